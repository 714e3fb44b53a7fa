// Drives the C++ ViCalibrator mirror (vicalib_b200/host/vicalibrator.h) the way VicalibTask does
// (vicalib-task.cc:128-147, 227-245, 339-363, 689): AddCamera / AddFrame / AddObservation /
// AddImuMeasurements, SetOptimizationFlags, Start, poll IsRunning, WriteCameraModels.
// Input: a flat binary problem written by tests/test_gpu_host_mirror.py.  Output: result text file.
#include <unistd.h>

#include <cstdint>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <vector>

#include "../../vicalib_b200/host/vicalib_task_checks.h"

using namespace visual_inertial_calibration;

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s problem.bin result.txt cameras.xml\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  // header: n_cams n_frames n_obs n_imu inertial has_initial_guess max_iters mode
  //   mode bit 0: emulate the residual-block duplication; bit 1: LEVENBERG_MARQUARDT instead of the reference's DOGLEG;
  //   bit 2: run, Clear(), load the problem again and run once more (a reused object must start from scratch);
  //   bit 3: remove_outliers; bit 4: the reference's default function tolerance 1e-6 (else 1e-10)
  std::vector<int64_t> hd = rd<int64_t>(f, 8);
  const int nc = static_cast<int>(hd[0]), nf = static_cast<int>(hd[1]);
  const int64_t nobs = hd[2], nimu = hd[3];
  const bool inertial = hd[4] != 0, has_guess = hd[5] != 0;
  std::vector<int32_t> model = rd<int32_t>(f, nc);
  std::vector<double> intr = rd<double>(f, 10 * nc), q = rd<double>(f, 4 * nc), p = rd<double>(f, 3 * nc);
  std::vector<double> T = rd<double>(f, 7 * nf), tm = rd<double>(f, nf);
  std::vector<int32_t> of = rd<int32_t>(f, nobs), oc = rd<int32_t>(f, nobs);
  std::vector<double> pw = rd<double>(f, 3 * nobs), pc = rd<double>(f, 2 * nobs);
  std::vector<double> it = rd<double>(f, nimu), iw = rd<double>(f, 3 * nimu), ia = rd<double>(f, 3 * nimu);
  std::vector<double> g = rd<double>(f, 2);
  fclose(f);

  static const char* kTypes[5] = {"calibu_fu_fv_u0_v0", "calibu_fu_fv_u0_v0_w", "calibu_fu_fv_u0_v0_k1_k2",
                                  "calibu_fu_fv_u0_v0_k1_k2_k3", "calibu_fu_fv_u0_v0_kb4"};
  static const int kK[5] = {4, 5, 6, 7, 8};
  CalibratorFlags flags;
  flags.calibrate_imu = inertial;
  flags.max_iters = static_cast<int>(hd[6]);
  flags.remove_outliers = (hd[7] & 8) != 0;
  if (hd[7] & 2) flags.trust_region_strategy = 0;
  ViCalibrator cal(flags);
  std::vector<CameraAndPose> input_cameras;
  for (int pass = 0; pass < ((hd[7] & 4) ? 2 : 1); ++pass) {
  if (pass > 0) cal.Clear();
  cal.SetEmulateBlockDuplication((hd[7] & 1) != 0);
  input_cameras.clear();
  for (int c = 0; c < nc; ++c) {
    std::vector<double> params0(intr.begin() + 10 * c, intr.begin() + 10 * c + kK[model[c]]);
    input_cameras.push_back(CameraAndPose(std::shared_ptr<CameraInterface>(new CameraInterface(kTypes[model[c]], 640, 480, params0)),
                                          SE3d(&q[4 * c], &p[3 * c])));
  }
  for (int c = 0; c < nc; ++c) {
    std::vector<double> params(intr.begin() + 10 * c, intr.begin() + 10 * c + kK[model[c]]);
    std::shared_ptr<CameraInterface> cam(new CameraInterface(kTypes[model[c]], 640, 480, params));
    cal.AddCamera(cam, SE3d(&q[4 * c], &p[3 * c]));
  }
  for (int fr = 0; fr < nf; ++fr) cal.AddFrame(SE3d(&T[7 * fr], &T[7 * fr + 4]), tm[fr]);
  for (int64_t i = 0; i < nobs; ++i)
    cal.AddObservation(of[i], oc[i], Vector3d{{pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]}}, Vector2d{{pc[2 * i], pc[2 * i + 1]}}, tm[of[i]]);
  for (int64_t i = 0; i < nimu; ++i)
    cal.AddImuMeasurements(Vector3d{{iw[3 * i], iw[3 * i + 1], iw[3 * i + 2]}}, Vector3d{{ia[3 * i], ia[3 * i + 1], ia[3 * i + 2]}}, it[i]);
  cal.SetFunctionTolerance((hd[7] & 16) ? 1e-6 : 1e-10);
  if (has_guess) cal.SetOptimizationFlags(true, inertial, false, true);  // vicalib-task.cc:227-235
  cal.Start();
  while (cal.IsRunning()) usleep(2000);  // the reference polls every 30 ms (vicalib-engine.cc:388-400)
  cal.Stop();
  }
  cal.WriteCameraModels(argv[3]);

  FILE* o = fopen(argv[2], "w");
  std::fprintf(o, "solves %d iterations %u mse %.17g ts %.17g\n", cal.num_solves(), cal.GetNumIterations(), cal.MeanSquaredError(),
               cal.time_offset());
  for (int c = 0; c < nc; ++c) {
    std::fprintf(o, "cam %d rmse %.17g params", c, cal.GetCameraProjRMSE()[c]);
    for (double v : cal.GetCamera(c).camera->GetParams()) std::fprintf(o, " %.17g", v);
    std::fprintf(o, " T_ck");
    for (int k = 0; k < 7; ++k) std::fprintf(o, " %.17g", cal.GetCamera(c).T_ck.d[k]);
    std::fprintf(o, "\n");
  }
  Vector6d b = cal.GetBiases(), sf = cal.GetScaleFactor();
  std::fprintf(o, "biases");
  for (double v : b) std::fprintf(o, " %.17g", v);
  std::fprintf(o, "\nscale");
  for (double v : sf) std::fprintf(o, " %.17g", v);
  std::fprintf(o, "\n");
  // cameras.xml round trip through the rig reader (calibu::ReadXmlRig stand-in): exact for %.17g output
  double worst = 0.0;
  const auto rig = ReadXmlRig(argv[3]);
  if (static_cast<int>(rig.size()) != nc) worst = 1e300;
  for (int c = 0; c < nc && worst < 1e300; ++c) {
    const CameraInterface& w = *cal.GetCamera(c).camera;
    if (rig[c]->Type() != w.Type() || rig[c]->Width() != w.Width() || rig[c]->GetParams().size() != w.GetParams().size()) { worst = 1e300; break; }
    for (size_t k = 0; k < w.GetParams().size(); ++k) worst = std::max(worst, std::fabs(rig[c]->GetParams()[k] - w.GetParams()[k]));
    double Ma[12], Mb[12];
    rig[c]->Pose().matrix3x4(Ma);
    w.Pose().matrix3x4(Mb);
    for (int k = 0; k < 12; ++k) worst = std::max(worst, std::fabs(Ma[k] - Mb[k]));
    for (int k = 0; k < 9; ++k) worst = std::max(worst, std::fabs(rig[c]->RDF()[k] - w.RDF()[k]));
  }
  std::fprintf(o, "xml_roundtrip %.3g\n", worst);
  // VicalibTask::IsSuccessful (vicalib-task.cc:837-863) with the flag defaults
  const std::vector<double> max_err(nc, SuccessThresholds().max_reprojection_error);
  std::fprintf(o, "success %d %d\n", IsSuccessful(cal, max_err, false, input_cameras, Vector6d{{0, 0, 0, 0, 0, 0}}) ? 1 : 0,
               IsSuccessful(cal, max_err, true, input_cameras, Vector6d{{0, 0, 0, 0, 0, 0}}) ? 1 : 0);
  // GetIntegrationPoses (vicalibrator.h:508-533): end pose of the first interval's integration vs the second frame
  if (inertial && nf > 1) {
    const std::vector<ImuPose> ip = cal.GetIntegrationPoses(0);
    double dp = -1;
    if (!ip.empty()) {
      dp = 0;
      for (int k = 0; k < 3; ++k) dp = std::max(dp, std::fabs(ip.back().t_wp_.d[4 + k] - cal.GetFrame(1)->t_wp_.d[4 + k]));
    }
    std::fprintf(o, "integration_poses %zu %.6g %.17g\n", ip.size(), dp, ip.empty() ? 0.0 : ip.back().time_);
  }
  // GetSolutionCovariance (vicalibrator.h:802-857): G, asymmetry, smallest / largest diagonal entry of the active blocks
  try {
    int G = 0;
    const std::vector<double> cov = cal.GetSolutionCovariance(&G);
    double asym = 0, dmin = 1e300, dmax = 0;
    for (int r = 0; r < G; ++r) {
      for (int c = 0; c < G; ++c) asym = std::max(asym, std::fabs(cov[r * G + c] - cov[c * G + r]));
      if (cov[r * G + r] != 0.0) { dmin = std::min(dmin, cov[r * G + r]); dmax = std::max(dmax, cov[r * G + r]); }
    }
    std::fprintf(o, "covariance %d %.6g %.6g %.6g", G, asym, dmin, dmax);
    for (int r = 0; r < G; ++r) std::fprintf(o, " %.17g", cov[r * G + r]);
    std::fprintf(o, "\n");
  } catch (const std::exception& e) {  // a short inertial trajectory leaves J^T J singular to working precision
    std::fprintf(o, "covariance -1\n");
  }
  // pose seeds the way VicalibTask::AddSuperFrame gets them (vicalib-task.cc:322-349): PnP on camera 0's corners of the
  // first frames, against the solved frame poses
  {
    const int nv = std::min(nf, 6);
    std::vector<ViCalibrator::PnpView> views(nv);
    for (int v = 0; v < nv; ++v) views[v].camera_id = 0;
    for (int64_t i = 0; i < nobs; ++i)
      if (oc[i] == 0 && of[i] < nv) {
        views[of[i]].pixels.push_back(Vector2d{{pc[2 * i], pc[2 * i + 1]}});
        views[of[i]].target_points.push_back(Vector3d{{pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]}});
      }
    std::vector<bool> ok;
    const std::vector<SE3d> seeds = cal.InitialPosesFromTarget(views, &ok);
    double worst = 0;
    int n_ok = 0;
    for (int v = 0; v < nv; ++v) {
      if (!ok[v]) continue;
      ++n_ok;
      for (int k = 0; k < 3; ++k) worst = std::max(worst, std::fabs(seeds[v].d[4 + k] - cal.GetFrame(v)->t_wp_.d[4 + k]));
    }
    std::fprintf(o, "pnp_seeds %d %d %.6g\n", nv, n_ok, worst);
  }
  fclose(o);
  return 0;
}
