// ORACLE — TEST INFRASTRUCTURE ONLY.
// oracle/_ref/libvicalib_ref.so: the reference's OWN headers, compiled UNMODIFIED from /root/reference/include
//   include/vicalib/types.h                  GetGravityVector, ImuPoseT, ImuMeasurementT, ImuResidualT::IntegratePose /
//                                            GetPoseDerivative / IntegrateImu / IntegrateResidual (values, Jacobians, covariance)
//   include/vicalib/vicalibrator-utils.h     powi, dLog_dq, dqExp_dw, dq1q2_dq1/2, dqx_dq, dt1t2_dt1, dLog_dSE3
//   include/vicalib/interpolation-buffer.h   InterpolationBufferT (AddElement, HasElement, GetElement, GetNext, GetRange)
//   include/vicalib/ceres-cost-functions.h   IntegratePoseJet .. IntegrateResidualJet, SwitchedFullImuCostFunction,
//                                            ImuReprojectionCostFunctor
//   include/vicalib/local-param-se3.h        LocalParamSe3 / LocalParamSo3 (Plus, ComputeJacobian)
// against stand-ins for the libraries that are not in this image (oracle/ref_shim: Eigen, Sophus, ceres::Jet, glog,
// Calibu).  This file only marshals flat arrays in and out and restates the ~30 lines of glue that live inside
// ViCalibrator itself (vicalibrator.h:726-796, the body of UpdateImuWeights; vicalibrator.h:412-453, which functor is
// bound to which camera model) — ViCalibrator cannot be compiled here (ceres::Problem, HAL, Calibu's camera classes).
// Camera-model Project bodies are NOT reference text (Calibu is un-vendored): they come from oracle/camera_models.h.
// Built by `make -C oracle _ref` (only where /root/reference exists); used by tests/test_cpu_oracle_ref.py to pin the
// oracle restatement to the reference's source text.
#include <cstring>
#include <vector>

#include <ceres/ceres.h>
#include <glog/logging.h>

// Calibu-style Project bodies instantiated with ceres::Jet: oracle/camera_models.h calls vo::sqrt etc. qualified
namespace vo {
template <class T, int N> inline ceres::Jet<T, N> sqrt(const ceres::Jet<T, N>& x) { return ceres::sqrt(x); }
template <class T, int N> inline ceres::Jet<T, N> sin(const ceres::Jet<T, N>& x) { return ceres::sin(x); }
template <class T, int N> inline ceres::Jet<T, N> cos(const ceres::Jet<T, N>& x) { return ceres::cos(x); }
template <class T, int N> inline ceres::Jet<T, N> tan(const ceres::Jet<T, N>& x) { return ceres::tan(x); }
template <class T, int N> inline ceres::Jet<T, N> atan(const ceres::Jet<T, N>& x) { return ceres::atan(x); }
template <class T, int N> inline ceres::Jet<T, N> atan2(const ceres::Jet<T, N>& y, const ceres::Jet<T, N>& x) { return ceres::atan2(y, x); }
}  // namespace vo
#include "camera_models.h"

// include order of include/vicalib/vicalibrator.h:51-55
#include <vicalib/types.h>
#include <vicalib/vicalibrator-utils.h>
#include <vicalib/interpolation-buffer.h>
#include <vicalib/ceres-cost-functions.h>
#include <vicalib/local-param-se3.h>

namespace visual_inertial_calibration {
int debug_level_threshold = 0;
int debug_level = 0;
}  // namespace visual_inertial_calibration

using namespace visual_inertial_calibration;  // NOLINT
typedef InterpolationBufferT<ImuMeasurementT, double> ImuBuffer;

static void fill_buffer(ImuBuffer* buf, int n, const double* t, const double* w, const double* a) {
  for (int i = 0; i < n; ++i)
    buf->AddElement(ImuMeasurementT<double>(Eigen::Vector3d(w[3 * i], w[3 * i + 1], w[3 * i + 2]),
                                            Eigen::Vector3d(a[3 * i], a[3 * i + 1], a[3 * i + 2]), t[i]));
}
static Sophus::SE3d make_se3(const double* x) {  // [qx qy qz qw | t]: the parameter-block layout (local-param-se3.h:34)
  return Sophus::SE3d(Eigen::Map<const Sophus::SE3d>(x));
}

// ImuReprojectionCostFunctor<Cam>::operator() (ceres-cost-functions.h:342-377) as AutoDiffCostFunction<., 2, 7, 4, 3, K>
// evaluates it (vicalibrator.h:412-453), tangent Jacobian through LocalParamSe3 / LocalParamSo3:
// J 2 x 22 row-major = (6 pose | 3 w_ck | 3 p_ck | K, zero padded).  Cam::Project is the oracle's restatement of Calibu.
template <class Cam>
static int reproj(const double* x_wk, const double* q_ck, const double* p_ck, const double* intr, const double* pw, const double* pc,
                  double* r, double* J) {
  constexpr int K = Cam::K;
  ImuReprojectionCostFunctor<Cam> f(Eigen::Vector3d(pw[0], pw[1], pw[2]), Eigen::Vector2d(pc[0], pc[1]));
  if (!f(x_wk, q_ck, p_ck, intr, r)) return 1;
  if (!J) return 0;
  typedef ceres::Jet<double, 22> JetT;
  JetT jx[7], jq[4], jp[3], ji[K], jr[2];
  int k = 0;
  for (int i = 0; i < 7; ++i) jx[i] = JetT(x_wk[i], k++);
  for (int i = 0; i < 4; ++i) jq[i] = JetT(q_ck[i], k++);
  for (int i = 0; i < 3; ++i) jp[i] = JetT(p_ck[i], k++);
  for (int i = 0; i < K; ++i) ji[i] = JetT(intr[i], k++);
  if (!f(jx, jq, jp, ji, jr)) return 1;
  double L7[42], L4[12];
  LocalParamSe3().ComputeJacobian(x_wk, L7);
  LocalParamSo3().ComputeJacobian(q_ck, L4);
  for (int row = 0; row < 2; ++row) {
    double* o = J + row * 22;
    for (int c = 0; c < 22; ++c) o[c] = 0.0;
    for (int c = 0; c < 6; ++c)
      for (int q = 0; q < 7; ++q) o[c] += jr[row].v[q] * L7[q * 6 + c];
    for (int c = 0; c < 3; ++c)
      for (int q = 0; q < 4; ++q) o[6 + c] += jr[row].v[7 + q] * L4[q * 3 + c];
    for (int c = 0; c < 3 + K; ++c) o[9 + c] = jr[row].v[11 + c];
  }
  return 0;
}
extern "C" {

// InterpolationBufferT::GetRange (interpolation-buffer.h:208-226); out rows: time, w (3), a (3)
int ref_get_range(int n, const double* t, const double* w, const double* a, double t0, double t1, double ts, double* out, int max_n) {
  ImuBuffer buf(n);
  fill_buffer(&buf, n, t, w, a);
  aligned_vector<ImuMeasurementT<double> > meas;
  buf.GetRange(t0, t1, ts, &meas);
  const int m = static_cast<int>(meas.size());
  for (int i = 0; i < m && i < max_n; ++i) {
    out[7 * i] = meas[i].time;
    for (int k = 0; k < 3; ++k) { out[7 * i + 1 + k] = meas[i].w_[k]; out[7 * i + 4 + k] = meas[i].a_[k]; }
  }
  return m;
}

// SwitchedFullImuCostFunction::operator() (ceres-cost-functions.h:379-490) evaluated the way
// ceres::AutoDiffCostFunction<ViFullCost, 9, 7, 7, 3, 3, 2, 6, 6, 1> does (vicalibrator.h:618-619): Jet<double, 35>,
// one unit derivative per ambient coordinate, then J_tangent = J_ambient * LocalParamSe3::ComputeJacobian for the two
// pose blocks (vicalibrator.h:604).  J: 9 x 33 row-major = (pose2 6 | pose1 6 | v2 3 | v1 3 | g 2 | b 6 | sf 6 | ts 1)
int ref_imu_eval(int n, const double* t, const double* w, const double* a, double t_start, double t_end, const double* W,
                 int rotation_only, const double* x2, const double* x1, const double* v2, const double* v1, const double* g,
                 const double* b, const double* sf, double ts, double* r, double* J) {
  ImuBuffer buf(n);
  fill_buffer(&buf, n, t, w, a);
  Eigen::Matrix<double, 9, 9> Wm;
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) Wm(i, j) = W[i * 9 + j];
  const bool sw = rotation_only != 0;
  SwitchedFullImuCostFunction<double> f(&buf, t_start, t_end, Wm, &sw);
  {
    double rr[9];
    if (!f(x2, x1, v2, v1, g, b, sf, &ts, rr)) return 1;
    for (int i = 0; i < 9; ++i) r[i] = rr[i];
  }
  if (!J) return 0;
  typedef ceres::Jet<double, 35> JetT;
  JetT jx2[7], jx1[7], jv2[3], jv1[3], jg[2], jb[6], jsf[6], jts, jr[9];
  int k = 0;
  for (int i = 0; i < 7; ++i) jx2[i] = JetT(x2[i], k++);
  for (int i = 0; i < 7; ++i) jx1[i] = JetT(x1[i], k++);
  for (int i = 0; i < 3; ++i) jv2[i] = JetT(v2[i], k++);
  for (int i = 0; i < 3; ++i) jv1[i] = JetT(v1[i], k++);
  for (int i = 0; i < 2; ++i) jg[i] = JetT(g[i], k++);
  for (int i = 0; i < 6; ++i) jb[i] = JetT(b[i], k++);
  for (int i = 0; i < 6; ++i) jsf[i] = JetT(sf[i], k++);
  jts = JetT(ts, k++);
  if (!f(jx2, jx1, jv2, jv1, jg, jb, jsf, &jts, jr)) return 1;
  LocalParamSe3 lp;
  double L2[42], L1[42];  // 7 x 6 row-major
  lp.ComputeJacobian(x2, L2);
  lp.ComputeJacobian(x1, L1);
  for (int row = 0; row < 9; ++row) {
    double* o = J + row * 33;
    for (int c = 0; c < 6; ++c) {
      double s2 = 0, s1 = 0;
      for (int q = 0; q < 7; ++q) { s2 += jr[row].v[q] * L2[q * 6 + c]; s1 += jr[row].v[7 + q] * L1[q * 6 + c]; }
      o[c] = s2;
      o[6 + c] = s1;
    }
    for (int c = 0; c < 21; ++c) o[12 + c] = jr[row].v[14 + c];
  }
  return 0;
}

// The loop body of ViCalibrator::UpdateImuWeights (vicalibrator.h:726-796) for one interval.  Returns 0 and leaves
// W untouched when the interval has no measurements (:731-733).
int ref_update_weight(int n, const double* t, const double* w, const double* a, double t_start, double t_end, const double* x1,
                      const double* v1, const double* x2, const double* v2, const double* g, const double* b, const double* sf,
                      double ts, double sigma_g, double sigma_a, double* W, double* mahalanobis) {
  ImuBuffer buf(n);
  fill_buffer(&buf, n, t, w, a);
  aligned_vector<ImuMeasurementT<double> > measurements;
  buf.GetRange(t_start, t_end, ts, &measurements);  // cost_functor()->GetMeasurements(imu_.time_offset_, ...)
  if (measurements.size() == 0) return 0;
  PoseT<double> start_pose;
  const Sophus::SE3d t_w2 = make_se3(x2);
  Sophus::SE3d t_2w = t_w2.inverse();
  start_pose.t_wp_ = make_se3(x1);
  start_pose.v_w_ = Eigen::Vector3d(v1[0], v1[1], v1[2]);
  start_pose.time_ = measurements.front().time;
  aligned_vector<ImuPoseT<double> > poses_d;
  Eigen::Matrix<double, 10, 6> jb_q;
  Eigen::Matrix<double, 10, 10> c_imu_pose;
  c_imu_pose.setZero();
  const Eigen::Matrix<double, 6, 6> r((Eigen::Matrix<double, 6, 1>() << powi(sigma_g, 2), powi(sigma_g, 2), powi(sigma_g, 2),
                                       powi(sigma_a, 2), powi(sigma_a, 2), powi(sigma_a, 2))
                                          .finished()
                                          .asDiagonal());
  Vector6d biases, scale_factors;
  for (int i = 0; i < 6; ++i) { biases[i] = b[i]; scale_factors[i] = sf[i]; }
  const Eigen::Matrix<double, 2, 1> gdir(g[0], g[1]);
  ImuPoseT<double> imu_pose = ImuResidualT<double>::IntegrateResidual(
      ImuPoseT<double>(start_pose), measurements, biases.head<3>(), biases.tail<3>(), scale_factors,
      GetGravityVector(gdir, gravity()), poses_d, &jb_q, nullptr, &c_imu_pose, &r);
  const Eigen::Matrix<double, 6, 7> dlog_dse3 = dLog_dSE3(imu_pose.t_wp_ * t_2w);
  const Eigen::Matrix<double, 7, 7> dt1t2_dt2 = dt1t2_dt1(imu_pose.t_wp_, t_2w);
  const Eigen::Matrix<double, 6, 7> dse3t1t2_dt2 = dlog_dse3 * dt1t2_dt2;
  Eigen::Matrix<double, 9, 10> dse3t1t2v_dt2;
  dse3t1t2v_dt2.setZero();
  dse3t1t2v_dt2.topLeftCorner<6, 7>() = dse3t1t2_dt2;
  dse3t1t2v_dt2.bottomRightCorner<3, 3>().setIdentity();
  const Eigen::Matrix<double, 9, 9> cov = (dse3t1t2v_dt2 * c_imu_pose * dse3t1t2v_dt2.transpose()).inverse();
  Eigen::Matrix<double, 9, 1> residuals;
  residuals.head<6>() = Sophus::SE3d::log(imu_pose.t_wp_ * t_2w);
  residuals.tail<3>() = imu_pose.v_w_ - Eigen::Vector3d(v2[0], v2[1], v2[2]);
  const Eigen::Matrix<double, 1, 1> dist = residuals.transpose() * cov * residuals;
  if (mahalanobis) *mahalanobis = dist(0, 0);
  const Eigen::Matrix<double, 9, 9> ws = cov.sqrt();
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) W[i * 9 + j] = ws(i, j);
  return 1;
}

// ImuResidualT::IntegrateResidual with Jacobians and covariance (types.h:611-687): end pose (p3 q4 v3), dpose/db 10 x 6,
// dpose/dpose 10 x 10, covariance 10 x 10, all row-major
int ref_integrate(int n, const double* t, const double* w, const double* a, double t_start, double t_end, const double* x1,
                  const double* v1, const double* g, const double* b, const double* sf, double ts, double sigma_g, double sigma_a,
                  double* y_out, double* dy_db, double* dy_dy, double* cov) {
  ImuBuffer buf(n);
  fill_buffer(&buf, n, t, w, a);
  aligned_vector<ImuMeasurementT<double> > measurements;
  buf.GetRange(t_start, t_end, ts, &measurements);
  if (measurements.size() == 0) return 0;
  PoseT<double> start_pose;
  start_pose.t_wp_ = make_se3(x1);
  start_pose.v_w_ = Eigen::Vector3d(v1[0], v1[1], v1[2]);
  start_pose.time_ = measurements.front().time;
  aligned_vector<ImuPoseT<double> > poses;
  Eigen::Matrix<double, 10, 6> jb;
  Eigen::Matrix<double, 10, 10> jy, c;
  c.setZero();
  Eigen::Matrix<double, 6, 6> r;
  r.setZero();
  for (int i = 0; i < 3; ++i) { r(i, i) = powi(sigma_g, 2); r(3 + i, 3 + i) = powi(sigma_a, 2); }
  Vector6d biases, scale_factors;
  for (int i = 0; i < 6; ++i) { biases[i] = b[i]; scale_factors[i] = sf[i]; }
  const Eigen::Matrix<double, 2, 1> gdir(g[0], g[1]);
  const ImuPoseT<double> y = ImuResidualT<double>::IntegrateResidual(ImuPoseT<double>(start_pose), measurements, biases.head<3>(),
                                                                    biases.tail<3>(), scale_factors, GetGravityVector(gdir, gravity()),
                                                                    poses, &jb, &jy, &c, &r);
  const Eigen::Matrix<double, 10, 1> yv = y;
  for (int i = 0; i < 10; ++i) y_out[i] = yv[i];
  for (int i = 0; i < 10; ++i) {
    for (int j = 0; j < 6; ++j) dy_db[i * 6 + j] = jb(i, j);
    for (int j = 0; j < 10; ++j) { dy_dy[i * 10 + j] = jy(i, j); cov[i * 10 + j] = c(i, j); }
  }
  return static_cast<int>(measurements.size());
}

// LocalParamSe3 / LocalParamSo3 (local-param-se3.h:14-26, 28-91, 107-119, 121-157)
void ref_se3_plus(const double* x, const double* d, double* out) { LocalParamSe3().Plus(x, d, out); }
void ref_se3_jacobian(const double* x, double* J42) { LocalParamSe3().ComputeJacobian(x, J42); }
void ref_so3_plus(const double* x, const double* d, double* out) { LocalParamSo3().Plus(x, d, out); }
void ref_so3_jacobian(const double* x, double* J12) { LocalParamSo3().ComputeJacobian(x, J12); }
void ref_gravity(const double* g2, double* out) {
  const Eigen::Vector3d v = GetGravityVector(Eigen::Matrix<double, 2, 1>(g2[0], g2[1]), gravity());
  for (int i = 0; i < 3; ++i) out[i] = v[i];
}
// the hand-derived derivative tables of vicalibrator-utils.h, row-major
void ref_dLog_dSE3(const double* x, double* out42) {
  const Eigen::Matrix<double, 6, 7> m = dLog_dSE3(make_se3(x));
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 7; ++j) out42[i * 7 + j] = m(i, j);
}
void ref_dqExp_dw(const double* w3, double* out12) {
  const Eigen::Matrix<double, 4, 3> m = dqExp_dw<double>(Eigen::Vector3d(w3[0], w3[1], w3[2]));
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) out12[i * 3 + j] = m(i, j);
}

int ref_reproj(int model, const double* x_wk, const double* q_ck, const double* p_ck, const double* intr, const double* pw,
               const double* pc, double* r, double* J) {
  switch (model) {
    case 0: return reproj<vo::LinearCam>(x_wk, q_ck, p_ck, intr, pw, pc, r, J);
    case 1: return reproj<vo::FovCam>(x_wk, q_ck, p_ck, intr, pw, pc, r, J);
    case 2: return reproj<vo::Poly2Cam>(x_wk, q_ck, p_ck, intr, pw, pc, r, J);
    case 3: return reproj<vo::Poly3Cam>(x_wk, q_ck, p_ck, intr, pw, pc, r, J);
    case 4: return reproj<vo::Kb4Cam>(x_wk, q_ck, p_ck, intr, pw, pc, r, J);
  }
  return 2;
}

}  // extern "C"
