// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// Problem container + Ceres-structured evaluation and trust-region solve, restating
//   ViCalibrator::SetupProblem       vicalibrator.h:548-679   (parameter blocks, constant masks)
//   ceres::Solve w/ solver_options_  vicalibrator.h:141-152,956  (trust region; Ceres not vendored:
//        loop / LM / dogleg restated from the published algorithm)
//   per-camera Evaluate              vicalibrator.h:958-971
//   RemoveOutliers                   vicalibrator.h:859-916
// One dual-number evaluation per residual block (like AutoDiffCostFunction), tangent-space
// Jacobians through the LocalParam* Jacobians, loss "corrector" scaling, block-sparse normal
// equations (frame blocks block-(tri)diagonal + dense arrow of globals), block Cholesky.
#ifndef VICALIB_ORACLE_PROBLEM_H_
#define VICALIB_ORACLE_PROBLEM_H_

#include <cstdint>
#include <string>
#include <vector>

#include "cost_functors.h"

namespace vo {

struct Flags {
  int inertial = 0;        // is_inertial_active_ && FLAGS_calibrate_imu
  int rotation_only = 0;   // optimize_rotation_only_
  int bias_active = 0;     // is_bias_active_
  int scale_active = 0;    // is_scale_factor_active_
  int optimize_ts = 0;     // optimize_time_offset_
  int fix_intrinsics = 0;  // fix_intrinsics_
  int visual = 1;          // is_visual_active_
  double visual_mult = 1;  // residual-block multiplicity quirk
  double imu_mult = 1;
};

struct Options {
  int max_iters = 200;              // FLAGS_max_iters (vicalib-engine.cc:94)
  double function_tol = 1e-6;       // vicalibrator.h:149
  double gradient_tol = 1e-10;      // Ceres default
  double param_tol = 1e-8;          // Ceres default
  double init_radius = 1e4;         // Ceres default
  double max_radius = 1e16;
  double min_radius = 1e-32;
  double min_rel_decrease = 1e-3;
  int strategy = 0;                 // 0 = LM (north_star), 1 = DOGLEG (vicalibrator.h:151)
  int jacobi_scaling = 1;
  int num_threads = 1;
  int update_imu_weights = 1;       // callback + pre-solve UpdateImuWeights (vicalibrator.h:691,955)
  double callback_gnorm_stop = 1e-9;  // vicalibrator.h:713-717
};

struct IterationRow {
  int iteration;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, rho, radius;
  int accepted;
};

struct Summary {
  int iterations = 0;          // accepted + rejected steps
  int successful_steps = 0;
  double initial_cost = 0, final_cost = 0;
  int termination = 0;         // 0 no-convergence(max iters), 1 function tol, 2 gradient tol,
                               // 3 param tol, 4 callback stop, 5 radius underflow
  int num_residuals = 0;
  std::vector<IterationRow> rows;
};

struct NormalEq {
  int nf = 0, fd = 6, G = 0;
  std::vector<double> B;   // nf * fd*fd   (row-major)
  std::vector<double> U;   // nf * fd*fd   H[f-1, f]  (U[0] unused)
  std::vector<double> E;   // nf * fd*G
  std::vector<double> gf;  // nf * fd
  std::vector<double> C;   // G*G
  std::vector<double> gc;  // G
  double cost = 0;
  void Resize(int nf_, int fd_, int G_);
  void Zero();
};

class Problem {
 public:
  // ---- data (set through the C API)
  int n_cams = 0;
  std::vector<int> model;          // per cam
  std::vector<double> intr;        // n_cams*10
  std::vector<double> q_ck, p_ck;  // n_cams*4 / *3
  int n_frames = 0;
  std::vector<double> T_wp, v_w, ftime;  // 7n, 3n, n
  int64_t n_obs = 0;
  std::vector<int32_t> obs_frame, obs_cam;
  std::vector<double> p_w, p_c;    // 3n, 2n
  std::vector<uint8_t> obs_active; // outlier mask
  InterpolationBuffer imu;
  double g[2] = {0, 0}, b[6] = {0, 0, 0, 0, 0, 0}, sf[6] = {1, 1, 1, 1, 1, 1}, ts = 0;
  double sigma_g = 5.3088444e-5, sigma_a = 0.001883649;  // types.h:34-35
  std::vector<double> w_sqrt;      // (n_frames-1)*81, init 500*I (vicalibrator.h:616)
  Flags flags;
  Options opts;

  // ---- derived layout
  int FrameDim() const { return flags.inertial ? 9 : 6; }
  int NumGlobals() const;
  int CamOffset(int c) const;      // start of [w_ck(3) p_ck(3) intr(K)] in the global vector
  int ImuOffset() const;           // start of [g2 b6 sf6 ts1]
  void GlobalMask(std::vector<double>* mask) const;
  int NumResiduals() const;

  void ResetImuWeights();
  // vicalibrator.h:723-799
  void UpdateImuWeights();

  // residuals (optionally Jacobians) of one observation / one IMU interval, tangent space.
  // J layouts: reprojection 2 x (6 pose | 3 w_ck | 3 p_ck | K); IMU 9 x 33 =
  //   (pose2 6 | pose1 6 | v2 3 | v1 3 | g 2 | b 6 | sf 6 | ts 1)  (ceres-cost-functions.h:403-406)
  void EvalReprojection(int64_t i, double r[2], double* J) const;
  void EvalImu(int interval, double r[9], double* J) const;

  // cost = 1/2 sum rho(|r|^2) (multiplicities applied); ne may be null (cost only)
  double Evaluate(NormalEq* ne) const;
  // vicalibrator.h:958-966 / 873-887: per-camera residuals, loss not applied
  double EvaluateCamera(int cam, std::vector<double>* residuals) const;
  // vicalibrator.h:859-916. returns number removed
  int RemoveOutliers(const std::vector<double>& rmse, double threshold);

  Summary Solve();

  // state update x (+) delta, delta = [frames nf*fd | globals G]
  void Plus(const std::vector<double>& delta);
  struct State {
    std::vector<double> intr, q_ck, p_ck, T_wp, v_w;
    double g[2], b[6], sf[6], ts;
  };
  State Save() const;
  void Restore(const State& s);
  double StateNorm() const;
};

// Solve (H + diag(D2)) x = -g for the arrow system in ne, with column scaling `scale`
// (Jacobi) applied symmetrically. D2 is in the scaled space. Returns false if not PD.
// delta is in the *scaled* space (caller multiplies by scale).
bool SolveArrow(const NormalEq& ne, const std::vector<double>& scale, const std::vector<double>& D2,
                std::vector<double>* delta_scaled);
// y = H_scaled * x  (no damping)
void ArrowMatVec(const NormalEq& ne, const std::vector<double>& scale, const std::vector<double>& x,
                 std::vector<double>* y);
void ArrowDiagonal(const NormalEq& ne, std::vector<double>* diag);
void ArrowGradient(const NormalEq& ne, std::vector<double>* grad);
// dense reference path for tests: assembles the full H (n x n, row-major)
void ArrowToDense(const NormalEq& ne, std::vector<double>* H, std::vector<double>* grad);

}  // namespace vo
#endif
