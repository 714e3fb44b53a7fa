/* vcgpu.h — C-ABI of the B200 calibration-solve library (libvcgpu.so).
 *
 * Drop-in boundary for the path that sits behind `ceres::Solve` in the reference:
 * it replaces ceres::Problem construction + ceres::Solve + Problem::Evaluate as used by
 * ViCalibrator (include/vicalib/vicalibrator.h:152, 548-679, 859-916, 956-971).  Plain
 * pointers and sizes only; every entry point returns 0 on success and a negative code on
 * error (message via vcgpu_last_error); nothing throws or aborts across this boundary.
 *
 * All arithmetic is FP64.  Quaternions are (x, y, z, w); SE3 is [q(4) | t(3)], the storage
 * the reference's parameter blocks use (local-param-se3.h:34, vicalibrator.h:460-461).
 */
#ifndef VCGPU_H_
#define VCGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vcgpu_handle vcgpu_handle;

/* camera model ids; intrinsics are fu, fv, u0, v0, <distortion...> (vicalib-engine.cc:205-250) */
enum {
  VCGPU_CAM_LINEAR = 0, /* K=4  calibu_fu_fv_u0_v0         vicalibrator.h:448 */
  VCGPU_CAM_FOV = 1,    /* K=5  calibu_fu_fv_u0_v0_w       vicalibrator.h:412 */
  VCGPU_CAM_POLY2 = 2,  /* K=6  calibu_fu_fv_u0_v0_k1_k2   vicalibrator.h:419 */
  VCGPU_CAM_POLY3 = 3,  /* K=7  ..._k1_k2_k3               vicalibrator.h:427 */
  VCGPU_CAM_KB4 = 4     /* K=8  calibu_fu_fv_u0_v0_kb4     vicalibrator.h:441 */
};
#define VCGPU_MAX_INTR 10 /* stride of the intrinsics array per camera */

enum {
  VCGPU_OK = 0,
  VCGPU_ERR_INVALID = -1,  /* bad argument / call order (reference: glog CHECK, vicalibrator.h:254,356,396) */
  VCGPU_ERR_CUDA = -2,     /* CUDA runtime failure (no CPU fallback exists) */
  VCGPU_ERR_NUMERIC = -3,  /* normal equations not positive definite even after damping */
  VCGPU_ERR_COMM = -4      /* NCCL failure */
};

typedef struct {
  int device; /* CUDA device ordinal, -1 = current device */
} vcgpu_config;

/* Which parameter blocks are variable — ViCalibrator::SetupProblem's constant masks
 * (vicalibrator.h:572-592, 651-676) and SetOptimizationFlags (vicalibrator.h:252-260). */
typedef struct {
  int inertial;         /* is_inertial_active_ && FLAGS_calibrate_imu: IMU residuals on, cam0 q_ck variable */
  int rotation_only;    /* optimize_rotation_only_: IMU residual rows 0-2,6-8 zeroed, gravity & cam0 p_ck constant */
  int bias_active;      /* is_bias_active_ */
  int scale_active;     /* is_scale_factor_active_ */
  int optimize_ts;      /* optimize_time_offset_ */
  int fix_intrinsics;   /* fix_intrinsics_ (vicalibrator.h:346,591) */
  int visual;           /* is_visual_active_ */
  double visual_mult;   /* residual-block multiplicity of the staged flow (vicalibrator.h:641-656 re-adds blocks) */
  double imu_mult;
} vcgpu_flags;

/* ceres::Solver::Options as ViCalibrator sets them (vicalibrator.h:141-152) */
typedef struct {
  int max_iters;           /* FLAGS_max_iters, default 200 (vicalib-engine.cc:94) */
  double function_tol;     /* 1e-6 (vicalibrator.h:149) */
  double gradient_tol;     /* Ceres default 1e-10 */
  double param_tol;        /* Ceres default 1e-8 */
  double init_radius;      /* Ceres default 1e4 */
  int strategy;            /* 0 = LEVENBERG_MARQUARDT, 1 = DOGLEG (vicalibrator.h:151) */
  int jacobi_scaling;      /* Ceres default on */
  int update_imu_weights;  /* run UpdateImuWeights before the solve and after every iteration (vicalibrator.h:691,955) */
  int update_state_every_iteration; /* copy the state to the host mirrors before each callback (vicalibrator.h:148) */
} vcgpu_options;

/* ceres::IterationSummary fields the reference's callback reads (vicalibrator.h:696-714) */
typedef struct {
  int iteration;
  int step_is_successful;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
} vcgpu_iteration;

/* return nonzero to stop (ceres::SOLVER_TERMINATE_SUCCESSFULLY, vicalibrator.h:717) */
typedef int (*vcgpu_iter_cb)(const vcgpu_iteration* it, void* user);

enum {
  VCGPU_TERM_NO_CONVERGENCE = 0, /* max iterations */
  VCGPU_TERM_FUNCTION_TOL = 1,
  VCGPU_TERM_GRADIENT_TOL = 2,
  VCGPU_TERM_PARAM_TOL = 3,
  VCGPU_TERM_CALLBACK = 4,       /* callback asked to stop, or gradient norm < 1e-9 rule (vicalibrator.h:713-717) */
  VCGPU_TERM_RADIUS = 5
};

/* ceres::Solver::Summary subset (vicalibrator.h:975-976) */
typedef struct {
  int iterations;
  int successful_steps;
  int termination;
  int num_residuals;
  double initial_cost;
  double final_cost;
  double device_seconds; /* CUDA-event time of the iteration loop */
  int kernel_launches;   /* kernels this library launched during the solve */
} vcgpu_summary;

int vcgpu_create(const vcgpu_config* cfg, vcgpu_handle** out);
int vcgpu_destroy(vcgpu_handle* h);
const char* vcgpu_last_error(const vcgpu_handle* h);

/* ---- problem upload (host buffers are copied; caller keeps ownership) --------------------
 * replaces AddParameterBlock / AddResidualBlock (vicalibrator.h:559-604, 628-632, 645-654) */
int vcgpu_set_cameras(vcgpu_handle* h, int n_cams, const int32_t* model, const double* intr /*n*10*/,
                      const double* q_ck /*n*4*/, const double* p_ck /*n*3*/);
int vcgpu_set_frames(vcgpu_handle* h, int n_frames, const double* T_wp /*n*7*/, const double* v_w /*n*3*/,
                     const double* time /*n*/);
/* any order; the library groups by (camera, frame) internally and keeps the caller's order for outputs */
int vcgpu_set_observations(vcgpu_handle* h, int64_t n_obs, const int32_t* frame_id, const int32_t* cam_id,
                           const double* p_w /*n*3*/, const double* p_c /*n*2*/);
/* strictly increasing timestamps (vicalibrator.h:373-378) */
int vcgpu_set_imu(vcgpu_handle* h, int n, const double* t, const double* w /*n*3*/, const double* a /*n*3*/,
                  double sigma_g, double sigma_a);
int vcgpu_set_imu_params(vcgpu_handle* h, const double g[2], const double b[6], const double sf[6], double ts);
int vcgpu_set_flags(vcgpu_handle* h, const vcgpu_flags* flags);
int vcgpu_set_options(vcgpu_handle* h, const vcgpu_options* opts);
void vcgpu_default_flags(vcgpu_flags* flags);
void vcgpu_default_options(vcgpu_options* opts);

/* host mirrors updated before every callback when update_state_every_iteration is set, and at
 * the end of vcgpu_solve; any pointer may be NULL (vicalibrator.h:148: the GUI thread polls them) */
int vcgpu_register_mirrors(vcgpu_handle* h, double* intr, double* q_ck, double* p_ck, double* T_wp,
                           double* v_w, double* g, double* b, double* sf, double* ts);

/* ---- the hot path ---------------------------------------------------------------------- */
/* ceres::Solve(solver_options_, problem_, &summary)  (vicalibrator.h:956) */
int vcgpu_solve(vcgpu_handle* h, vcgpu_iter_cb cb, void* user, vcgpu_summary* out);
/* run exactly n trust-region iterations with no convergence test and no host round trip per
 * iteration (benchmark entry: LM iterations / second) */
int vcgpu_iterate(vcgpu_handle* h, int n_iterations, vcgpu_summary* out);

/* Problem::Evaluate with residual_blocks = camera `cam` (or all visual blocks if cam < 0),
 * apply_loss_function = false: cost = 1/2 sum |r|^2, residuals in caller observation order
 * (2 per active observation of that camera)  (vicalibrator.h:873-887, 958-966) */
int vcgpu_evaluate(vcgpu_handle* h, int cam, double* cost, double* residuals_or_null, int64_t* n_blocks);
/* total robustified cost 1/2 sum rho(|r|^2) over all active residual blocks at the current state */
int vcgpu_cost(vcgpu_handle* h, double* cost);
/* RemoveOutliers (vicalibrator.h:859-916): drops observations with |r| > threshold * rmse[cam] */
int vcgpu_remove_outliers(vcgpu_handle* h, const double* rmse, double threshold, int64_t* n_removed);
int vcgpu_get_obs_active(vcgpu_handle* h, uint8_t* active /*n_obs, caller order*/);
/* UpdateImuWeights (vicalibrator.h:723-799) */
int vcgpu_update_imu_weights(vcgpu_handle* h);
int vcgpu_get_imu_weights(vcgpu_handle* h, double* w_sqrt /*(n_frames-1)*81*/);
int vcgpu_set_imu_weights(vcgpu_handle* h, const double* w_sqrt);

int vcgpu_get_state(vcgpu_handle* h, double* intr, double* q_ck, double* p_ck, double* T_wp, double* v_w,
                    double* g, double* b, double* sf, double* ts);
int vcgpu_num_residuals(vcgpu_handle* h, int* out);
int vcgpu_frame_dim(vcgpu_handle* h, int* out);   /* 6, or 9 with inertial terms */
int vcgpu_num_globals(vcgpu_handle* h, int* out); /* sum_c (6 + K_c) (+15 with inertial terms) */

/* Batched pose initialisation — replaces the per-frame, per-camera `PosePnPRansac(camera, ellipses, target.Circles3D(),
 * ellipse_target_map, 0, 0, &t_cw)` of VicalibTask (src/vicalib-task.cc:322-325; Calibu + OpenCV solvePnP, un-vendored).
 * View v uses camera cam_id[v] (model / intrinsics from vcgpu_set_cameras) and the correspondences
 * [start[v], start[v] + count[v]) of pix (pixels) / pw (target points on the z = 0 plane, vicalib-task.cc:357-358).
 * robust_its > 0: RANSAC over 4-point homographies with inlier tolerance robust_tol (normalised image coordinates);
 * the reference calls it with 0, 0.  Out: T_cw[v] = (qx qy qz qw tx ty tz), p_c = R p_w + t; rmse[v] in normalised
 * coordinates; n_used[v] = points in the final fit (0: fewer than 4 usable points, T_cw[v] = identity).
 * The frame pose the solve is seeded with is T_wp = T_cw^-1 * T_ck (vicalib-task.cc:341-349). */
int vcgpu_pose_pnp_ransac(vcgpu_handle* h, int n_views, const int32_t* cam_id, const int64_t* start, const int32_t* count,
                          const double* pix, const double* pw, int robust_its, double robust_tol, double* T_cw /*n_views*7*/,
                          double* rmse_or_null, int32_t* n_used_or_null);

/* GetSolutionCovariance (vicalibrator.h:802-857; compiled out upstream behind COMPUTE_VICALIB_COVARIANCE): the
 * [globals, globals] block of (J^T J)^-1 at the current state, tangent space, G x G row-major with G = vcgpu_num_globals:
 * per camera (w_ck 3 | p_ck 3 | intrinsics K), then with inertial terms (g 2 | b 6 | sf 6 | ts 1).  Rows / columns of
 * constant parameter blocks are zero (ceres::Covariance convention). */
int vcgpu_get_covariance(vcgpu_handle* h, double* cov);

/* ---- inspection hooks used by the parity tests (what AutoDiffCostFunction::Evaluate and the
 * normal-equation build produce inside Ceres) ------------------------------------------------ */
/* loss-free residuals and tangent-space Jacobians per observation, caller order.
 * J: n_obs * 2 * 22, row-major 2 x (6 pose | 3 w_ck | 3 p_ck | K), zero padded to 22 columns. */
int vcgpu_eval_reproj(vcgpu_handle* h, double* r /*n*2*/, double* J_or_null);
/* per IMU interval: r 9, J 9 x 33 = (pose2 6 | pose1 6 | v2 3 | v1 3 | g 2 | b 6 | sf 6 | ts 1) */
int vcgpu_eval_imu(vcgpu_handle* h, double* r /*(n_frames-1)*9*/, double* J_or_null);
/* block normal equations at the current state: B nf*fd*fd, U nf*fd*fd (U[f] = H[f-1,f]),
 * E nf*fd*G, gf nf*fd, C G*G, gc G; any may be NULL */
int vcgpu_normal_equations(vcgpu_handle* h, double* B, double* U, double* E, double* gf, double* C,
                           double* gc, double* cost);
/* one damped solve (H*scale^2 + diag(D2)) x = -g*scale with the device arrow solver */
int vcgpu_solve_arrow(vcgpu_handle* h, const double* scale, const double* D2, double* x);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------
 * profile (bit mask): 1 = bracket every kernel stage of the multi-launch engine with CUDA events on the
 *          launching stream; 2 = two-pass path that materialises the Jacobian in HBM; 4 = multi-launch engine
 *          even where the persistent kernel applies; 8 = persistent kernel records per-phase device clocks
 *          (read back through vcgpu_get_stage_times)
 * flush_l2: overwrite a 256 MiB scratch buffer before every iteration and time iterations
 *           individually, so no iteration starts with its inputs resident in the 126 MB L2 */
enum {
  VCGPU_STAGE_DIAG = 0,         /* LM diagonal */
  VCGPU_STAGE_FRAME_SOLVE = 1,  /* per-frame Cholesky + Schur contribution (or IMU chain elimination) */
  VCGPU_STAGE_GLOBAL_SOLVE = 2, /* dense reduced system */
  VCGPU_STAGE_BACKSUB = 3,      /* back-substitution + x (+) delta */
  VCGPU_STAGE_EVAL_REPROJ = 4,  /* reprojection residual + Jacobian */
  VCGPU_STAGE_BUILD = 5,        /* per-frame block normal equations */
  VCGPU_STAGE_REDUCE = 6,       /* global block tree reduction, level 1 */
  VCGPU_STAGE_FINALIZE = 7,     /* level 2 + cost + gradient norms */
  VCGPU_STAGE_IMU_EVAL = 8,     /* IMU residual + Jacobian */
  VCGPU_STAGE_IMU_WEIGHTS = 9,  /* UpdateImuWeights */
  VCGPU_STAGE_GRID_SYNC = 10,   /* persistent vision kernel: the two grid barriers of an iteration */
  VCGPU_STAGE_EVAL_TASKS = 11,  /* persistent inertial kernels: IMU residual/Jacobian + reprojection evaluate/build (one task queue) */
  VCGPU_STAGE_IMU_ACCUM = 12,   /* persistent inertial kernels: J^T J of the IMU factors into the frame blocks */
  VCGPU_STAGE_COUNT = 16
};
/* Measured FP64 throughput of `device` in TFLOP/s (FMA = 2 flop): independent DFMA chains and independent
 * mma.sync.m8n8k4.f64 chains on every SM.  The roofline denominators for the FP64-bound kernels. */
int vcgpu_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops);
int vcgpu_set_profiling(vcgpu_handle* h, int profile, int flush_l2);
int vcgpu_get_stage_times(vcgpu_handle* h, double ms_total[VCGPU_STAGE_COUNT], int64_t launches[VCGPU_STAGE_COUNT]);
/* raw device phase clocks (ns, summed over the iterations since vcgpu_set_profiling) of the persistent inertial kernels:
 * chain_solve_kernel: elimination level l at [l], Schur reduce [10], dense solve [11], back-substitution level at [12 + k],
 * state update + step statistics [22], tail of the deferred UpdateImuWeights [23]; eval_mega_kernel: tasks [32], IMU
 * accumulate [33], reduce [34], decide [35], weights [36] (only when the update is not deferred into the solve launch) */
int vcgpu_get_phase_clocks(vcgpu_handle* h, uint64_t ns[64]);

/* ---- multi-GPU: one process per GPU of one node, frames sharded contiguously.  Two reductions per LM iteration (the
 * reduced system before the dense solve; global blocks + cost / step scalars before the decision), both INSIDE the
 * persistent kernels over NVLink peer memory: vcgpu_comm_init maps a 32 MiB totals buffer of every rank into every rank
 * (CUDA IPC; NCCL is used to pass the handles).  If the mapping is not possible the same reductions run as NCCL
 * all-reduces between kernel launches.  Every rank must issue the same sequence of calls. ---------------------------- */
#define VCGPU_UNIQUE_ID_BYTES 128
int vcgpu_comm_unique_id(uint8_t id[VCGPU_UNIQUE_ID_BYTES]);
int vcgpu_comm_init(vcgpu_handle* h, const uint8_t id[VCGPU_UNIQUE_ID_BYTES], int rank, int nranks);

#ifdef __cplusplus
}
#endif
#endif /* VCGPU_H_ */
