// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
// Flat C entry points (ctypes) over vo::Problem.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.
#include <chrono>
#include <cstring>

#include "problem.h"

using vo::Problem;

extern "C" {

void* vo_create() { return new Problem(); }
void vo_destroy(void* h) { delete static_cast<Problem*>(h); }

void vo_set_cameras(void* h, int n, const int* model, const double* intr10, const double* q_ck,
                    const double* p_ck) {
  Problem& P = *static_cast<Problem*>(h);
  P.n_cams = n;
  P.model.assign(model, model + n);
  P.intr.assign(intr10, intr10 + 10 * n);
  P.q_ck.assign(q_ck, q_ck + 4 * n);
  P.p_ck.assign(p_ck, p_ck + 3 * n);
}
void vo_set_frames(void* h, int n, const double* T_wp, const double* v_w, const double* time) {
  Problem& P = *static_cast<Problem*>(h);
  P.n_frames = n;
  P.T_wp.assign(T_wp, T_wp + 7 * n);
  P.v_w.assign(v_w, v_w + 3 * n);
  P.ftime.assign(time, time + n);
  P.ResetImuWeights();
}
void vo_set_observations(void* h, int64_t n, const int32_t* frame, const int32_t* cam, const double* p_w,
                         const double* p_c) {
  Problem& P = *static_cast<Problem*>(h);
  P.n_obs = n;
  P.obs_frame.assign(frame, frame + n);
  P.obs_cam.assign(cam, cam + n);
  P.p_w.assign(p_w, p_w + 3 * n);
  P.p_c.assign(p_c, p_c + 2 * n);
  P.obs_active.assign(n, 1);
}
void vo_set_imu(void* h, int n, const double* t, const double* w, const double* a, double sigma_g,
                double sigma_a) {
  Problem& P = *static_cast<Problem*>(h);
  P.imu = vo::InterpolationBuffer();
  for (int i = 0; i < n; ++i)
    P.imu.AddElement({{w[3 * i], w[3 * i + 1], w[3 * i + 2]}, {a[3 * i], a[3 * i + 1], a[3 * i + 2]}, t[i]});
  P.sigma_g = sigma_g;
  P.sigma_a = sigma_a;
}
void vo_set_imu_params(void* h, const double* g, const double* b, const double* sf, double ts) {
  Problem& P = *static_cast<Problem*>(h);
  std::memcpy(P.g, g, sizeof P.g);
  std::memcpy(P.b, b, sizeof P.b);
  std::memcpy(P.sf, sf, sizeof P.sf);
  P.ts = ts;
}
void vo_set_flags(void* h, int inertial, int rotation_only, int bias_active, int scale_active,
                  int optimize_ts, int fix_intrinsics, int visual, double visual_mult, double imu_mult) {
  Problem& P = *static_cast<Problem*>(h);
  P.flags = {inertial, rotation_only, bias_active, scale_active, optimize_ts, fix_intrinsics, visual,
             visual_mult, imu_mult};
}
void vo_set_options(void* h, int max_iters, double function_tol, double gradient_tol, double param_tol,
                    double init_radius, int strategy, int jacobi_scaling, int num_threads,
                    int update_imu_weights) {
  Problem& P = *static_cast<Problem*>(h);
  P.opts.max_iters = max_iters;
  P.opts.function_tol = function_tol;
  P.opts.gradient_tol = gradient_tol;
  P.opts.param_tol = param_tol;
  P.opts.init_radius = init_radius;
  P.opts.strategy = strategy;
  P.opts.jacobi_scaling = jacobi_scaling;
  P.opts.num_threads = num_threads;
  P.opts.update_imu_weights = update_imu_weights;
}
int vo_frame_dim(void* h) { return static_cast<Problem*>(h)->FrameDim(); }
int vo_num_globals(void* h) { return static_cast<Problem*>(h)->NumGlobals(); }
int vo_num_residuals(void* h) { return static_cast<Problem*>(h)->NumResiduals(); }
void vo_global_mask(void* h, double* out) {
  std::vector<double> m;
  static_cast<Problem*>(h)->GlobalMask(&m);
  std::memcpy(out, m.data(), m.size() * sizeof(double));
}

// residual blocks. J stride per observation = 2*22 (row-major 2 x (12+K), zero padded to 22 cols)
void vo_eval_reproj(void* h, int64_t i0, int64_t n, double* r, double* J) {
  const Problem& P = *static_cast<Problem*>(h);
  for (int64_t i = 0; i < n; ++i) {
    double Jl[2 * 22];
    P.EvalReprojection(i0 + i, r + 2 * i, J ? Jl : nullptr);
    if (J) {
      const int NT = 12 + vo::NumIntrinsics(P.model[P.obs_cam[i0 + i]]);
      double* o = J + 44 * i;
      std::memset(o, 0, 44 * sizeof(double));
      for (int row = 0; row < 2; ++row)
        for (int k = 0; k < NT; ++k) o[row * 22 + k] = Jl[row * NT + k];
    }
  }
}
void vo_eval_imu(void* h, int k0, int n, double* r, double* J) {
  const Problem& P = *static_cast<Problem*>(h);
  for (int k = 0; k < n; ++k) P.EvalImu(k0 + k, r + 9 * k, J ? J + 297 * k : nullptr);
}

// cost (+ normal equations into caller buffers, any may be null)
double vo_evaluate(void* h, double* B, double* U, double* E, double* gf, double* C, double* gc) {
  const Problem& P = *static_cast<Problem*>(h);
  if (!B && !U && !E && !gf && !C && !gc) return P.Evaluate(nullptr);
  vo::NormalEq ne;
  const double cost = P.Evaluate(&ne);
  if (B) std::memcpy(B, ne.B.data(), ne.B.size() * sizeof(double));
  if (U) std::memcpy(U, ne.U.data(), ne.U.size() * sizeof(double));
  if (E) std::memcpy(E, ne.E.data(), ne.E.size() * sizeof(double));
  if (gf) std::memcpy(gf, ne.gf.data(), ne.gf.size() * sizeof(double));
  if (C) std::memcpy(C, ne.C.data(), ne.C.size() * sizeof(double));
  if (gc) std::memcpy(gc, ne.gc.data(), ne.gc.size() * sizeof(double));
  return cost;
}
double vo_evaluate_camera(void* h, int cam, double* residuals) {
  const Problem& P = *static_cast<Problem*>(h);
  std::vector<double> res;
  const double c = P.EvaluateCamera(cam, residuals ? &res : nullptr);
  if (residuals) std::memcpy(residuals, res.data(), res.size() * sizeof(double));
  return c;
}
int vo_remove_outliers(void* h, const double* rmse, double threshold) {
  Problem& P = *static_cast<Problem*>(h);
  return P.RemoveOutliers(std::vector<double>(rmse, rmse + P.n_cams), threshold);
}
void vo_get_obs_active(void* h, uint8_t* out) {
  Problem& P = *static_cast<Problem*>(h);
  std::memcpy(out, P.obs_active.data(), P.obs_active.size());
}
void vo_update_imu_weights(void* h) { static_cast<Problem*>(h)->UpdateImuWeights(); }
void vo_get_imu_weights(void* h, double* out) {
  Problem& P = *static_cast<Problem*>(h);
  std::memcpy(out, P.w_sqrt.data(), P.w_sqrt.size() * sizeof(double));
}
void vo_set_imu_weights(void* h, const double* in) {
  Problem& P = *static_cast<Problem*>(h);
  std::memcpy(P.w_sqrt.data(), in, P.w_sqrt.size() * sizeof(double));
}

// one linear solve with the arrow solver, for solver unit tests:
// (H*scale^2 + diag(D2)) x = -g*scale, x returned in scaled space
int vo_solve_arrow(void* h, const double* scale, const double* D2, double* x_out) {
  const Problem& P = *static_cast<Problem*>(h);
  vo::NormalEq ne;
  P.Evaluate(&ne);
  const size_t n = static_cast<size_t>(ne.nf) * ne.fd + ne.G;
  std::vector<double> x;
  const bool ok = vo::SolveArrow(ne, std::vector<double>(scale, scale + n), std::vector<double>(D2, D2 + n), &x);
  if (ok) std::memcpy(x_out, x.data(), n * sizeof(double));
  return ok ? 0 : -1;
}
void vo_plus(void* h, const double* delta) {
  Problem& P = *static_cast<Problem*>(h);
  const size_t n = static_cast<size_t>(P.n_frames) * P.FrameDim() + P.NumGlobals();
  P.Plus(std::vector<double>(delta, delta + n));
}

// summary_out: [iterations, successful, initial_cost, final_cost, termination, num_residuals, seconds]
// rows_out (optional, max_rows x 9): iteration,cost,cost_change,gmax,gnorm,step_norm,rho,radius,accepted
int vo_solve(void* h, double* summary_out, double* rows_out, int max_rows) {
  Problem& P = *static_cast<Problem*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  const vo::Summary s = P.Solve();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  summary_out[0] = s.iterations;
  summary_out[1] = s.successful_steps;
  summary_out[2] = s.initial_cost;
  summary_out[3] = s.final_cost;
  summary_out[4] = s.termination;
  summary_out[5] = s.num_residuals;
  summary_out[6] = secs;
  const int nr = static_cast<int>(s.rows.size());
  if (rows_out)
    for (int i = 0; i < nr && i < max_rows; ++i) {
      const vo::IterationRow& r = s.rows[i];
      double* o = rows_out + 9 * i;
      o[0] = r.iteration; o[1] = r.cost; o[2] = r.cost_change; o[3] = r.gradient_max_norm;
      o[4] = r.gradient_norm; o[5] = r.step_norm; o[6] = r.rho; o[7] = r.radius; o[8] = r.accepted;
    }
  return nr;
}

void vo_get_state(void* h, double* intr10, double* q_ck, double* p_ck, double* T_wp, double* v_w, double* g,
                  double* b, double* sf, double* ts) {
  const Problem& P = *static_cast<Problem*>(h);
  if (intr10) std::memcpy(intr10, P.intr.data(), P.intr.size() * sizeof(double));
  if (q_ck) std::memcpy(q_ck, P.q_ck.data(), P.q_ck.size() * sizeof(double));
  if (p_ck) std::memcpy(p_ck, P.p_ck.data(), P.p_ck.size() * sizeof(double));
  if (T_wp) std::memcpy(T_wp, P.T_wp.data(), P.T_wp.size() * sizeof(double));
  if (v_w) std::memcpy(v_w, P.v_w.data(), P.v_w.size() * sizeof(double));
  if (g) std::memcpy(g, P.g, sizeof P.g);
  if (b) std::memcpy(b, P.b, sizeof P.b);
  if (sf) std::memcpy(sf, P.sf, sizeof P.sf);
  if (ts) *ts = P.ts;
}

// stand-alone helpers for unit tests -------------------------------------------------------
void vo_se3_exp(const double* d, double* out7) { vo::se3_to(vo::se3_exp(d), out7); }
void vo_se3_log(const double* x7, double* out6) { vo::se3_log(vo::se3_from(x7), out6); }
void vo_se3_plus(const double* x7, const double* d6, double* out7) { vo::LocalParamSe3Plus(x7, d6, out7); }
void vo_so3_plus(const double* x4, const double* d3, double* out4) { vo::LocalParamSo3Plus(x4, d3, out4); }
void vo_project(int model, const double* ray, const double* params, double* pix) {
  switch (model) {
    case vo::kLinear: vo::LinearCam::Project(ray, params, pix); break;
    case vo::kFov: vo::FovCam::Project(ray, params, pix); break;
    case vo::kPoly2: vo::Poly2Cam::Project(ray, params, pix); break;
    case vo::kPoly3: vo::Poly3Cam::Project(ray, params, pix); break;
    case vo::kKb4: vo::Kb4Cam::Project(ray, params, pix); break;
  }
}
// IMU range extraction (scalar): returns count, writes up to max entries of (t, w3, a3)
int vo_imu_get_range(void* h, double t0, double t1, double ts, double* out, int max) {
  const Problem& P = *static_cast<Problem*>(h);
  std::vector<vo::ImuMeas<double>> m;
  P.imu.GetRange(t0, t1, ts, &m);
  const int n = static_cast<int>(m.size());
  for (int i = 0; i < n && i < max; ++i) {
    double* o = out + 7 * i;
    o[0] = m[i].time; o[1] = m[i].w.x; o[2] = m[i].w.y; o[3] = m[i].w.z;
    o[4] = m[i].a.x; o[5] = m[i].a.y; o[6] = m[i].a.z;
  }
  return n;
}
}
