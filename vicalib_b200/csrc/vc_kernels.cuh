// CUDA kernels of the calibration solve (sm_100a, FP64).
//
// One trust-region iteration, vision terms (the fused evaluate+build kernel is in vc_fused.cuh):
//   frame_solve_kernel    per-frame (damped, Jacobi-scaled) 6x6 Cholesky, X = (B+D)^-1 [E | g],
//                         Schur contribution E^T X; one warp per frame
//   sum_partials_kernel   fixed-order sum of the per-CTA Schur partials           (vc_chain.cuh)
//   global_solve_kernel   dense Cholesky of the reduced (globals) system
//   backsub_update_kernel back-substitution + x (+) delta into the trial state
//   fused_build_kernel    residuals + Jacobians + block normal equations at the trial point
//   reduce_finalize_kernel  global blocks, cost, gradient norms and — in the CTA that arrives last —
//                         decide_step: accept / reject, trust-region radius, termination tests, on the device
// Every kernel picks its double buffer through Ctl::cur, so the host only enqueues.  The single- and
// multi-GPU vision solves run the same phases inside one persistent kernel instead (vc_mega.cuh).
// The reference does all of this inside ceres::Solve (vicalibrator.h:956).
#pragma once
#include "vc_internal.h"
#include "vc_math.cuh"

namespace vc {

__device__ __forceinline__ int pick(const Ctl* c, int which) { return which ? 1 - c->cur : c->cur; }

// scalars produced by the evaluation / step kernels for decide_step (indices into d_scalars)
enum {
  kScCost = 0,      // robust cost of the evaluated point
  kScGmax = 1,      // gradient max norm
  kScGnorm2 = 2,    // gradient squared 2-norm
  kScDotG = 3,      // step . g        (scaled space)
  kScDotD = 4,      // step . D2 step  (scaled space)
  kScStep2 = 5,     // |x_trial - x|^2 (ambient)
  kScXnorm2 = 6,    // |x_trial|^2 (ambient)
  kScNotPD = 7,     // >0 if a Cholesky pivot failed
  kScCount = 16
};

// ---------------------------------------------------------------- reprojection evaluate (two-pass path + hooks)
struct EvalArgs {
  const double* state[2];
  const Ctl* ctl;
  int which;
  int64_t cam_off;         // offset of this camera's 17 state doubles
  const int32_t* frame;    // per observation (camera-local arrays from here on)
  const double *pw, *pc;   // AoS [n][3], [n][2]
  const double* mask;      // 6+K column mask (w_ck, p_ck, intr)
  double *r0, *r1;
  double* J;               // column c at J + c*n
  double* cost_part;       // one per block
  int n;
  int apply_loss;          // 1: SoftLOne corrector + robust cost; 0: raw residuals, cost 1/2|r|^2
  double mult;
};

template <int MODEL, bool JAC>
__global__ void __launch_bounds__(256) eval_reproj_kernel(EvalArgs a) {
  constexpr int K = Cam<MODEL>::K, NT = 12 + K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* state = a.state[pick(a.ctl, a.which)];
  const double* cam = state + a.cam_off;
  double cost = 0.0;
  if (i < a.n) {
    const double* T = state + 7 * static_cast<int64_t>(a.frame[i]);
    const Q4 q{T[0], T[1], T[2], T[3]};
    const V3 t{T[4], T[5], T[6]};
    const V3 pw{a.pw[3 * i], a.pw[3 * i + 1], a.pw[3 * i + 2]};
    const V3 pk = qrot(qconj(q), pw - t);  // T_wk^-1 * p_w
    const Q4 qc{cam[0], cam[1], cam[2], cam[3]};
    double R[9];
    qmat(qc, R);
    const V3 pc = mat_mul(R, pk) + V3{cam[4], cam[5], cam[6]};
    double z[2], dzp[6], dzi[2 * K];
    Cam<MODEL>::project(pc, cam + 7, z, JAC ? dzp : nullptr, JAC ? dzi : nullptr);
    double r0 = z[0] - a.pc[2 * i], r1 = z[1] - a.pc[2 * i + 1];
    const double s = r0 * r0 + r1 * r1;
    double sc = 1.0;
    if (a.apply_loss) {
      double rho0, rho1;
      soft_l_one(s, &rho0, &rho1);
      cost = 0.5 * rho0 * a.mult;
      sc = sqrt(rho1);  // Ceres corrector for rho'' <= 0: scale residual and Jacobian by sqrt(rho')
    } else {
      cost = 0.5 * s;
    }
    a.r0[i] = r0 * sc;
    a.r1[i] = r1 * sc;
    if (JAC) {
      const int64_t n = a.n;
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double* d = dzp + 3 * row;
        const double m0 = d[0] * R[0] + d[1] * R[3] + d[2] * R[6];
        const double m1 = d[0] * R[1] + d[1] * R[4] + d[2] * R[7];
        const double m2 = d[0] * R[2] + d[1] * R[5] + d[2] * R[8];
        const double w0 = m1 * pk.z - m2 * pk.y;
        const double w1 = m2 * pk.x - m0 * pk.z;
        const double w2 = m0 * pk.y - m1 * pk.x;
        double* Jc = a.J + static_cast<int64_t>(row * NT) * n + i;
        Jc[0 * n] = -m0 * sc;
        Jc[1 * n] = -m1 * sc;
        Jc[2 * n] = -m2 * sc;
        Jc[3 * n] = w0 * sc;
        Jc[4 * n] = w1 * sc;
        Jc[5 * n] = w2 * sc;
        Jc[6 * n] = -w0 * sc * a.mask[0];
        Jc[7 * n] = -w1 * sc * a.mask[1];
        Jc[8 * n] = -w2 * sc * a.mask[2];
        Jc[9 * n] = d[0] * sc * a.mask[3];
        Jc[10 * n] = d[1] * sc * a.mask[4];
        Jc[11 * n] = d[2] * sc * a.mask[5];
#pragma unroll
        for (int k = 0; k < K; ++k) Jc[(12 + k) * n] = dzi[row * K + k] * sc * a.mask[6 + k];
      }
    }
  }
  __shared__ double wsum[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost += __shfl_down_sync(0xffffffffu, cost, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cost;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += wsum[w];
    a.cost_part[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------- per-frame block build (two-pass path)
struct BuildArgs {
  DevProblem dp;
  const Ctl* ctl;
  int which;
  const int32_t *grp_start, *grp_count, *group_of;
  const double* r;   // [2][n_obs]
  const double* J;
  int64_t n_obs;
  Blocks out[2];
  double* Cg;
};

constexpr int kBuildThreads = 128;
constexpr int kBuildChunk = 128;  // observations staged per pass

template <int FD>
__global__ void __launch_bounds__(kBuildThreads) build_frames_kernel(BuildArgs a) {
  extern __shared__ double smem[];
  double* tile = smem;                              // [2*kBuildChunk][W]
  double* smB = smem + 2 * kBuildChunk * kMaxW;     // [FD*FD]
  double* smg = smB + FD * FD;                      // [FD]
  const Blocks& out = a.out[pick(a.ctl, a.which)];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int G = a.dp.G, nf = a.dp.n_frames;
  for (int k = tid; k < FD * FD + FD; k += kBuildThreads) smB[k] = 0.0;
  double* Ef = out.E + static_cast<int64_t>(f) * FD * G;
  for (int k = tid; k < FD * G; k += kBuildThreads) Ef[k] = 0.0;
  int ei[2], ej[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int e = tid + s * kBuildThreads;
    int i = static_cast<int>((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    while (i * (i + 1) / 2 > e) --i;
    ei[s] = i;
    ej[s] = e - i * (i + 1) / 2;
  }
  __syncthreads();
  for (int c = 0; c < a.dp.n_cams; ++c) {
    const int g = a.group_of[c * nf + f];
    if (g < 0) continue;
    const CamInfo& ci = a.dp.cams[c];
    const int NT = 12 + ci.K, W = NT + 1, ntri = W * (W + 1) / 2;
    const int start = a.grp_start[g], cnt = a.grp_count[g];
    const int64_t nc = ci.n_obs;
    const double* Jc = a.J + ci.joff + (start - ci.obs_start);
    double acc[2] = {0.0, 0.0};
    for (int ch = 0; ch < cnt; ch += kBuildChunk) {
      const int m = min(kBuildChunk, cnt - ch);
      for (int col = 0; col < 2 * NT; ++col) {
        const int row = col / NT, k = col - row * NT;
        const double* src = Jc + col * nc + ch;
        for (int li = tid; li < m; li += kBuildThreads) tile[(2 * li + row) * W + k] = src[li];
      }
      for (int li = tid; li < m; li += kBuildThreads) {
        tile[(2 * li) * W + NT] = a.r[start + ch + li];
        tile[(2 * li + 1) * W + NT] = a.r[a.n_obs + start + ch + li];
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (tid + s * kBuildThreads < ntri) {
          const int i = ei[s], j = ej[s];
          double sum = 0.0;
          for (int row = 0; row < 2 * m; ++row) sum += tile[row * W + i] * tile[row * W + j];
          acc[s] += sum;
        }
      }
      __syncthreads();
    }
    double* Cgg = a.Cg + static_cast<int64_t>(g) * kCgStride;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (tid + s * kBuildThreads >= ntri) continue;
      const int i = ei[s], j = ej[s];
      const double v = acc[s] * a.dp.visual_mult;
      if (i < 6) {
        smB[i * FD + j] += v;
        if (i != j) smB[j * FD + i] += v;
      } else if (i < W - 1) {
        if (j < 6) Ef[j * G + ci.goff + (i - 6)] = v;
        else Cgg[(i - 6) * (i - 5) / 2 + (j - 6)] = v;
      } else {
        if (j < 6) smg[j] += v;
        else if (j < W - 1) Cgg[105 + (j - 6)] = v;
      }
    }
  }
  __syncthreads();
  double* Bf = out.B + static_cast<int64_t>(f) * FD * FD;
  for (int k = tid; k < FD * FD; k += kBuildThreads) Bf[k] = smB[k];
  for (int k = tid; k < FD; k += kBuildThreads) out.gf[static_cast<int64_t>(f) * FD + k] = smg[k];
}

// ---------------------------------------------------------------- accept / reject on the device
// mode 0: initial point (iteration 0): adopt cost / gradient norms
// mode 1: one iteration of Ceres' TrustRegionMinimizer loop with the LM strategy's radius rules
__device__ inline void decide_step(Ctl* c, double* sc, int mode) {
  if (c->done) return;
  if (mode == 0) {
    c->cost = sc[kScCost];
    c->initial_cost = sc[kScCost];
    c->gmax = sc[kScGmax];
    c->gnorm = sqrt(sc[kScGnorm2]);
    c->last_accepted = 1;
    if (!c->fixed && c->gmax <= c->gradient_tol) c->done = 1 + VCGPU_TERM_GRADIENT_TOL;
    sc[kScNotPD] = 0.0;
    return;
  }
  const int iter = ++c->iter;
  const double cand = sc[kScCost];
  // model_cost_change = -step.g - step.H.step/2, with (H + D2) step = -g
  const double model_change = -0.5 * sc[kScDotG] + 0.5 * sc[kScDotD];
  const bool invalid = sc[kScNotPD] > 0.0 || !(model_change > 0.0) || !isfinite(cand);
  sc[kScNotPD] = 0.0;
  c->last_accepted = 0;
  c->last_cand_cost = cand;
  c->last_rho = 0.0;
  c->last_cost_change = 0.0;
  c->last_step_norm = 0.0;
  if (invalid) {  // StepIsInvalid == StepRejected for the LM strategy
    c->radius /= c->decrease_factor;
    c->decrease_factor *= 2.0;
  } else {
    const double step_norm = sqrt(sc[kScStep2]);
    const double cost_change = c->cost - cand;
    c->last_step_norm = step_norm;
    c->last_cost_change = cost_change;
    if (!c->fixed && step_norm <= c->param_tol * (c->x_norm + c->param_tol)) {
      c->done = 1 + VCGPU_TERM_PARAM_TOL;
      return;
    }
    if (!c->fixed && fabs(cost_change) <= c->function_tol * c->cost) {
      c->done = 1 + VCGPU_TERM_FUNCTION_TOL;
      return;
    }
    const double rho = cost_change / model_change;
    c->last_rho = rho;
    if (rho > 1e-3) {
      c->cur = 1 - c->cur;  // the candidate's blocks were built speculatively: accept = flip
      c->cost = cand;
      c->gmax = sc[kScGmax];
      c->gnorm = sqrt(sc[kScGnorm2]);
      c->x_norm = sqrt(sc[kScXnorm2]);
      ++c->successful;
      const double t = 2.0 * rho - 1.0;
      c->radius = fmin(1e16, c->radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      c->decrease_factor = 2.0;
      c->last_accepted = 1;
      if (!c->fixed) {
        if (c->gmax <= c->gradient_tol) c->done = 1 + VCGPU_TERM_GRADIENT_TOL;
        else if (c->gnorm > 0.0 && c->gnorm < 1e-9) c->done = 1 + VCGPU_TERM_CALLBACK;  // vicalibrator.h:713-717
      }
    } else {
      c->radius /= c->decrease_factor;
      c->decrease_factor *= 2.0;
    }
  }
  if (c->radius < 1e-32) {
    if (c->fixed) { c->radius = 1e4; c->decrease_factor = 2.0; }
    else if (!c->done) c->done = 1 + VCGPU_TERM_RADIUS;
  }
  if (!c->done && iter >= c->max_iters) c->done = 1 + VCGPU_TERM_NO_CONVERGENCE;
}

// ---------------------------------------------------------------- reduce + finalize + decide in one launch
// Level 1 (every CTA): a slice of the per-group global blocks, cost partials, gradient norms and step
// reductions.  Level 2 (the CTA that arrives last, found with a self-resetting ticket): fixed-order sum
// of the level-1 partials, C / gc / scalars, then the accept-reject decision.  One launch instead of
// three; deterministic because level 2 always sums in block order.
struct RedFinArgs {
  DevProblem dp;        // cameras to reduce (n_cams = 0 when visual terms are off)
  Ctl* ctl;
  int which;
  int decide_mode;      // -1: none (multi-GPU: decide after the all-reduce), 0: initial point, 1: iteration
  int multi;            // 1: frame-sharded run: scalars cover this rank's frames only (gc joins after the all-reduce)
  int level1_only;      // 1: stop after the per-CTA partials (large G: level 2 runs as its own parallel launches)
  int64_t gf_skip_below, gf_skip_from;  // sharded inertial run: separator frames' gradients are normed after the all-reduce
  const double* Cg;
  const double* imuCg;  // [ni][kImuCgStride] or null
  int ni, imu_goff, imu_stride;
  double* Cpart;        // [gridDim][G*G+G]
  double* red_part;     // [gridDim][8]
  const double* cost_part;
  int n_cost_part;
  const double* imu_cost_part;
  int n_imu_cost_part;
  const double* step_part;  // [n_step_part][4] or null
  int n_step_part;
  int n_frames_fd;      // n_frames * fd
  Blocks out[2];
  double* scalars;
  unsigned* counter;
};
// Level 1 of the reduction for slice `bid` of `nb` (256 threads): the slice's share of the per-group global blocks
// -> Cpart[bid], of the cost / gradient-norm / step partials -> red_part[bid].  acc: NS doubles of shared memory.
__device__ __forceinline__ void reduce_level1(const RedFinArgs& a, const DevProblem& dp, const Blocks& out, double* acc, double (*shr)[8],
                                              int bid, int nb) {
  const int G = dp.G, tid = threadIdx.x, NS = G * G + G, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < NS; k += 256) acc[k] = 0.0;
  __syncthreads();
  for (int c = 0; c < dp.n_cams; ++c) {
    const CamInfo& ci = dp.cams[c];
    const int NG = 6 + ci.K, nsym = NG * (NG + 1) / 2;
    const int lo = static_cast<int>(static_cast<int64_t>(ci.n_groups) * bid / nb);
    const int hi = static_cast<int>(static_cast<int64_t>(ci.n_groups) * (bid + 1) / nb);
    for (int e = tid; e < nsym + NG; e += 256) {
      int src, dst, dst2 = -1;
      if (e < nsym) {
        int i = static_cast<int>((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        while (i * (i + 1) / 2 > e) --i;
        const int j = e - i * (i + 1) / 2;
        src = e;
        dst = (ci.goff + i) * G + ci.goff + j;
        if (i != j) dst2 = (ci.goff + j) * G + ci.goff + i;
      } else {
        src = 105 + (e - nsym);
        dst = G * G + ci.goff + (e - nsym);
      }
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const double* p = a.Cg + static_cast<int64_t>(ci.group_start + lo) * kCgStride + src;
      int g = lo;
      for (; g + 3 < hi; g += 4, p += 4 * kCgStride) {
        s0 += p[0]; s1 += p[kCgStride]; s2 += p[2 * kCgStride]; s3 += p[3 * kCgStride];
      }
      for (; g < hi; ++g, p += kCgStride) s0 += *p;
      const double s = (s0 + s1) + (s2 + s3);
      acc[dst] += s;
      if (dst2 >= 0) acc[dst2] += s;
    }
    __syncthreads();
  }
  if (a.imuCg) {
    const int io = a.imu_goff;
    const int lo = static_cast<int>(static_cast<int64_t>(a.ni) * bid / nb);
    const int hi = static_cast<int>(static_cast<int64_t>(a.ni) * (bid + 1) / nb);
    for (int q = tid; q < 135; q += 256) {
      double s0 = 0.0, s1 = 0.0;
      const double* p = a.imuCg + static_cast<int64_t>(lo) * a.imu_stride + q;
      int k = lo;
      for (; k + 1 < hi; k += 2, p += 2 * a.imu_stride) { s0 += p[0]; s1 += p[a.imu_stride]; }
      if (k < hi) s0 += *p;
      const double s = s0 + s1;
      if (q < 120) {
        int i = static_cast<int>((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        while (i * (i + 1) / 2 > q) --i;
        const int j = q - i * (i + 1) / 2;
        acc[(io + i) * G + io + j] = s;
        if (i != j) acc[(io + j) * G + io + i] = s;
      } else {
        acc[G * G + io + (q - 120)] = s;
      }
    }
    __syncthreads();
  }
  double* cp = a.Cpart + static_cast<int64_t>(bid) * NS;
  for (int k = tid; k < NS; k += 256) cp[k] = acc[k];
  // slice reductions: cost, |gf|^2, step sums (4), |gf|_inf
  double v[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  {
    const int lo = static_cast<int>(static_cast<int64_t>(a.n_cost_part) * bid / nb), hi = static_cast<int>(static_cast<int64_t>(a.n_cost_part) * (bid + 1) / nb);
    for (int k = lo + tid; k < hi; k += 256) v[0] += a.cost_part[k];
    const int li = static_cast<int>(static_cast<int64_t>(a.n_imu_cost_part) * bid / nb), hi2 = static_cast<int>(static_cast<int64_t>(a.n_imu_cost_part) * (bid + 1) / nb);
    for (int k = li + tid; k < hi2; k += 256) v[0] += a.imu_cost_part[k];
    const int64_t lg = static_cast<int64_t>(a.n_frames_fd) * bid / nb, hg = static_cast<int64_t>(a.n_frames_fd) * (bid + 1) / nb;
    for (int64_t k = lg + tid; k < hg; k += 256) {
      if (k < a.gf_skip_below || k >= a.gf_skip_from) continue;
      const double g = out.gf[k];
      v[6] = fmax(v[6], fabs(g));
      v[1] += g * g;
    }
    if (a.step_part) {
      const int ls = static_cast<int>(static_cast<int64_t>(a.n_step_part) * bid / nb), hs = static_cast<int>(static_cast<int64_t>(a.n_step_part) * (bid + 1) / nb);
      for (int k = ls + tid; k < hs; k += 256)
        for (int q = 0; q < 4; ++q) v[2 + q] += a.step_part[4 * k + q];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 6; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
    v[6] = fmax(v[6], __shfl_down_sync(0xffffffffu, v[6], o));
  }
  if (lane == 0)
    for (int q = 0; q < 7; ++q) shr[warp][q] = v[q];
  __syncthreads();
  if (tid == 0) {
    double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < 8; ++w) {
      for (int q = 0; q < 6; ++q) t[q] += shr[w][q];
      t[6] = fmax(t[6], shr[w][6]);
    }
    for (int q = 0; q < 7; ++q) a.red_part[8 * bid + q] = t[q];
  }
}

__global__ void __launch_bounds__(256) reduce_finalize_kernel(RedFinArgs a) {
  extern __shared__ double acc[];
  __shared__ double shr[8][8];
  __shared__ double sh4[4][64];
  __shared__ int is_last;
  if (a.ctl->done) return;
  const Blocks& out = a.out[pick(a.ctl, a.which)];
  const int G = a.dp.G, tid = threadIdx.x, NS = G * G + G, lane = tid & 31, warp = tid >> 5;
  const int nb = gridDim.x, bid = blockIdx.x;
  reduce_level1(a, a.dp, out, acc, shr, bid, nb);
  if (tid == 0) {
    if (a.level1_only) { is_last = 0; }
    __threadfence();
    if (!a.level1_only) {
      const unsigned ticket = atomicInc(a.counter, static_cast<unsigned>(nb - 1));
      is_last = ticket == static_cast<unsigned>(nb - 1);
    }
  }
  __threadfence();
  __syncthreads();
  if (!is_last) return;
  // ---- level 2 (one CTA)
  __threadfence();
  double gm = 0.0, g2 = 0.0;
  const int el = tid & 63, slice = tid >> 6;
  for (int e0 = 0; e0 < NS; e0 += 64) {
    const int e = e0 + el;
    double s0 = 0.0, s1 = 0.0;
    if (e < NS) {
      int b = slice;
      for (; b + 4 < nb; b += 8) {
        s0 += __ldcg(a.Cpart + static_cast<int64_t>(b) * NS + e);
        s1 += __ldcg(a.Cpart + static_cast<int64_t>(b + 4) * NS + e);
      }
      if (b < nb) s0 += __ldcg(a.Cpart + static_cast<int64_t>(b) * NS + e);
    }
    sh4[slice][el] = s0 + s1;
    __syncthreads();
    if (slice == 0 && e < NS) {
      const double s = (sh4[0][el] + sh4[1][el]) + (sh4[2][el] + sh4[3][el]);
      if (e < G * G) {
        out.C[e] = s;
      } else {
        out.gc[e - G * G] = s;
        gm = fmax(gm, fabs(s));
        g2 += s * s;
      }
    }
    __syncthreads();
  }
  double w[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (tid < nb)
    for (int q = 0; q < 7; ++q) w[q] = __ldcg(a.red_part + 8 * tid + q);
  if (!a.multi) {
    w[1] += g2;
    w[6] = fmax(w[6], gm);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 6; ++q) w[q] += __shfl_down_sync(0xffffffffu, w[q], o);
    w[6] = fmax(w[6], __shfl_down_sync(0xffffffffu, w[6], o));
  }
  if (lane == 0)
    for (int q = 0; q < 7; ++q) shr[warp][q] = w[q];
  __syncthreads();
  if (tid == 0) {
    double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int ww = 0; ww < 8; ++ww) {
      for (int q = 0; q < 6; ++q) t[q] += shr[ww][q];
      t[6] = fmax(t[6], shr[ww][6]);
    }
    *out.cost = t[0];
    a.scalars[kScCost] = t[0];
    a.scalars[kScGmax] = t[6];
    a.scalars[kScGnorm2] = t[1];
    if (a.step_part) {
      a.scalars[kScDotG] = t[2];
      a.scalars[kScDotD] = t[3];
      a.scalars[kScStep2] = t[4];
      a.scalars[kScXnorm2] = t[5];
    }
    if (a.decide_mode >= 0) decide_step(a.ctl, a.scalars, a.decide_mode);
  }
}

// level 2 for large global blocks: C | gc were summed by sum_partials_sel_kernel; this one-CTA launch
// finishes the scalars (fixed order) and decides
__global__ void __launch_bounds__(256) finalize_small_kernel(RedFinArgs a, int nb) {
  __shared__ double shr[8][8];
  if (a.ctl->done) return;
  const Blocks& out = a.out[pick(a.ctl, a.which)];
  const int G = a.dp.G, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double w[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (tid < nb)
    for (int q = 0; q < 7; ++q) w[q] = a.red_part[8 * tid + q];
  if (!a.multi)
    for (int k = tid; k < G; k += 256) {
      const double v = out.gc[k];
      w[1] += v * v;
      w[6] = fmax(w[6], fabs(v));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 6; ++q) w[q] += __shfl_down_sync(0xffffffffu, w[q], o);
    w[6] = fmax(w[6], __shfl_down_sync(0xffffffffu, w[6], o));
  }
  if (lane == 0)
    for (int q = 0; q < 7; ++q) shr[warp][q] = w[q];
  __syncthreads();
  if (tid == 0) {
    double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int ww = 0; ww < 8; ++ww) {
      for (int q = 0; q < 6; ++q) t[q] += shr[ww][q];
      t[6] = fmax(t[6], shr[ww][6]);
    }
    *out.cost = t[0];
    a.scalars[kScCost] = t[0];
    a.scalars[kScGmax] = t[6];
    a.scalars[kScGnorm2] = t[1];
    if (a.step_part) {
      a.scalars[kScDotG] = t[2];
      a.scalars[kScDotD] = t[3];
      a.scalars[kScStep2] = t[4];
      a.scalars[kScXnorm2] = t[5];
    }
    if (a.decide_mode >= 0) decide_step(a.ctl, a.scalars, a.decide_mode);
  }
}

// ---------------------------------------------------------------- multi-GPU: pack / unpack around the all-reduce
// buffer: [C G*G | gc G | cost, |gf|^2, dotG, dotD, step2, xnorm2 | per-rank |gf|_inf slots]
// sharded inertial runs append [diag(B) of separator frames P*9 | g of separator frames P*9]
__global__ void mg_pack_kernel(int G, const Ctl* ctl, int which, Blocks b0, Blocks b1, const double* scalars, int rank,
                               int nranks, double* buf, int sep_fd, int n_frames, int ghost) {
  if (ctl->done) return;
  const Blocks& b = pick(ctl, which) ? b1 : b0;
  const int NS = G * G + G, tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = tid; k < NS; k += gridDim.x * blockDim.x) buf[k] = k < G * G ? b.C[k] : b.gc[k - G * G];
  if (sep_fd > 0) {
    double* sd = buf + NS + 6 + nranks;
    double* sg = sd + nranks * sep_fd;
    for (int k = tid; k < nranks * sep_fd; k += gridDim.x * blockDim.x) {
      const int slot = k / sep_fd, r = k - slot * sep_fd;
      const int f = slot == rank ? 0 : (ghost && slot == rank + 1 ? n_frames - 1 : -1);
      sd[k] = f >= 0 ? b.B[(static_cast<int64_t>(f) * sep_fd + r) * sep_fd + r] : 0.0;
      sg[k] = f >= 0 ? b.gf[static_cast<int64_t>(f) * sep_fd + r] : 0.0;
    }
  }
  if (tid == 0) {
    buf[NS + 0] = scalars[kScCost];
    buf[NS + 1] = scalars[kScGnorm2];
    buf[NS + 2] = scalars[kScDotG];
    buf[NS + 3] = scalars[kScDotD];
    buf[NS + 4] = scalars[kScStep2];
    buf[NS + 5] = scalars[kScXnorm2];
    for (int r = 0; r < nranks; ++r) buf[NS + 6 + r] = r == rank ? scalars[kScGmax] : 0.0;
  }
}
__global__ void mg_unpack_decide_kernel(int G, Ctl* ctl, int which, Blocks b0, Blocks b1, double* scalars, int nranks,
                                        const double* buf, int decide_mode, int sep_fd, double* sep_out) {
  if (ctl->done) return;
  const Blocks& b = pick(ctl, which) ? b1 : b0;
  const int NS = G * G + G, tid = threadIdx.x;
  __shared__ double shm[256], shs[256];
  double gm = 0.0, g2 = 0.0;
  for (int k = tid; k < NS; k += 256) {
    const double v = buf[k];
    if (k < G * G) {
      b.C[k] = v;
    } else {
      b.gc[k - G * G] = v;
      gm = fmax(gm, fabs(v));
      g2 += v * v;
    }
  }
  if (sep_fd > 0) {  // separator frames: keep the summed diagonal / gradient, norm the gradient once
    const double* sd = buf + NS + 6 + nranks;
    for (int k = tid; k < 2 * nranks * sep_fd; k += 256) {
      sep_out[k] = sd[k];
      if (k >= nranks * sep_fd) {
        gm = fmax(gm, fabs(sd[k]));
        g2 += sd[k] * sd[k];
      }
    }
  }
  shm[tid] = gm;
  shs[tid] = g2;
  __syncthreads();
  if (tid == 0) {
    gm = 0.0; g2 = 0.0;
    for (int k = 0; k < 256; ++k) { gm = fmax(gm, shm[k]); g2 += shs[k]; }
    for (int r = 0; r < nranks; ++r) gm = fmax(gm, buf[NS + 6 + r]);
    *b.cost = buf[NS];
    scalars[kScCost] = buf[NS];
    scalars[kScGmax] = gm;
    scalars[kScGnorm2] = buf[NS + 1] + g2;
    scalars[kScDotG] = buf[NS + 2];
    scalars[kScDotD] = buf[NS + 3];
    scalars[kScStep2] = buf[NS + 4];
    scalars[kScXnorm2] = buf[NS + 5];
    if (decide_mode >= 0) decide_step(ctl, scalars, decide_mode);
  }
}

// ---------------------------------------------------------------- Jacobi scaling
// scale = 1/(1+sqrt(diag(J'J)))   (Ceres TrustRegionMinimizer, jacobi_scaling, computed once)
__global__ void jacobi_scale_kernel(DevProblem dp, Blocks b0, Blocks b1, const Ctl* ctl, double* out, const double* sepdiag) {
  const Blocks& b = ctl->cur ? b1 : b0;
  const int64_t n = static_cast<int64_t>(dp.n_frames) * dp.fd + dp.G;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t nfp = static_cast<int64_t>(dp.n_frames) * dp.fd;
  double d;
  if (i < nfp) {
    const int64_t f = i / dp.fd;
    const int k = static_cast<int>(i - f * dp.fd);
    d = b.B[(f * dp.fd + k) * dp.fd + k];
    if (sepdiag) {  // separator frames: the diagonal summed over the two ranks that hold a piece of it
      if (f == 0) d = sepdiag[dp.rank * dp.fd + k];
      else if (dp.ghost && f == dp.n_frames - 1) d = sepdiag[(dp.rank + 1) * dp.fd + k];
    }
  } else {
    const int k = static_cast<int>(i - nfp);
    d = b.C[k * dp.G + k];
  }
  out[i] = 1.0 / (1.0 + sqrt(d));
}

// ---------------------------------------------------------------- per-frame solve (no coupling)
struct SolveArgs {
  DevProblem dp;
  Blocks b[2];
  const Ctl* ctl;
  const double* scale;
  const double* D2x;  // explicit damping (inspection hook) or null: lm_damp(diag, scale, 1/radius)
  double* X;          // [nf][fd][G+1]  = (B + D)^-1 [E | g]   (scaled space)
  double* Spart;      // [gridDim][G*G+G]
  double* scalars;
};
constexpr int kSolveThreads = 128;
constexpr int kSolveWarps = kSolveThreads / 32;

// One warp per frame: every lane factors the (damped, scaled) FD x FD block in registers and the
// lanes split the right-hand-side columns [E | g]; the CTA then accumulates E^T X for its four
// frames with a fixed thread->entry ownership (deterministic).
template <int FD>
__global__ void __launch_bounds__(kSolveThreads) frame_solve_kernel(SolveArgs a) {
  extern __shared__ double sm[];
  if (a.ctl->done) return;
  const Blocks& b = a.b[a.ctl->cur];
  const double rinv = 1.0 / a.ctl->radius;
  const int G = a.dp.G, M = G + 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NS = G * G + G;
  double* Sacc = sm;                                   // [G*G+G]
  double* Esw = Sacc + NS + warp * (2 * FD * M);       // per warp: Es [FD][M] (col G = scaled g)
  double* Xw = Esw + FD * M;                           // per warp: X  [FD][M]
  const double* sc = a.scale + static_cast<int64_t>(a.dp.n_frames) * FD;
  for (int k = tid; k < NS; k += kSolveThreads) Sacc[k] = 0.0;
  for (int base = blockIdx.x * kSolveWarps; base < a.dp.n_frames; base += gridDim.x * kSolveWarps) {
    const int f = base + warp;
    __syncthreads();
    if (f < a.dp.n_frames) {
      const double* sf = a.scale + static_cast<int64_t>(f) * FD;
      const double* Bf = b.B + static_cast<int64_t>(f) * FD * FD;
      const double* Ef = b.E + static_cast<int64_t>(f) * FD * G;
      double s[FD], L[FD][FD], iL[FD];
#pragma unroll
      for (int i = 0; i < FD; ++i) s[i] = sf[i];
#pragma unroll
      for (int i = 0; i < FD; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          const double bij = Bf[i * FD + j];
          double v = bij * s[i] * s[j];
          if (i == j) v += a.D2x ? a.D2x[static_cast<int64_t>(f) * FD + i] : lm_damp(bij, s[i], rinv);
          L[i][j] = v;
        }
      bool ok = true;
#pragma unroll
      for (int j = 0; j < FD; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        d = sqrt(d);
        L[j][j] = d;
        const double inv = 1.0 / d;
        iL[j] = inv;
#pragma unroll
        for (int i = j + 1; i < FD; ++i) {
          double t = L[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
          L[i][j] = t * inv;
        }
      }
      if (!ok && lane == 0) a.scalars[kScNotPD] = 1.0;
      double* Xf = a.X + static_cast<int64_t>(f) * FD * M;
      for (int c = lane; c < M; c += 32) {
        double x[FD];
        const double scc = c < G ? sc[c] : 1.0;
#pragma unroll
        for (int i = 0; i < FD; ++i) {
          const double r = (c < G ? Ef[i * G + c] : b.gf[static_cast<int64_t>(f) * FD + i]) * s[i] * scc;
          Esw[i * M + c] = r;
          x[i] = r;
        }
#pragma unroll
        for (int i = 0; i < FD; ++i) {
          double t = x[i];
#pragma unroll
          for (int k = 0; k < i; ++k) t -= L[i][k] * x[k];
          x[i] = t * iL[i];
        }
#pragma unroll
        for (int i = FD - 1; i >= 0; --i) {
          double t = x[i];
#pragma unroll
          for (int k = i + 1; k < FD; ++k) t -= L[k][i] * x[k];
          x[i] = t * iL[i];
        }
#pragma unroll
        for (int i = 0; i < FD; ++i) {
          Xw[i * M + c] = x[i];
          Xf[i * M + c] = x[i];
        }
      }
    }
    __syncthreads();
    const int nact = min(kSolveWarps, a.dp.n_frames - base);
    for (int e = tid; e < NS; e += kSolveThreads) {
      const int ra = e < G * G ? e / G : e - G * G;
      const int cb = e < G * G ? e - ra * G : G;
      double sum = 0.0;
      for (int w = 0; w < nact; ++w) {
        const double* Ew = sm + NS + w * (2 * FD * M);
        const double* Xv = Ew + FD * M;
#pragma unroll
        for (int k = 0; k < FD; ++k) sum += Ew[k * M + ra] * Xv[k * M + cb];
      }
      Sacc[e] += sum;
    }
  }
  __syncthreads();
  double* out = a.Spart + static_cast<int64_t>(blockIdx.x) * NS;
  for (int k = tid; k < NS; k += kSolveThreads) out[k] = Sacc[k];
}

// ---------------------------------------------------------------- reduced dense solve
struct GlobalSolveArgs {
  DevProblem dp;
  Blocks b[2];
  const Ctl* ctl;
  const double* scale;
  const double* D2x;
  const double* Ssum;  // [G*G+G] summed Schur partials (n_part == 0), or
  const double* Spart; // [n_part][G*G+G] per-CTA partials summed here in fixed order
  int n_part;
  double* delta;       // scaled step; globals written at delta[nf*fd ...]
  double* scalars;
};
__global__ void __launch_bounds__(256) global_solve_kernel(GlobalSolveArgs a) {
  extern __shared__ double sm[];
  __shared__ double sh4[4][64];
  if (a.ctl->done) return;
  const Blocks& b = a.b[a.ctl->cur];
  const double rinv = 1.0 / a.ctl->radius;
  const int G = a.dp.G, tid = threadIdx.x, NS = G * G + G;
  double* S = sm;           // [G*G]
  double* rhs = sm + G * G; // [G]
  __shared__ int bad;
  if (tid == 0) bad = 0;
  const int64_t nfp = static_cast<int64_t>(a.dp.n_frames) * a.dp.fd;
  const double* sc = a.scale + nfp;
  const int el = tid & 63, slice = tid >> 6;
  for (int e0 = 0; e0 < NS; e0 += 64) {
    const int e = e0 + el;
    double p = 0.0;
    if (a.n_part > 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      if (e < NS) {
        int bb = slice;
        for (; bb + 12 < a.n_part; bb += 16) {
          s0 += a.Spart[static_cast<int64_t>(bb) * NS + e];
          s1 += a.Spart[static_cast<int64_t>(bb + 4) * NS + e];
          s2 += a.Spart[static_cast<int64_t>(bb + 8) * NS + e];
          s3 += a.Spart[static_cast<int64_t>(bb + 12) * NS + e];
        }
        for (; bb < a.n_part; bb += 4) s0 += a.Spart[static_cast<int64_t>(bb) * NS + e];
      }
      sh4[slice][el] = (s0 + s1) + (s2 + s3);
      __syncthreads();
      p = (sh4[0][el] + sh4[1][el]) + (sh4[2][el] + sh4[3][el]);
      __syncthreads();
    } else if (e < NS) {
      p = a.Ssum[e];
    }
    if (slice == 0 && e < NS) {
      if (e < G * G) {
        const int r = e / G, c = e - r * G;
        double v = b.C[e] * sc[r] * sc[c] - p;
        if (r == c) v += a.D2x ? a.D2x[nfp + r] : lm_damp(b.C[e], sc[r], rinv);
        S[e] = v;
      } else {
        const int r = e - G * G;
        rhs[r] = -b.gc[r] * sc[r] + p;
      }
    }
  }
  __syncthreads();
  for (int j = 0; j < G; ++j) {
    if (tid == 0) {
      double d = S[j * G + j];
      if (!(d > 0.0)) { bad = 1; d = 1.0; }
      S[j * G + j] = sqrt(d);
    }
    __syncthreads();
    const double dj = S[j * G + j];
    for (int i = j + 1 + tid; i < G; i += 256) S[i * G + j] /= dj;
    __syncthreads();
    const int rem = G - j - 1;
    for (int e = tid; e < rem * rem; e += 256) {
      const int i = j + 1 + e / rem, k = j + 1 + e % rem;
      if (k <= i) S[i * G + k] -= S[i * G + j] * S[k * G + j];
    }
    __syncthreads();
  }
  if (tid < 32) {  // warp-cooperative triangular solves (column oriented), L L^T x = rhs
    const int lane = tid;
    for (int i = 0; i < G; ++i) {
      const double xi = rhs[i] / S[i * G + i];
      __syncwarp();
      if (lane == 0) rhs[i] = xi;
      for (int k = i + 1 + lane; k < G; k += 32) rhs[k] -= S[k * G + i] * xi;
      __syncwarp();
    }
    for (int i = G - 1; i >= 0; --i) {
      const double xi = rhs[i] / S[i * G + i];
      __syncwarp();
      if (lane == 0) rhs[i] = xi;
      for (int k = lane; k < i; k += 32) rhs[k] -= S[i * G + k] * xi;
      __syncwarp();
    }
    if (bad && lane == 0) a.scalars[7] = 1.0;
  }
  __syncthreads();
  for (int i = tid; i < G; i += 256) a.delta[nfp + i] = bad ? 0.0 : rhs[i];
}

// ---------------------------------------------------------------- back-substitution + x (+) delta
struct UpdateArgs {
  DevProblem dp;
  Blocks b[2];
  const Ctl* ctl;
  const double* scale;
  const double* D2x;
  const double* X;    // null: the chain solver already wrote the frame steps into delta
  double* delta;
  double* state[2];
  double* step_part;  // [gridDim+1][4]
  const double* sepdiag;  // sharded inertial run: summed diagonal of the separator frames, else null
};
constexpr int kUpdateWarps = 4;

template <int FD>
__global__ void __launch_bounds__(32 * kUpdateWarps) backsub_update_kernel(UpdateArgs a) {
  if (a.ctl->done) return;
  const int cur = a.ctl->cur;
  const Blocks& b = a.b[cur];
  const double* x_cur = a.state[cur];
  double* x_new = a.state[1 - cur];
  const double rinv = 1.0 / a.ctl->radius;
  const int G = a.dp.G, M = G + 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nf = a.dp.n_frames;
  const int64_t nfp = static_cast<int64_t>(nf) * FD;
  const double* dc = a.delta + nfp;
  __shared__ double part[kUpdateWarps][4];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int f = blockIdx.x * kUpdateWarps + warp;
  if (f < nf) {
    double d[FD];
    if (a.X) {
      const double* Xf = a.X + static_cast<int64_t>(f) * FD * M;
#pragma unroll
      for (int r = 0; r < FD; ++r) {
        double s = 0.0;
        for (int c = lane; c < G; c += 32) s += Xf[r * M + c] * dc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        d[r] = -Xf[r * M + G] - s;
      }
    } else {
#pragma unroll
      for (int r = 0; r < FD; ++r) d[r] = a.delta[static_cast<int64_t>(f) * FD + r];
    }
    if (lane == 0) {
      double du[FD];
      const bool is_ghost = a.dp.ghost && f == nf - 1;
#pragma unroll
      for (int r = 0; r < FD; ++r) {
        const int64_t k = static_cast<int64_t>(f) * FD + r;
        const double sc = a.scale[k];
        double d2;
        if (a.D2x) d2 = a.D2x[k];
        else if (a.sepdiag && f == 0) d2 = lm_damp(a.sepdiag[a.dp.rank * FD + r], sc, rinv);
        else if (is_ghost) d2 = 0.0;  // damped (and counted) by its owner
        else d2 = lm_damp(b.B[k * FD + r], sc, rinv);
        a.delta[k] = d[r];
        acc[0] += d[r] * b.gf[k] * sc;
        acc[1] += d[r] * d[r] * d2;
        du[r] = d[r] * sc;
      }
      const double* x = x_cur + 7 * static_cast<int64_t>(f);
      double xo[7];
      se3_plus(x, du, xo);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        x_new[7 * static_cast<int64_t>(f) + k] = xo[k];
        if (!is_ghost) {
          acc[2] += (xo[k] - x[k]) * (xo[k] - x[k]);
          acc[3] += xo[k] * xo[k];
        }
      }
      const double* v = x_cur + a.dp.off_v + 3 * static_cast<int64_t>(f);
      double* vo = x_new + a.dp.off_v + 3 * static_cast<int64_t>(f);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double nv = (FD == 9) ? v[k] + du[(FD == 9) ? 6 + k : 0] : v[k];
        vo[k] = nv;
        if (FD == 9 && !is_ghost) {
          acc[2] += (nv - v[k]) * (nv - v[k]);
          acc[3] += nv * nv;
        }
      }
    }
  }
  if (lane == 0)
    for (int q = 0; q < 4; ++q) part[warp][q] = acc[q];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) {
      double s = 0.0;
      for (int w = 0; w < kUpdateWarps; ++w) s += part[w][q];
      a.step_part[4 * static_cast<int64_t>(blockIdx.x) + q] = s;
    }
  }
  // globals: cameras + IMU parameters (one thread of block 0)
  if (blockIdx.x == 0 && threadIdx.x == 32 * kUpdateWarps - 1) {
    double g4[4] = {0.0, 0.0, 0.0, 0.0};
    const double* sc = a.scale + nfp;
    for (int c = 0; c < a.dp.n_cams; ++c) {
      const CamInfo& ci = a.dp.cams[c];
      const double* x = x_cur + a.dp.off_cam + kCamStateStride * c;
      double* xo = x_new + a.dp.off_cam + kCamStateStride * c;
      double du[3];
      for (int k = 0; k < 3; ++k) du[k] = dc[ci.goff + k] * sc[ci.goff + k];
      double qo[4];
      so3_plus(x, du, qo);
      for (int k = 0; k < 4; ++k) xo[k] = qo[k];
      for (int k = 0; k < 3; ++k) xo[4 + k] = x[4 + k] + dc[ci.goff + 3 + k] * sc[ci.goff + 3 + k];
      for (int k = 0; k < 10; ++k)
        xo[7 + k] = x[7 + k] + (k < ci.K ? dc[ci.goff + 6 + k] * sc[ci.goff + 6 + k] : 0.0);
      for (int k = 0; k < 7 + ci.K; ++k) {
        g4[2] += (xo[k] - x[k]) * (xo[k] - x[k]);
        g4[3] += xo[k] * xo[k];
      }
    }
    {
      const double* x = x_cur + a.dp.off_imu;
      double* xo = x_new + a.dp.off_imu;
      for (int k = 0; k < kImuStateSize; ++k) {
        const double dd = a.dp.inertial ? dc[a.dp.imu_goff + k] * sc[a.dp.imu_goff + k] : 0.0;
        xo[k] = x[k] + dd;
        if (a.dp.inertial) {
          g4[2] += dd * dd;
          g4[3] += xo[k] * xo[k];
        }
      }
    }
    for (int k = 0; k < G; ++k) {
      const double d2 = a.D2x ? a.D2x[nfp + k] : lm_damp(b.C[k * G + k], sc[k], rinv);
      g4[0] += dc[k] * b.gc[k] * sc[k];
      g4[1] += dc[k] * dc[k] * d2;
    }
    for (int q = 0; q < 4; ++q) a.step_part[4 * static_cast<int64_t>(gridDim.x) + q] = g4[q];
  }
}

}  // namespace vc
