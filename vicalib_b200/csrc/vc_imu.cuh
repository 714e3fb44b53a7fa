// IMU residual family on the device and the block-tridiagonal frame chain it induces.
//
//   imu_eval_kernel      one warp per frame interval, one tangent direction per lane (forward-mode
//                        dual numbers through the RK4 integrator): r (9) and J (9x33)
//                        [SwitchedFullImuCostFunction, ceres-cost-functions.h:379-490]
//   imu_accumulate_kernel adds J^T J of the two intervals touching a frame into the frame blocks
//                        (B, off-diagonal U, E columns of the 15 IMU globals, gradients)
//   chain_*              partitioned elimination of the block-tridiagonal + arrow system
//   imu_weights_kernel   UpdateImuWeights (vicalibrator.h:723-799)
#pragma once
#include <algorithm>
#include <vector>

#include "vc_imu_math.cuh"
#include "vc_internal.h"

namespace vc {

constexpr int kImuCgStride = 136;  // 15*16/2 = 120 packed sym + 15 gradient (+1 pad)

// ---------------------------------------------------------------- IMU evaluate
struct ImuEvalArgs {
  DevProblem dp;
  imu::ImuBuf buf;
  const Ctl* ctl;
  int which;
  const double* states[2];
  const double* ftime;
  const double* wsqrt;  // [ni][81]
  const double* mask;   // 15 IMU global columns
  double *r, *J, *cost;
  int ni, apply_loss;
  double mult;
};

// column `col` (0..5) of LocalParamSe3::ComputeJacobian (local-param-se3.h:61-88), 7 entries
__device__ __forceinline__ void se3_local_column(const double* x, int col, double out[7]) {
  const double q1 = x[0], q2 = x[1], q3 = x[2], q0 = x[3];
#pragma unroll
  for (int i = 0; i < 7; ++i) out[i] = 0.0;
  switch (col) {
    case 0: out[4] = 1.0 - 2.0 * (q2 * q2 + q3 * q3); out[5] = 2.0 * (q1 * q2 + q0 * q3); out[6] = 2.0 * (q1 * q3 - q0 * q2); break;
    case 1: out[4] = 2.0 * (q1 * q2 - q0 * q3); out[5] = 1.0 - 2.0 * (q1 * q1 + q3 * q3); out[6] = 2.0 * (q2 * q3 + q0 * q1); break;
    case 2: out[4] = 2.0 * (q1 * q3 + q0 * q2); out[5] = 2.0 * (q2 * q3 - q0 * q1); out[6] = 1.0 - 2.0 * (q1 * q1 + q2 * q2); break;
    case 3: out[0] = 0.5 * q0; out[1] = 0.5 * q3; out[2] = -0.5 * q2; out[3] = -0.5 * q1; break;
    case 4: out[0] = -0.5 * q3; out[1] = 0.5 * q0; out[2] = 0.5 * q1; out[3] = -0.5 * q2; break;
    case 5: out[0] = 0.5 * q2; out[1] = -0.5 * q1; out[2] = 0.5 * q0; out[3] = -0.5 * q3; break;
  }
}

// the same fields as ImuEvalArgs / ImuAccArgs with the problem description by reference
struct ImuEvalView {
  const DevProblem& dp;
  imu::ImuBuf buf;
  const double* ftime;
  const double* wsqrt;
  const double* mask;
  double *r, *J, *cost;
  int ni, apply_loss;
  double mult;
};
struct ImuAccView {
  const DevProblem& dp;
  const double *r, *J;
  double* Cg;
  int ni;
};

constexpr int kImuWarps = 4;

// residual (9) and tangent Jacobian (9 x 33) of interval k by one warp: lane -> Jacobian column
// A: ImuEvalArgs or a view with the same fields (the persistent kernel binds `dp` by reference to its own parameter)
template <class A>
__device__ __forceinline__ void imu_eval_interval(const A& a, int k, int lane, const double* state) {
  using imu::D1;
  // lane -> Jacobian column: pose2 0-5 | pose1 6-11 | (v2 12-14 analytic) | v1 15-17 | g 18-19 | b 20-25 | sf 26-31 | ts 32
  const int col = lane < 12 ? lane : lane + 3;
  const double* X2 = state + 7 * static_cast<int64_t>(k + 1);
  const double* X1 = state + 7 * static_cast<int64_t>(k);
  const double* V2 = state + a.dp.off_v + 3 * static_cast<int64_t>(k + 1);
  const double* V1 = state + a.dp.off_v + 3 * static_cast<int64_t>(k);
  const double* P = state + a.dp.off_imu;  // g2 b6 sf6 ts
  double s2[7], s1[7];
  se3_local_column(X2, col < 6 ? col : 0, s2);
  se3_local_column(X1, (col >= 6 && col < 12) ? col - 6 : 0, s1);
  D1 x2[7], x1[7], v2[3], v1[3], g2[2], b[6], sf[6], ts;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    x2[i] = D1(X2[i], col < 6 ? s2[i] : 0.0);
    x1[i] = D1(X1[i], (col >= 6 && col < 12) ? s1[i] : 0.0);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v2[i] = D1(V2[i], 0.0);
    v1[i] = D1(V1[i], col == 15 + i ? 1.0 : 0.0);
  }
  g2[0] = D1(P[0], col == 18 ? 1.0 : 0.0);
  g2[1] = D1(P[1], col == 19 ? 1.0 : 0.0);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    b[i] = D1(P[2 + i], col == 20 + i ? 1.0 : 0.0);
    sf[i] = D1(P[8 + i], col == 26 + i ? 1.0 : 0.0);
  }
  ts = D1(P[14], col == 32 ? 1.0 : 0.0);
  D1 raw[9];
  const bool ok = imu::imu_raw_residual<D1>(a.buf, a.ftime[k], a.ftime[k + 1], x2, x1, v2, v1, g2, b, sf, ts, raw);
  const double* W = a.wsqrt + static_cast<int64_t>(k) * 81;
  double* rk = a.r + static_cast<int64_t>(k) * 9;
  double* Jk = a.J + static_cast<int64_t>(k) * 297;
  if (!ok) {  // no measurements: zero residual (ceres-cost-functions.h:452-455)
    if (lane == 0) {
      for (int j = 0; j < 9; ++j) rk[j] = 0.0;
      a.cost[k] = 0.0;
    }
    for (int j = 0; j < 9; ++j) {
      if (lane < 30) Jk[j * 33 + col] = 0.0;
      if (lane >= 30) { Jk[j * 33 + 12] = 0.0; Jk[j * 33 + 13] = 0.0; Jk[j * 33 + 14] = 0.0; }
    }
    return;
  }
  // residuals_vec = (r^T W)^T  (:476-477); rotation-only switch zeroes rows 0-2 and 6-8 (:479-482)
  D1 rw[9];
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double va = 0.0, vv = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      va += raw[i].a * W[i * 9 + j];
      vv += raw[i].v * W[i * 9 + j];
    }
    const bool off = a.dp.rotation_only && (j < 3 || j >= 6);
    rw[j] = off ? D1(0.0, 0.0) : D1(va, vv);
    s += rw[j].a * rw[j].a;
  }
  double sc = 1.0, cost = 0.5 * s;
  if (a.apply_loss) {
    double rho0, rho1;
    cauchy(s, &rho0, &rho1);  // CauchyLoss(100), vicalibrator.h:133
    cost = 0.5 * rho0 * a.mult;
    sc = sqrt(rho1);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 9; ++j) rk[j] = rw[j].a * sc;
    a.cost[k] = cost;
  }
  const double m = (col >= 18 && col < 33) ? a.mask[col - 18] : 1.0;  // lanes 30, 31 carry no column
  if (lane < 30) {
#pragma unroll
    for (int j = 0; j < 9; ++j) Jk[j * 33 + col] = rw[j].v * sc * m;
  } else if (lane == 30) {
    // d r / d v2 = -W[6+i][:]  (v_end - v2 enters rows 6-8 of the raw residual)
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const bool off = a.dp.rotation_only && (j < 3 || j >= 6);
#pragma unroll
      for (int i = 0; i < 3; ++i) Jk[j * 33 + 12 + i] = off ? 0.0 : -W[(6 + i) * 9 + j] * sc;
    }
  }
}

__global__ void __launch_bounds__(32 * kImuWarps) imu_eval_kernel(ImuEvalArgs a) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kImuWarps + (threadIdx.x >> 5);
  if (k >= a.ni || a.ctl->done) return;
  imu_eval_interval(a, k, lane, a.states[a.which ? 1 - a.ctl->cur : a.ctl->cur]);
}

// ---------------------------------------------------------------- IMU -> frame blocks
struct ImuAccArgs {
  DevProblem dp;
  const Ctl* ctl;
  int which;
  const double *r, *J;  // [ni][9], [ni][9*33] (loss-corrected, masked)
  Blocks outs[2];
  double* Cg;  // [ni][kImuCgStride]
  int ni;
};
// local column layout of an interval: frame k (pose1 6, v1 3) = 0..8 | frame k+1 (pose2 6, v2 3) = 9..17 | globals 18..32
__device__ __forceinline__ int imu_local_col(int c) {
  return c < 6 ? 6 + c : c < 9 ? 15 + (c - 6) : c < 15 ? (c - 9) : c < 18 ? 12 + (c - 15) : c;
}
// frame f by a group of NT threads: interval f-1 (f is its second frame) and interval f (f is its first frame).
// Jl: the group's [2][9][34] staging area ([which][row][local col 0..32, 33 = residual]); SYNC: the group's barrier
template <int NT, class A, class SYNC>
__device__ __forceinline__ void imu_accumulate_frame(const A& a, const Blocks& out, int f, int tid, double (*Jl)[9][34],
                                                     SYNC sync) {
  const int G = a.dp.G, nf = a.dp.n_frames, io = a.dp.imu_goff;
  const bool hasP = f > 0, hasN = f < nf - 1;
  double* Bf = out.B + static_cast<int64_t>(f) * 81;
  double* Uf = out.U + static_cast<int64_t>(f) * 81;
  double* Ef = out.E + static_cast<int64_t>(f) * 9 * G;
  double* gf = out.gf + static_cast<int64_t>(f) * 9;
  constexpr int kIn = 2 * 9 * 34, kOut = 81 + 81 + 135 + 9 + 135;
  constexpr int kItIn = (kIn + NT - 1) / NT, kItOut = (kOut + NT - 1) / NT;
  // every global load first (the two intervals' Jacobians and the entries the reprojection build left in B and the
  // gradient), then the arithmetic: a load-use loop would pay one memory round trip per iteration
  double vin[kItIn], old[kItOut];
#pragma unroll
  for (int it = 0; it < kItIn; ++it) {
    const int e = tid + it * NT;
    const int which = e / (9 * 34), row = (e / 34) % 9, c = e % 34;
    const int k = which == 0 ? f - 1 : f;
    vin[it] = 0.0;
    if (e < kIn && k >= 0 && k < a.ni)
      vin[it] = c < 33 ? a.J[static_cast<int64_t>(k) * 297 + row * 33 + imu_local_col(c)] : a.r[static_cast<int64_t>(k) * 9 + row];
  }
#pragma unroll
  for (int it = 0; it < kItOut; ++it) {
    const int e = tid + it * NT;
    old[it] = e < 81 ? Bf[e] : (e >= 297 && e < 306) ? gf[e - 297] : 0.0;
  }
#pragma unroll
  for (int it = 0; it < kItIn; ++it) {
    const int e = tid + it * NT;
    if (e < kIn) Jl[e / (9 * 34)][(e / 34) % 9][e % 34] = vin[it];
  }
  sync();
  const double m = a.dp.imu_mult;
  // B (81) | U (81) | E imu columns (9*15) | g (9) | Cg of interval f (120 + 15)
#pragma unroll
  for (int it = 0; it < kItOut; ++it) {
    const int e = tid + it * NT;
    if (e >= kOut) break;
    if (e < 81) {
      const int i = e / 9, j = e % 9;
      double s = 0.0;
#pragma unroll
      for (int row = 0; row < 9; ++row) {
        if (hasP) s += Jl[0][row][9 + i] * Jl[0][row][9 + j];
        if (hasN) s += Jl[1][row][i] * Jl[1][row][j];
      }
      Bf[e] = old[it] + s * m;
    } else if (e < 162) {
      const int i = (e - 81) / 9, j = (e - 81) % 9;  // U[f] = H[f-1, f] = J1(f-1)^T J2(f-1)
      double s = 0.0;
      if (hasP) {
#pragma unroll
        for (int row = 0; row < 9; ++row) s += Jl[0][row][i] * Jl[0][row][9 + j];
      }
      Uf[e - 81] = s * m;
    } else if (e < 297) {
      const int i = (e - 162) / 15, j = (e - 162) % 15;
      double s = 0.0;
#pragma unroll
      for (int row = 0; row < 9; ++row) {
        if (hasP) s += Jl[0][row][9 + i] * Jl[0][row][18 + j];
        if (hasN) s += Jl[1][row][i] * Jl[1][row][18 + j];
      }
      Ef[i * G + io + j] = s * m;
    } else if (e < 306) {
      const int i = e - 297;
      double s = 0.0;
#pragma unroll
      for (int row = 0; row < 9; ++row) {
        if (hasP) s += Jl[0][row][9 + i] * Jl[0][row][33];
        if (hasN) s += Jl[1][row][i] * Jl[1][row][33];
      }
      gf[i] = old[it] + s * m;
    } else if (hasN) {
      const int q = e - 306;  // packed lower triangle of the 15x15 global block, then gradient
      double s = 0.0;
      if (q < 120) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        const int j = q - i * (i + 1) / 2;
#pragma unroll
        for (int row = 0; row < 9; ++row) s += Jl[1][row][18 + i] * Jl[1][row][18 + j];
      } else {
#pragma unroll
        for (int row = 0; row < 9; ++row) s += Jl[1][row][18 + (q - 120)] * Jl[1][row][33];
      }
      a.Cg[static_cast<int64_t>(f) * kImuCgStride + q] = s * m;
    }
  }
}
__global__ void __launch_bounds__(128) imu_accumulate_kernel(ImuAccArgs a) {
  __shared__ double Jl[2][9][34];
  if (a.ctl->done) return;
  imu_accumulate_frame<128>(a, a.outs[a.which ? 1 - a.ctl->cur : a.ctl->cur], blockIdx.x, threadIdx.x, Jl, [] { __syncthreads(); });
}

}  // namespace vc
