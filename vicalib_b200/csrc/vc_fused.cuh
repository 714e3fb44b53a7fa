// Fused evaluate + block-build kernel for the reprojection terms.
//
// One CTA per frame.  For each camera that sees the frame, every thread evaluates one grid corner
// (pose chain, camera model, analytic Jacobian, SoftLOne corrector) and parks its two Jacobian rows
// in a shared-memory tile, column-major over corners; the CTA then forms the frame's normal-equation
// blocks  [Jf Jg r]^T [Jf Jg r]  (24 x 24, symmetric) from that tile with FP64 tensor-core
// mma.sync.m8n8k4 (DMMA) — both operand fragments are read straight from the tile.  The Jacobian
// never goes to HBM: the pass reads 44 B per corner and writes ~2 KB per (frame, camera).
// Replaces the residual+Jacobian evaluation and J^T J build Ceres does inside ceres::Solve
// (vicalibrator.h:956) for the ImuReprojectionCostFunctor blocks (ceres-cost-functions.h:342-377).
#pragma once
#include "vc_internal.h"
#include "vc_math.cuh"

namespace vc {

constexpr int kFusedThreads = 160;                   // 5 warps
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedChunk = 144;                     // corners staged per pass
constexpr int kFusedLd = 2 * kFusedChunk + 4;        // tile leading dimension (== 4 mod 16: conflict-free fragments)
constexpr int kFusedCols = 24;                       // 3 column blocks of 8 (W <= 21)
static_assert(kFusedLd % 16 == 4, "fragment loads need ld == 4 (mod 16)");

struct FusedArgs {
  DevProblem dp;
  const Ctl* ctl;
  int which;
  const double* state[2];
  const int32_t *grp_start, *grp_count, *group_of;
  const double *pw, *pc;  // AoS [n_obs][3], [n_obs][2]
  int64_t n_obs;
  const double* mask;  // [G]
  Blocks out[2];
  double* Cg;          // [n_groups][kCgStride]
  double* cost_part;   // [n_frames]
  // apply_update: this launch first forms the trial state x_new = x_cur (+) step for its frame (and
  // block 0 for the globals) — the back-substitution of the arrow solve — then evaluates there.
  int apply_update;
  const double* scale;
  const double* D2x;   // explicit damping or null (LM rule)
  const double* X;     // [nf][FD][G+1] per-frame solutions, or null when the chain solver wrote delta
  double* delta;       // scaled step [nf*FD + G]
  double* states_rw[2];
  double* step_part;   // [n_frames + 1][4]
};

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// residual + tangent Jacobian of one observation, loss-corrected and masked, written to the tile
template <int MODEL>
__device__ __forceinline__ double eval_obs_to_tile(const double* T, const double* cam, const double* mask, V3 pw, double pcu,
                                                   double pcv, double mult, double* tile, int k0, int k1) {
  constexpr int K = Cam<MODEL>::K, NT = 12 + K;
  const Q4 q{T[0], T[1], T[2], T[3]};
  const V3 t{T[4], T[5], T[6]};
  const V3 pk = qrot(qconj(q), pw - t);
  const Q4 qc{cam[0], cam[1], cam[2], cam[3]};
  double R[9];
  qmat(qc, R);
  const V3 pc = mat_mul(R, pk) + V3{cam[4], cam[5], cam[6]};
  double z[2], dzp[6], dzi[2 * K];
  Cam<MODEL>::project(pc, cam + 7, z, dzp, dzi);
  const double r0 = z[0] - pcu, r1 = z[1] - pcv;
  double rho0, rho1;
  soft_l_one(r0 * r0 + r1 * r1, &rho0, &rho1);
  const double sc = sqrt(rho1);
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    const double* d = dzp + 3 * row;
    const double m0 = d[0] * R[0] + d[1] * R[3] + d[2] * R[6];
    const double m1 = d[0] * R[1] + d[1] * R[4] + d[2] * R[7];
    const double m2 = d[0] * R[2] + d[1] * R[5] + d[2] * R[8];
    const double w0 = m1 * pk.z - m2 * pk.y;
    const double w1 = m2 * pk.x - m0 * pk.z;
    const double w2 = m0 * pk.y - m1 * pk.x;
    double* o = tile + (row == 0 ? k0 : k1);
    o[0 * kFusedLd] = -m0 * sc;
    o[1 * kFusedLd] = -m1 * sc;
    o[2 * kFusedLd] = -m2 * sc;
    o[3 * kFusedLd] = w0 * sc;
    o[4 * kFusedLd] = w1 * sc;
    o[5 * kFusedLd] = w2 * sc;
    o[6 * kFusedLd] = -w0 * sc * mask[0];
    o[7 * kFusedLd] = -w1 * sc * mask[1];
    o[8 * kFusedLd] = -w2 * sc * mask[2];
    o[9 * kFusedLd] = d[0] * sc * mask[3];
    o[10 * kFusedLd] = d[1] * sc * mask[4];
    o[11 * kFusedLd] = d[2] * sc * mask[5];
#pragma unroll
    for (int k = 0; k < K; ++k) o[(12 + k) * kFusedLd] = dzi[row * K + k] * sc * mask[6 + k];
    o[NT * kFusedLd] = (row == 0 ? r0 : r1) * sc;
#pragma unroll
    for (int k = NT + 1; k < kFusedCols; ++k) o[k * kFusedLd] = 0.0;
  }
  return 0.5 * rho0 * mult;
}

template <int FD>
__global__ void __launch_bounds__(kFusedThreads) fused_build_kernel(FusedArgs a) {
  extern __shared__ double smem[];
  if (a.ctl->done) return;
  const int buf = a.which ? 1 - a.ctl->cur : a.ctl->cur;
  const double* state = a.state[buf];
  const Blocks& out = a.out[buf];
  double* tile = smem;                                   // [24][kFusedLd]
  double* red = tile + kFusedCols * kFusedLd;            // [warps][6][64]
  double* smB = red + kFusedWarps * 6 * 64;              // [FD*FD]
  double* smg = smB + FD * FD;                           // [FD]
  double* wcost = smg + FD;                              // [warps]
  double* smT = wcost + kFusedWarps;                     // [7] trial frame pose
  double* smCam = smT + 8;                               // [kMaxCams][kCamStateStride] trial camera states
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = a.dp.G, nf = a.dp.n_frames;
  for (int k = tid; k < FD * FD + FD; k += kFusedThreads) smB[k] = 0.0;
  double* Ef = out.E + static_cast<int64_t>(f) * FD * G;
  for (int k = tid; k < FD * G; k += kFusedThreads) Ef[k] = 0.0;
  const double* T = state + 7 * static_cast<int64_t>(f);
  const double* camBase = state + a.dp.off_cam;
  if (a.apply_update) {
    const int cur = a.ctl->cur;
    const Blocks& bc = a.out[cur];
    const double* x_cur = a.states_rw[cur];
    double* x_new = a.states_rw[1 - cur];
    const double rinv = 1.0 / a.ctl->radius;
    const int M = G + 1;
    const int64_t nfp = static_cast<int64_t>(nf) * FD;
    const double* dc = a.delta + nfp;
    const double* scg = a.scale + nfp;
    if (warp == 0) {  // this frame's step and pose
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
        double acc[4] = {0.0, 0.0, 0.0, 0.0}, du[FD];
#pragma unroll
        for (int r = 0; r < FD; ++r) {
          const int64_t k = static_cast<int64_t>(f) * FD + r;
          const double sc = a.scale[k];
          const double d2 = a.D2x ? a.D2x[k] : lm_damp(bc.B[k * FD + r], sc, rinv);
          a.delta[k] = d[r];
          acc[0] += d[r] * bc.gf[k] * sc;
          acc[1] += d[r] * d[r] * d2;
          du[r] = d[r] * sc;
        }
        const double* x = x_cur + 7 * static_cast<int64_t>(f);
        double xo[7];
        se3_plus(x, du, xo);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          x_new[7 * static_cast<int64_t>(f) + k] = xo[k];
          smT[k] = xo[k];
          acc[2] += (xo[k] - x[k]) * (xo[k] - x[k]);
          acc[3] += xo[k] * xo[k];
        }
        const double* v = x_cur + a.dp.off_v + 3 * static_cast<int64_t>(f);
        double* vo = x_new + a.dp.off_v + 3 * static_cast<int64_t>(f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double nv = (FD == 9) ? v[k] + du[(FD == 9) ? 6 + k : 0] : v[k];
          vo[k] = nv;
          if (FD == 9) {
            acc[2] += (nv - v[k]) * (nv - v[k]);
            acc[3] += nv * nv;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) a.step_part[4 * static_cast<int64_t>(f) + q] = acc[q];
      }
    } else if (warp == 1 && lane < a.dp.n_cams) {  // trial camera states (every CTA needs them)
      const int c = lane;
      const CamInfo& ci = a.dp.cams[c];
      const double* x = x_cur + a.dp.off_cam + kCamStateStride * c;
      double* xo = smCam + kCamStateStride * c;
      double du[3], qo[4];
      for (int k = 0; k < 3; ++k) du[k] = dc[ci.goff + k] * scg[ci.goff + k];
      so3_plus(x, du, qo);
      for (int k = 0; k < 4; ++k) xo[k] = qo[k];
      for (int k = 0; k < 3; ++k) xo[4 + k] = x[4 + k] + dc[ci.goff + 3 + k] * scg[ci.goff + 3 + k];
      for (int k = 0; k < 10; ++k) xo[7 + k] = x[7 + k] + (k < ci.K ? dc[ci.goff + 6 + k] * scg[ci.goff + 6 + k] : 0.0);
      if (f == 0) {
        double* xg = x_new + a.dp.off_cam + kCamStateStride * c;
        for (int k = 0; k < kCamStateStride; ++k) xg[k] = xo[k];
      }
    } else if (warp == 2 && lane == 0 && f == 0) {  // IMU globals + the globals' share of the step reductions
      double g4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int c = 0; c < a.dp.n_cams; ++c) {
        const CamInfo& ci = a.dp.cams[c];
        const double* x = x_cur + a.dp.off_cam + kCamStateStride * c;
        double du[3], qo[4];
        for (int k = 0; k < 3; ++k) du[k] = dc[ci.goff + k] * scg[ci.goff + k];
        so3_plus(x, du, qo);
        for (int k = 0; k < 4; ++k) { g4[2] += (qo[k] - x[k]) * (qo[k] - x[k]); g4[3] += qo[k] * qo[k]; }
        for (int k = 0; k < 3 + ci.K; ++k) {
          const double dd = dc[ci.goff + 3 + k] * scg[ci.goff + 3 + k], nv = x[4 + k] + dd;
          g4[2] += dd * dd;
          g4[3] += nv * nv;
        }
      }
      const double* x = x_cur + a.dp.off_imu;
      double* xo = x_new + a.dp.off_imu;
      for (int k = 0; k < kImuStateSize; ++k) {
        const double dd = a.dp.inertial ? dc[a.dp.imu_goff + k] * scg[a.dp.imu_goff + k] : 0.0;
        xo[k] = x[k] + dd;
        if (a.dp.inertial) { g4[2] += dd * dd; g4[3] += xo[k] * xo[k]; }
      }
      for (int k = 0; k < G; ++k) {
        const double d2 = a.D2x ? a.D2x[nfp + k] : lm_damp(bc.C[k * G + k], scg[k], rinv);
        g4[0] += dc[k] * bc.gc[k] * scg[k];
        g4[1] += dc[k] * dc[k] * d2;
      }
      for (int q = 0; q < 4; ++q) a.step_part[4 * static_cast<int64_t>(nf) + q] = g4[q];
    }
    __syncthreads();
    T = smT;
    camBase = smCam;
  }
  double cost = 0.0;
  for (int c = 0; c < a.dp.n_cams; ++c) {
    const int g = a.group_of[c * nf + f];
    if (g < 0) continue;
    const CamInfo& ci = a.dp.cams[c];
    const int W = 13 + ci.K;
    const int start = a.grp_start[g], cnt = a.grp_count[g];
    const double* cam = camBase + kCamStateStride * c;
    const double* mask = a.mask + ci.goff;
    double acc[6][2];
#pragma unroll
    for (int b = 0; b < 6; ++b) acc[b][0] = acc[b][1] = 0.0;
    for (int ch = 0; ch < cnt; ch += kFusedChunk) {
      const int m = min(kFusedChunk, cnt - ch);
      const int m4 = (m + 3) & ~3;  // k range padded to the MMA depth
      __syncthreads();              // previous pass is done with the tile
      if (tid < m) {
        const int64_t i = start + ch + tid;
        const V3 pw{a.pw[3 * i], a.pw[3 * i + 1], a.pw[3 * i + 2]};
        const double pcu = a.pc[2 * i], pcv = a.pc[2 * i + 1];
        switch (ci.model) {
          case kLinear: cost += eval_obs_to_tile<kLinear>(T, cam, mask, pw, pcu, pcv, a.dp.visual_mult, tile, tid, kFusedChunk + tid); break;
          case kFov: cost += eval_obs_to_tile<kFov>(T, cam, mask, pw, pcu, pcv, a.dp.visual_mult, tile, tid, kFusedChunk + tid); break;
          case kPoly2: cost += eval_obs_to_tile<kPoly2>(T, cam, mask, pw, pcu, pcv, a.dp.visual_mult, tile, tid, kFusedChunk + tid); break;
          case kPoly3: cost += eval_obs_to_tile<kPoly3>(T, cam, mask, pw, pcu, pcv, a.dp.visual_mult, tile, tid, kFusedChunk + tid); break;
          default: cost += eval_obs_to_tile<kKb4>(T, cam, mask, pw, pcu, pcv, a.dp.visual_mult, tile, tid, kFusedChunk + tid); break;
        }
      } else if (tid < m4) {  // zero the padding rows of both halves
#pragma unroll
        for (int k = 0; k < kFusedCols; ++k) {
          tile[k * kFusedLd + tid] = 0.0;
          tile[k * kFusedLd + kFusedChunk + tid] = 0.0;
        }
      }
      __syncthreads();
      // SYRK over k: steps [0, m4/4) cover residual row 0, [m4/4, m4/2) residual row 1
      const int ns = m4 >> 2;
      const double* frag = tile + (lane >> 2) * kFusedLd + (lane & 3);
      for (int s = warp; s < 2 * ns; s += kFusedWarps) {
        const int k0 = s < ns ? 4 * s : kFusedChunk + 4 * (s - ns);
        const double a0 = frag[k0], a1 = frag[8 * kFusedLd + k0], a2 = frag[16 * kFusedLd + k0];
        dmma_m8n8k4(acc[0][0], acc[0][1], a0, a0);  // (0,0)
        dmma_m8n8k4(acc[1][0], acc[1][1], a1, a0);  // (1,0)
        dmma_m8n8k4(acc[2][0], acc[2][1], a1, a1);  // (1,1)
        dmma_m8n8k4(acc[3][0], acc[3][1], a2, a0);  // (2,0)
        dmma_m8n8k4(acc[4][0], acc[4][1], a2, a1);  // (2,1)
        dmma_m8n8k4(acc[5][0], acc[5][1], a2, a2);  // (2,2)
      }
    }
    // cross-warp reduction of the six 8x8 blocks
    {
      double* rw = red + warp * 384 + (lane >> 2) * 8 + 2 * (lane & 3);
#pragma unroll
      for (int b = 0; b < 6; ++b) { rw[b * 64] = acc[b][0]; rw[b * 64 + 1] = acc[b][1]; }
    }
    __syncthreads();
    double* Cgg = a.Cg + static_cast<int64_t>(g) * kCgStride;
    for (int e = tid; e < 384; e += kFusedThreads) {
      const int b = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
      const int bi = b == 0 ? 0 : b < 3 ? 1 : 2;
      const int bj = b == 0 ? 0 : b == 1 ? 0 : b == 2 ? 1 : b - 3;
      const int i = bi * 8 + rr, j = bj * 8 + cc;
      if (j > i || i >= W) continue;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kFusedWarps; ++w) v += red[w * 384 + e];
      v *= a.dp.visual_mult;
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
  // block cost
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost += __shfl_down_sync(0xffffffffu, cost, o);
  if (lane == 0) wcost[warp] = cost;
  __syncthreads();
  double* Bf = out.B + static_cast<int64_t>(f) * FD * FD;
  for (int k = tid; k < FD * FD; k += kFusedThreads) Bf[k] = smB[k];
  for (int k = tid; k < FD; k += kFusedThreads) out.gf[static_cast<int64_t>(f) * FD + k] = smg[k];
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kFusedWarps; ++w) s += wcost[w];
    a.cost_part[f] = s;
  }
}

}  // namespace vc
