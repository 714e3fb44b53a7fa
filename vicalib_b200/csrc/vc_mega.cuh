// Persistent trust-region kernel for the vision-only problem on one GPU.
//
// One cooperative launch runs up to n_iters Levenberg-Marquardt iterations; one CTA per SM, every
// frame owned by the same CTA for the whole solve.  An iteration is five phases separated by two grid
// barriers (instead of six launches):
//
//   S  per frame (warp):  L = chol(B_f + D_f),  X_f = (B_f + D_f)^-1 [E_f | g_f],  CTA partial of E^T X
//   ------------------------------------------------------------------------------- grid barrier
//   G  every CTA, redundantly: sum the 148 partials in fixed order, factor the reduced system
//      C + D - sum E^T X, solve for the globals' step
//   U  per frame (warp):  back-substitution, x (+) step, step statistics; trial camera states
//   B  per frame (team of 5 warps): fused evaluate + Gram build at the trial point (vc_fused.cuh),
//      the camera blocks accumulated per team in shared memory (no per-group global round trip)
//   ------------------------------------------------------------------------------- grid barrier
//   D  every CTA, redundantly: sum the partial global blocks / scalars in fixed order, then the
//      accept-reject decision of Ceres' TrustRegionMinimizer (decide_step) on its own copy of Ctl
//
// Every CTA sees bit-identical sums (same order), so every CTA takes the same decision and the loop
// needs no broadcast.  Replaces the body of ceres::Solve (vicalibrator.h:956) for the staged vision
// solves; the inertial and the frame-sharded paths keep the multi-launch engine (vc_engine.inl).
#pragma once
#include <cooperative_groups.h>

#include "vc_fused.cuh"
#include "vc_kernels.cuh"

namespace vc {
namespace cg = cooperative_groups;

constexpr int kTeamThreads = kFusedThreads;
constexpr int kTeamWarps = kFusedWarps;
constexpr int kMegaMaxTeams = 4;
constexpr int kMegaMaxThreads = kMegaMaxTeams * kTeamThreads;
// per team: tile | per-warp partials | Gram matrix | frame block | frame gradient (+ pad)
constexpr int kTeamDoubles = kFusedCols * kFusedLd + kTeamWarps * kFusedRed + 256 + 36 + 8;
constexpr int kMegaPartExtra = 8;  // scalars appended to each CTA's partial slot
enum { kPCost = 0, kPGf2, kPDotG, kPDotD, kPStep2, kPXnorm2, kPGfMax, kPNotPD };
enum { kProfS = 0, kProfG, kProfU, kProfB, kProfD, kProfSyncA, kProfSyncB, kProfCount };

struct MegaArgs {
  DevProblem dp;
  Ctl* ctl;
  double* state[2];
  Blocks blk[2];
  const int32_t *grp_start, *grp_count, *group_of;
  const double *pw, *pc;
  const double* mask;
  const double* scale;   // Jacobi scale [nf*6 + G]
  double* X;             // [nf][6][G+1]
  double* partS;         // [grid][NS + 8]
  double* partC;         // [grid][NS + 8]
  double* totS;          // [NS + 8] reduced over the grid
  double* totC;          // [NS + 8]
  double* delta;         // scaled step [nf*6 + G]
  double* scalars;       // kSc* of the last evaluated point (for the host)
  int n_iters;
  int n_teams;
  unsigned long long* prof;  // [kProfCount] ns per phase (CTA 0), or null
};

__host__ __device__ inline size_t mega_smem_doubles(int G, int n_teams) {
  const size_t NS = static_cast<size_t>(G) * G + G;
  return static_cast<size_t>(n_teams) * kTeamDoubles + (2 + n_teams) * NS + G + kMaxCams * (kCamStateStride + 9) + kScCount + 64 +
         sizeof(Ctl) / sizeof(double) + 8;
}

__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(kTeamThreads) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Two-stage fixed-order reduction of the CTAs' partial slots (the grid barrier is ~1 us, while every CTA
// reading every slot is ~50 MB of L2 traffic):  stage 1, entry e is summed over all CTAs by one warp of
// CTA (e mod grid) into tot[e];  (grid barrier)  stage 2, every CTA copies tot[] to shared memory.
// Entries listed in max_a / max_b combine with max instead of +.
__device__ inline void mega_reduce_stage1(const double* part, int stride, int nparts, int n, double* tot, int max_a, int max_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int e = blockIdx.x + warp * gridDim.x; e < n; e += nwarps * gridDim.x) {
    const bool is_max = e == max_a || e == max_b;
    const double* p = part + e;
    double s0 = 0.0, s1 = 0.0;
    int b = lane;
    if (is_max) {
      for (; b < nparts; b += 32) s0 = fmax(s0, __ldcg(p + static_cast<int64_t>(b) * stride));
    } else {
      for (; b + 32 < nparts; b += 64) {
        s0 += __ldcg(p + static_cast<int64_t>(b) * stride);
        s1 += __ldcg(p + static_cast<int64_t>(b + 32) * stride);
      }
      if (b < nparts) s0 += __ldcg(p + static_cast<int64_t>(b) * stride);
      s0 += s1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(0xffffffffu, s0, o);
      s0 = is_max ? fmax(s0, t) : s0 + t;
    }
    if (lane == 0) tot[e] = s0;
  }
}

__global__ void __launch_bounds__(kMegaMaxThreads, 1) lm_mega_kernel(MegaArgs a) {
  extern __shared__ double smem[];
  cg::grid_group grid = cg::this_grid();
  constexpr int FD = 6;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int bid = blockIdx.x, nb = gridDim.x;
  const int team = tid / kTeamThreads, ttid = tid - team * kTeamThreads, twarp = ttid >> 5;
  const int n_teams = a.n_teams;
  const int G = a.dp.G, M = G + 1, NS = G * G + G, PS = NS + kMegaPartExtra, nf = a.dp.n_frames;
  const int64_t nfp = static_cast<int64_t>(nf) * FD;
  const int nk = bid < nf ? (nf - bid + nb - 1) / nb : 0;  // frames of this CTA: f = bid + k * nb

  double* teams = smem;                                            // [n_teams][kTeamDoubles]; also phase scratch
  double* Cacc = teams + static_cast<size_t>(n_teams) * kTeamDoubles;  // [2][NS]  C | gc of the two points
  double* Cteam = Cacc + 2 * NS;                                   // [n_teams][NS]
  double* dg = Cteam + static_cast<size_t>(n_teams) * NS;          // [G] globals' step (scaled)
  double* smCam = dg + G;                                          // [kMaxCams][17] trial camera states
  double* smRc = smCam + kMaxCams * kCamStateStride;               // [kMaxCams][9]
  double* sc = smRc + kMaxCams * 9;                                // [kScCount]
  double* wred = sc + kScCount;                                    // [64] per-warp scalars
  Ctl* ctl = reinterpret_cast<Ctl*>(wred + 64);
  __shared__ int bad;
  double* Swork = Cteam;  // phases S / G / D: Schur accumulator, then the reduced system, then the reduced trial block
  const double* scg = a.scale + nfp;

  if (tid == 0) *ctl = *a.ctl;
  __syncthreads();
  if (ctl->done) return;
  {
    const Blocks& b0 = a.blk[ctl->cur];
    double* C0 = Cacc + ctl->cur * NS;
    for (int e = tid; e < NS; e += nthreads) C0[e] = e < G * G ? b0.C[e] : b0.gc[e - G * G];
  }
  unsigned long long t_prev = 0;
  const bool prof = a.prof != nullptr && bid == 0 && tid == 0;
  if (prof) t_prev = global_ns();
  auto mark = [&](int slot) {
    if (prof) {
      const unsigned long long t = global_ns();
      a.prof[slot] += t - t_prev;
      t_prev = t;
    }
  };

  for (int it = 0; it < a.n_iters; ++it) {
    __syncthreads();
    if (ctl->done) break;
    const int cur = ctl->cur;
    const Blocks& bc = a.blk[cur];
    const Blocks& bt = a.blk[1 - cur];
    const double* x_cur = a.state[cur];
    double* x_new = a.state[1 - cur];
    const double rinv = 1.0 / ctl->radius;
    const double* Ccur = Cacc + cur * NS;
    double notpd = 0.0;

    // ------------------------------------------------------------ S: per-frame solves + Schur partial
    {
      double* Sacc = Swork;
      double* scratch = teams;  // [nwarps][2][FD][M]
      for (int k = tid; k < NS; k += nthreads) Sacc[k] = 0.0;
      for (int base = 0; base < nk; base += nwarps) {
        const int k = base + warp;
        __syncthreads();
        if (k < nk) {
          const int f = bid + k * nb;
          double* Esw = scratch + static_cast<size_t>(warp) * (2 * FD * M);
          double* Xw = Esw + FD * M;
          const double* sf = a.scale + static_cast<int64_t>(f) * FD;
          const double* Bf = bc.B + static_cast<int64_t>(f) * FD * FD;
          const double* Ef = bc.E + static_cast<int64_t>(f) * FD * G;
          double s[FD], L[FD][FD], iL[FD];
#pragma unroll
          for (int i = 0; i < FD; ++i) s[i] = sf[i];
#pragma unroll
          for (int i = 0; i < FD; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
              const double bij = Bf[i * FD + j];
              double v = bij * s[i] * s[j];
              if (i == j) v += lm_damp(bij, s[i], rinv);
              L[i][j] = v;
            }
          bool ok = true;
#pragma unroll
          for (int j = 0; j < FD; ++j) {
            double d = L[j][j];
#pragma unroll
            for (int q = 0; q < j; ++q) d -= L[j][q] * L[j][q];
            if (!(d > 0.0)) { ok = false; d = 1.0; }
            d = sqrt(d);
            L[j][j] = d;
            const double inv = 1.0 / d;
            iL[j] = inv;
#pragma unroll
            for (int i = j + 1; i < FD; ++i) {
              double t = L[i][j];
#pragma unroll
              for (int q = 0; q < j; ++q) t -= L[i][q] * L[j][q];
              L[i][j] = t * inv;
            }
          }
          if (!ok) notpd = 1.0;
          double* Xf = a.X + static_cast<int64_t>(f) * FD * M;
          for (int c = lane; c < M; c += 32) {
            double x[FD];
            const double scc = c < G ? scg[c] : 1.0;
#pragma unroll
            for (int i = 0; i < FD; ++i) {
              const double r = (c < G ? Ef[i * G + c] : bc.gf[static_cast<int64_t>(f) * FD + i]) * s[i] * scc;
              Esw[i * M + c] = r;
              x[i] = r;
            }
#pragma unroll
            for (int i = 0; i < FD; ++i) {
              double t = x[i];
#pragma unroll
              for (int q = 0; q < i; ++q) t -= L[i][q] * x[q];
              x[i] = t * iL[i];
            }
#pragma unroll
            for (int i = FD - 1; i >= 0; --i) {
              double t = x[i];
#pragma unroll
              for (int q = i + 1; q < FD; ++q) t -= L[q][i] * x[q];
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
        const int nact = min(nwarps, nk - base);
        for (int e = tid; e < NS; e += nthreads) {
          const int ra = e < G * G ? e / G : e - G * G;
          const int cb = e < G * G ? e - ra * G : G;
          if (e < G * G && cb > ra) continue;  // lower triangle only
          double sum = 0.0;
          for (int w = 0; w < nact; ++w) {
            const double* Ew = scratch + static_cast<size_t>(w) * (2 * FD * M);
            const double* Xv = Ew + FD * M;
#pragma unroll
            for (int q = 0; q < FD; ++q) sum += Ew[q * M + ra] * Xv[q * M + cb];
          }
          Sacc[e] += sum;
        }
      }
      __syncthreads();
      double* out = a.partS + static_cast<int64_t>(bid) * PS;
      for (int k = tid; k < NS; k += nthreads) out[k] = Sacc[k];
      // any failed pivot in this CTA
      const unsigned any = __ballot_sync(0xffffffffu, notpd > 0.0);
      if (lane == 0) wred[warp] = any ? 1.0 : 0.0;
      __syncthreads();
      if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < nwarps; ++w) v = fmax(v, wred[w]);
        out[NS + kPNotPD] = v;
      }
    }
    mark(kProfS);
    grid.sync();
    mark(kProfSyncA);

    // ------------------------------------------------------------ G: reduced system, every CTA the same
    {
      double* S = Swork;  // [G*G] lower triangle, then rhs [G]
      mega_reduce_stage1(a.partS, PS, nb, PS, a.totS, NS + kPNotPD, -1);
      grid.sync();
      for (int e = tid; e < NS; e += nthreads) {
        const double p = __ldcg(a.totS + e);
        if (e < G * G) {
          const int r = e / G, c = e - r * G;
          if (c > r) continue;
          double v = Ccur[e] * scg[r] * scg[c] - p;
          if (r == c) v += lm_damp(Ccur[e], scg[r], rinv);
          S[e] = v;
        } else {
          const int r = e - G * G;
          S[e] = -Ccur[e] * scg[r] + p;
        }
      }
      if (tid == 0) { sc[kScNotPD] = __ldcg(a.totS + NS + kPNotPD); bad = 0; }
      __syncthreads();
      if (warp == 0) {
        // Cholesky of the reduced system with the right-hand side carried as row G (so L y = rhs comes for
        // free), reciprocal pivots kept for the back-substitution; one warp, rows spread over the lanes
        double* rhs = S + G * G;
        double* invd = dg;
        for (int j = 0; j < G; ++j) {
          double d = S[j * G + j];
          if (!(d > 0.0)) { if (lane == 0) bad = 1; d = 1.0; }
          const double inv = rsqrt(d);
          __syncwarp();
          if (lane == 0) { S[j * G + j] = d * inv; invd[j] = inv; }
          for (int r = j + 1 + lane; r <= G; r += 32) {
            double* row = r < G ? S + r * G : rhs;
            row[j] *= inv;
          }
          __syncwarp();
          for (int r = j + 1 + lane; r <= G; r += 32) {
            double* row = r < G ? S + r * G : rhs;
            const double lrj = row[j];
            const int qmax = min(r, G - 1);
            int q = j + 1;
            for (; q + 3 <= qmax; q += 4) {
              const double l0 = S[q * G + j], l1 = S[(q + 1) * G + j], l2 = S[(q + 2) * G + j], l3 = S[(q + 3) * G + j];
              const double r0 = row[q], r1 = row[q + 1], r2 = row[q + 2], r3 = row[q + 3];
              row[q] = r0 - lrj * l0; row[q + 1] = r1 - lrj * l1; row[q + 2] = r2 - lrj * l2; row[q + 3] = r3 - lrj * l3;
            }
            for (; q <= qmax; ++q) row[q] -= lrj * S[q * G + j];
          }
          __syncwarp();
        }
        for (int i = G - 1; i >= 0; --i) {  // L^T x = y
          const double xi = rhs[i] * invd[i];
          __syncwarp();
          if (lane == 0) rhs[i] = xi;
          for (int q = lane; q < i; q += 32) rhs[q] -= S[i * G + q] * xi;
          __syncwarp();
        }
      }
      __syncthreads();
      for (int i = tid; i < G; i += nthreads) {
        const double v = bad ? 0.0 : S[G * G + i];
        dg[i] = v;
        if (bid == 0) a.delta[nfp + i] = v;
      }
      if (tid == 0 && bad) sc[kScNotPD] = 1.0;
      __syncthreads();
    }
    mark(kProfG);

    // ------------------------------------------------------------ U: back-substitution, trial states
    double ustat[4] = {0.0, 0.0, 0.0, 0.0};  // lane 0 of each warp: dotG, dotD, step2, xnorm2
    {
      if (warp == nwarps - 1) {  // trial camera states (every CTA needs them); CTA 0 also stores the globals
        if (lane < a.dp.n_cams) {
          const int c = lane;
          const CamInfo& ci = a.dp.cams[c];
          const double* x = x_cur + a.dp.off_cam + kCamStateStride * c;
          double* xo = smCam + kCamStateStride * c;
          double du[3], qo[4], R[9];
          for (int q = 0; q < 3; ++q) du[q] = dg[ci.goff + q] * scg[ci.goff + q];
          so3_plus(x, du, qo);
          for (int q = 0; q < 4; ++q) xo[q] = qo[q];
          for (int q = 0; q < 3; ++q) xo[4 + q] = x[4 + q] + dg[ci.goff + 3 + q] * scg[ci.goff + 3 + q];
          for (int q = 0; q < 10; ++q) xo[7 + q] = x[7 + q] + (q < ci.K ? dg[ci.goff + 6 + q] * scg[ci.goff + 6 + q] : 0.0);
          qmat(Q4{qo[0], qo[1], qo[2], qo[3]}, R);
#pragma unroll
          for (int q = 0; q < 9; ++q) smRc[9 * c + q] = R[q];
          if (bid == 0) {
            double* xg = x_new + a.dp.off_cam + kCamStateStride * c;
            for (int q = 0; q < kCamStateStride; ++q) xg[q] = xo[q];
          }
        }
        __syncwarp();
        if (bid == 0 && lane == 0) {  // the globals' share of the step statistics, counted once
          for (int c = 0; c < a.dp.n_cams; ++c) {
            const CamInfo& ci = a.dp.cams[c];
            const double* x = x_cur + a.dp.off_cam + kCamStateStride * c;
            const double* xo = smCam + kCamStateStride * c;
            for (int q = 0; q < 7 + ci.K; ++q) {
              ustat[2] += (xo[q] - x[q]) * (xo[q] - x[q]);
              ustat[3] += xo[q] * xo[q];
            }
          }
          for (int q = 0; q < kImuStateSize; ++q) x_new[a.dp.off_imu + q] = x_cur[a.dp.off_imu + q];
          for (int q = 0; q < G; ++q) {
            const double d2 = lm_damp(Ccur[q * G + q], scg[q], rinv);
            ustat[0] += dg[q] * Ccur[G * G + q] * scg[q];
            ustat[1] += dg[q] * dg[q] * d2;
          }
        }
      }
      for (int k = warp; k < nk; k += nwarps) {
        const int f = bid + k * nb;
        const double* Xf = a.X + static_cast<int64_t>(f) * FD * M;
        double d[FD];
#pragma unroll
        for (int r = 0; r < FD; ++r) {
          double s = 0.0;
          for (int c = lane; c < G; c += 32) s += Xf[r * M + c] * dg[c];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          d[r] = -Xf[r * M + G] - s;
        }
        if (lane == 0) {
          double du[FD];
#pragma unroll
          for (int r = 0; r < FD; ++r) {
            const int64_t q = static_cast<int64_t>(f) * FD + r;
            const double s = a.scale[q];
            const double d2 = lm_damp(bc.B[q * FD + r], s, rinv);
            a.delta[q] = d[r];
            ustat[0] += d[r] * bc.gf[q] * s;
            ustat[1] += d[r] * d[r] * d2;
            du[r] = d[r] * s;
          }
          const double* x = x_cur + 7 * static_cast<int64_t>(f);
          double xo[7];
          se3_plus(x, du, xo);
#pragma unroll
          for (int q = 0; q < 7; ++q) {
            x_new[7 * static_cast<int64_t>(f) + q] = xo[q];
            ustat[2] += (xo[q] - x[q]) * (xo[q] - x[q]);
            ustat[3] += xo[q] * xo[q];
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) x_new[a.dp.off_v + 3 * static_cast<int64_t>(f) + q] = x_cur[a.dp.off_v + 3 * static_cast<int64_t>(f) + q];
        }
      }
      for (int e = tid; e < n_teams * NS; e += nthreads) Cteam[e] = 0.0;
      __syncthreads();
    }
    mark(kProfU);

    // ------------------------------------------------------------ B: fused evaluate + build at the trial point
    double cost = 0.0, gf2 = 0.0, gfmax = 0.0;
    if (team < n_teams) {
      double* tb = teams + static_cast<size_t>(team) * kTeamDoubles;
      double* tile = tb;
      double* red = tile + kFusedCols * kFusedLd;
      double* Gm = red + kTeamWarps * kFusedRed;
      double* smB = Gm + 256;
      double* smg = smB + 36;
      double* Ct = Cteam + static_cast<size_t>(team) * NS;
      for (int k = team; k < nk; k += n_teams) {
        const int f = bid + k * nb;
        const double* T = x_new + 7 * static_cast<int64_t>(f);
        team_sync(team);  // previous frame's stores of smB / smg are done
        for (int q = ttid; q < 42; q += kTeamThreads) smB[q] = 0.0;  // smB[36] | smg[6]
        double* Ef = bt.E + static_cast<int64_t>(f) * FD * G;
        for (int q = ttid; q < FD * G; q += kTeamThreads) Ef[q] = 0.0;
        for (int c = 0; c < a.dp.n_cams; ++c) {
          const int g = a.group_of[c * nf + f];
          if (g < 0) continue;
          const CamInfo& ci = a.dp.cams[c];
          const int K = ci.K, NG = 6 + K;
          const int start = a.grp_start[g], cnt = a.grp_count[g];
          const double* cam = smCam + kCamStateStride * c;
          const double* Rc = smRc + 9 * c;
          const double* mask = a.mask + ci.goff;
          double acc[3][2];
#pragma unroll
          for (int b = 0; b < 3; ++b) acc[b][0] = acc[b][1] = 0.0;
          for (int ch = 0; ch < cnt; ch += kFusedChunk) {
            const int m = min(kFusedChunk, cnt - ch);
            const int m4 = (m + 3) & ~3;
            team_sync(team);  // previous pass / epilogue is done with the tile and Gm
            if (ttid < m) {
              const int64_t i = start + ch + ttid;
              const V3 pw{a.pw[3 * i], a.pw[3 * i + 1], a.pw[3 * i + 2]};
              const double pcu = a.pc[2 * i], pcv = a.pc[2 * i + 1];
              switch (ci.model) {
                case kLinear: cost += eval_obs_to_tile<kLinear>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, tile, ttid, kFusedChunk + ttid); break;
                case kFov: cost += eval_obs_to_tile<kFov>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, tile, ttid, kFusedChunk + ttid); break;
                case kPoly2: cost += eval_obs_to_tile<kPoly2>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, tile, ttid, kFusedChunk + ttid); break;
                case kPoly3: cost += eval_obs_to_tile<kPoly3>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, tile, ttid, kFusedChunk + ttid); break;
                default: cost += eval_obs_to_tile<kKb4>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, tile, ttid, kFusedChunk + ttid); break;
              }
            } else if (ttid < m4) {
#pragma unroll
              for (int q = 0; q < kFusedCols; ++q) {
                tile[q * kFusedLd + ttid] = 0.0;
                tile[q * kFusedLd + kFusedChunk + ttid] = 0.0;
              }
            }
            team_sync(team);
            const int ns = m4 >> 2;
            const double* frag = tile + (lane >> 2) * kFusedLd + (lane & 3);
            for (int s = twarp; s < 2 * ns; s += kTeamWarps) {
              const int k0 = s < ns ? 4 * s : kFusedChunk + 4 * (s - ns);
              const double a0 = frag[k0], a1 = frag[8 * kFusedLd + k0];
              dmma_m8n8k4(acc[0][0], acc[0][1], a0, a0);
              dmma_m8n8k4(acc[1][0], acc[1][1], a1, a0);
              dmma_m8n8k4(acc[2][0], acc[2][1], a1, a1);
            }
          }
          {
            double* rw = red + twarp * kFusedRed + (lane >> 2) * 8 + 2 * (lane & 3);
#pragma unroll
            for (int b = 0; b < 3; ++b) { rw[b * 64] = acc[b][0]; rw[b * 64 + 1] = acc[b][1]; }
          }
          team_sync(team);
          for (int e = ttid; e < kFusedRed; e += kTeamThreads) {
            const int b = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
            const int i = (b == 0 ? 0 : 8) + rr, j = (b == 2 ? 8 : 0) + cc;
            if (j > i) continue;
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kTeamWarps; ++w) v += red[w * kFusedRed + e];
            v *= a.dp.visual_mult;
            Gm[i * 16 + j] = v;
            Gm[j * 16 + i] = v;
          }
          team_sync(team);
          const double* Grf = Gm + (6 + K) * 16;
          const int nsym = NG * (NG + 1) / 2;
          const int n_out = 36 + 6 + 6 * NG + nsym + NG;
          for (int e = ttid; e < n_out; e += kTeamThreads) {
            int o = e;
            if (o < 36) { smB[o] += Gm[(o / 6) * 16 + (o % 6)]; continue; }
            o -= 36;
            if (o < 6) { smg[o] += Grf[o]; continue; }
            o -= 6;
            if (o < 6 * NG) {
              const int j = o / NG, p = o - j * NG;
              Ef[j * G + ci.goff + p] = p < 6 ? mask[p] * times_A(Gm + j * 16, p, Rc) : Gm[p * 16 + j];
              continue;
            }
            o -= 6 * NG;
            if (o < nsym) {
              int p = static_cast<int>((sqrt(8.0 * o + 1.0) - 1.0) * 0.5);
              while ((p + 1) * (p + 2) / 2 <= o) ++p;
              while (p * (p + 1) / 2 > o) --p;
              const int q = o - p * (p + 1) / 2;
              double v;
              if (q >= 6) {
                v = Gm[p * 16 + q];
              } else if (p >= 6) {
                v = mask[q] * times_A(Gm + p * 16, q, Rc);
              } else {
                if (p < 3) {
                  v = -times_A(Gm + (3 + p) * 16, q, Rc);
                } else {
                  const double* r = Rc + 3 * (p - 3);
                  v = -(r[0] * times_A(Gm, q, Rc) + r[1] * times_A(Gm + 16, q, Rc) + r[2] * times_A(Gm + 32, q, Rc));
                }
                v *= mask[p] * mask[q];
              }
              Ct[(ci.goff + p) * G + ci.goff + q] += v;  // lower triangle; the same thread owns the entry every frame
              continue;
            }
            o -= nsym;
            Ct[G * G + ci.goff + o] += o < 6 ? mask[o] * times_A(Grf, o, Rc) : Grf[o];
          }
        }
        team_sync(team);
        double* Bf = bt.B + static_cast<int64_t>(f) * FD * FD;
        for (int q = ttid; q < 36; q += kTeamThreads) Bf[q] = smB[q];
        if (ttid < 6) {
          const double gv = smg[ttid];
          bt.gf[static_cast<int64_t>(f) * FD + ttid] = gv;
          gf2 += gv * gv;
          gfmax = fmax(gfmax, fabs(gv));
        }
      }
    }
    // CTA partials: scalars, then the teams' global blocks in fixed order
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cost += __shfl_xor_sync(0xffffffffu, cost, o);
        gf2 += __shfl_xor_sync(0xffffffffu, gf2, o);
        gfmax = fmax(gfmax, __shfl_xor_sync(0xffffffffu, gfmax, o));
      }
      __syncthreads();  // S-phase users of wred are long done; teams are done with Cteam
      if (lane == 0) {
        double* w = teams + warp * 8;  // team areas are free again
        w[0] = cost; w[1] = gf2; w[2] = gfmax;
        w[3] = ustat[0]; w[4] = ustat[1]; w[5] = ustat[2]; w[6] = ustat[3];
      }
      __syncthreads();
      double* out = a.partC + static_cast<int64_t>(bid) * PS;
      for (int e = tid; e < NS; e += nthreads) {
        double v = 0.0;
        for (int t = 0; t < n_teams; ++t) v += Cteam[static_cast<size_t>(t) * NS + e];
        out[e] = v;
      }
      if (tid == 0) {
        double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < nwarps; ++w) {
          const double* p = teams + w * 8;
          t[0] += p[0]; t[1] += p[1]; t[2] = fmax(t[2], p[2]);
          t[3] += p[3]; t[4] += p[4]; t[5] += p[5]; t[6] += p[6];
        }
        out[NS + kPCost] = t[0]; out[NS + kPGf2] = t[1]; out[NS + kPGfMax] = t[2];
        out[NS + kPDotG] = t[3]; out[NS + kPDotD] = t[4]; out[NS + kPStep2] = t[5]; out[NS + kPXnorm2] = t[6];
      }
    }
    mark(kProfB);
    grid.sync();
    mark(kProfSyncB);

    // ------------------------------------------------------------ D: global block of the trial point, decision
    {
      double* Ctrial = Cacc + (1 - cur) * NS;
      mega_reduce_stage1(a.partC, PS, nb, PS, a.totC, NS + kPGfMax, NS + kPNotPD);
      grid.sync();
      for (int e = tid; e < NS; e += nthreads) Swork[e] = __ldcg(a.totC + e);
      __syncthreads();
      for (int e = tid; e < NS; e += nthreads) {
        if (e < G * G) {
          const int r = e / G, c = e - r * G;
          Ctrial[e] = Swork[c > r ? c * G + r : e];
        } else {
          Ctrial[e] = Swork[e];
        }
      }
      if (warp == 0) {
        double v[7];
        v[0] = __ldcg(a.totC + NS + kPCost); v[1] = __ldcg(a.totC + NS + kPGf2); v[2] = __ldcg(a.totC + NS + kPDotG);
        v[3] = __ldcg(a.totC + NS + kPDotD); v[4] = __ldcg(a.totC + NS + kPStep2); v[5] = __ldcg(a.totC + NS + kPXnorm2);
        v[6] = __ldcg(a.totC + NS + kPGfMax);
        double g2 = 0.0, gm = 0.0;
        for (int q = lane; q < G; q += 32) {
          const double gv = Swork[G * G + q];
          g2 += gv * gv;
          gm = fmax(gm, fabs(gv));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          g2 += __shfl_xor_sync(0xffffffffu, g2, o);
          gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        }
        if (lane == 0) {
          sc[kScCost] = v[0];
          sc[kScGnorm2] = v[1] + g2;
          sc[kScGmax] = fmax(v[6], gm);
          sc[kScDotG] = v[2];
          sc[kScDotD] = v[3];
          sc[kScStep2] = v[4];
          sc[kScXnorm2] = v[5];
          decide_step(ctl, sc, 1);
          if (bid == 0) {
            *a.ctl = *ctl;
            for (int q = 0; q < 8; ++q) a.scalars[q] = sc[q];
            *bt.cost = v[0];
          }
        }
      }
      __syncthreads();
    }
    mark(kProfD);
  }
  // hand the accepted point's global block back to the multi-launch engine / inspection hooks
  __syncthreads();
  if (bid == 0) {
    const Blocks& bf = a.blk[ctl->cur];
    const double* Cf = Cacc + ctl->cur * NS;
    for (int e = tid; e < NS; e += nthreads) {
      if (e < G * G) bf.C[e] = Cf[e];
      else bf.gc[e - G * G] = Cf[e];
    }
  }
}

}  // namespace vc
