// Persistent trust-region kernel for the vision-only problem (one GPU, or frame shards on several).
//
// One cooperative launch runs up to n_iters Levenberg-Marquardt iterations; one CTA per SM, every
// frame owned by the same WARP of the same CTA for the whole solve (frame f -> CTA f mod grid, warp
// (f / grid) mod warps).  An iteration is a chain of per-frame warp phases and small every-CTA phases,
// separated by grid barriers (~1.5 us) instead of launches:
//
//   S  per frame (warp):  L = chol(B_f + D_f),  X_f = (B_f + D_f)^-1 [E_f | g_f],  CTA partial of E^T X
//   -- grid barrier; every CTA leaves here, together, if CTA 0 published "stop" (or a peer timed out)
//   R1 entry e of the Schur sum is added up over the CTAs by one warp of CTA (e mod grid) and published:
//      one GPU: plain totals + grid barrier; shards: tagged words into every rank's buffer, readers poll
//   G  every CTA, redundantly: reduced system C + D - sum E^T X, Cholesky, globals' step
//   U  per frame (warp):  back-substitution, x (+) step, step statistics; trial camera states
//   B  per frame (warp): fused evaluate + Gram build at the trial point, 32 corners at a time through the
//      warp's own shared-memory slab and FP64 DMMA (accumulators stay in registers across slabs; no block
//      barrier anywhere in the phase); the camera blocks are accumulated per warp, packed per camera
//   -- grid barrier
//   R2 packed camera blocks / scalars added up and published, as R1
//   D  every CTA, redundantly: expands the packed blocks, takes the accept-reject decision of Ceres'
//      TrustRegionMinimizer (decide_step) on its own copy of Ctl
//
// Every CTA (of every rank) reads bit-identical totals, so all take the same decision and the loop needs
// no broadcast; all sums run in a fixed order (deterministic).  Replaces the body of ceres::Solve
// (vicalibrator.h:956) for the staged vision solves; the inertial path keeps the multi-launch engine
// (vc_engine.inl).
#pragma once
#include <cooperative_groups.h>

#include "vc_fused.cuh"
#include "vc_kernels.cuh"

namespace vc {
namespace cg = cooperative_groups;

constexpr int kMegaMaxWarps = 16;
constexpr int kMegaMaxThreads = 32 * kMegaMaxWarps;
constexpr int kSlabLd = 64 + 4;                        // 32 corners x 2 residual rows (+4: conflict-free fragments)
constexpr int kSlabDoubles = kFusedCols * kSlabLd;     // one warp's tile
constexpr int kWarpDoubles = kSlabDoubles + 48;        // + frame block (36) + frame gradient (6), padded
constexpr int kGtabMax = 256;
constexpr int kMegaPartExtra = 8;  // scalars appended to each CTA's partial slot
// Totals buffer: every rank owns one (32 MiB, the vision kernel uses the first 4); in a frame-sharded run (one process per GPU) all ranks of the
// node map all of them (CUDA IPC).  A total is published as two 64-bit words {low half | tag}, {high half | tag}
// (tag = exchange number): an 8-byte store is single-copy atomic, so a reader that sees both tags has the value —
// no fence, no flag, no barrier between publishing and reading.  Layout in 64-bit words: Schur totals
// [2 parities][ranks][NS+8][2] at 0, packed camera blocks [2][ranks][NP+8][2] at kXchgCOff, control words at
// kXchgCtlOff: exchange counters (S, C), exit word (bit 0: CTA 0 decided to stop, bit 1: a reader timed out).
constexpr int kMaxRanks = 8;
constexpr size_t kXchgBytes = 32u << 20;  // the inertial kernels' regions follow the vision kernel's (vc_xchg.cuh)
constexpr int kXchgCOff = 409600;
constexpr int kXchgCtlOff = 442368;
// sharded runs: the totals summed over the ranks, as plain doubles, for this rank's CTAs (one reader per entry polls
// the tagged words and writes here; a grid barrier later everybody reads the plain copy)
constexpr int kXchgPlainS = kXchgCtlOff + 64;
constexpr int kXchgPlainC = kXchgPlainS + 32768;
constexpr int kMegaCommFailed = 1000;  // Ctl::done value when a peer never showed up
static_assert(kSlabLd % 16 == 4, "fragment loads need ld == 4 (mod 16)");
enum { kPCost = 0, kPGf2, kPDotG, kPDotD, kPStep2, kPXnorm2, kPGfMax, kPNotPD };
enum { kProfS = 0, kProfG, kProfU, kProfB, kProfD, kProfSync, kProfCount };

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
  double* partS;         // [grid][NS + 8]   Schur partial (lower triangle) | flags
  double* partC;         // [grid][n_cams*kCgStride + 8]   packed camera blocks | scalars
  double* delta;         // scaled step [nf*6 + G]
  double* scalars;       // kSc* of the last evaluated point (for the host)
  int n_iters;
  int n_warps;
  int rank, nranks;                      // frame shards (nranks = 1: a single GPU)
  unsigned long long* xbuf[kMaxRanks];   // totals buffer of every rank (peer pointers; [rank] is local)
  unsigned long long* prof;  // [kProfCount] ns per phase (CTA 0), or null
};

__host__ __device__ inline size_t mega_smem_doubles(int G, int n_cams, int n_warps) {
  const size_t NS = static_cast<size_t>(G) * G + G;
  return static_cast<size_t>(n_warps) * (kWarpDoubles + n_cams * kCgStride) + 3 * NS + G + kMaxCams * (kCamStateStride + 9) +
         kScCount + 8 * kMegaMaxWarps + n_cams * kCgStride + kMegaPartExtra + sizeof(Ctl) / sizeof(double) + 8;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Single GPU: entry e is added up over the CTAs by one warp of CTA (e mod grid) into tot[e]; a grid barrier
// later every CTA reads tot[] (measured 2 us per iteration faster than polling tagged words when nobody is remote).
// (bid of nb: the CTAs sharing the work — the whole grid unless the caller says otherwise)
__device__ inline void mega_reduce_stage1(const double* part, int stride, int nparts, int n, double* tot, int max_a, int max_b,
                                          int bid = blockIdx.x, int nb = gridDim.x) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int e = bid + warp * nb; e < n; e += nwarps * nb) {
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

// ---- the two reductions of an iteration: CTA partials -> totals (over the CTAs and, sharded, over the GPUs) ----
// Publish: entry e is added up over this GPU's CTAs by one warp of CTA (e mod grid) — fixed order — and stored,
// tagged, into slot [parity][my rank][e] of EVERY rank's totals buffer (lane r -> rank r; NVLink peer stores).
__device__ inline void mega_publish(const double* part, int stride, int nparts, int n, const MegaArgs& a, int off, int parity,
                                    unsigned tag, int max_a, int max_b) {
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
    if (lane < a.nranks) {
      const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(s0));
      const unsigned long long hi_tag = static_cast<unsigned long long>(tag) << 32;
      unsigned long long* w = a.xbuf[lane] + off + 2 * (static_cast<size_t>(parity * a.nranks + a.rank) * stride + e);
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"((bits & 0xffffffffull) | hi_tag),
                   "l"((bits >> 32) | hi_tag)
                   : "memory");
    }
  }
}
// Read: the total of entry e over the ranks, in rank order; spins until every rank's words carry `tag`.
// A reader that waits longer than 2 s sets the abort bit of the exit word and gives up (the value is then junk;
// every CTA leaves the loop together at the next grid barrier).
__device__ inline double mega_total(const MegaArgs& a, int off, int stride, int parity, unsigned tag, int e, bool is_max) {
  unsigned long long* ctlw = a.xbuf[a.rank] + kXchgCtlOff;
  const unsigned long long* w = a.xbuf[a.rank] + off + 2 * (static_cast<size_t>(parity) * a.nranks * stride + e);
  double s = 0.0;
  for (int r = 0; r < a.nranks; ++r, w += 2 * static_cast<size_t>(stride)) {
    unsigned long long w0, w1;
    unsigned polls = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
      if (static_cast<unsigned>(w0 >> 32) == tag && static_cast<unsigned>(w1 >> 32) == tag) break;
      __nanosleep(polls < 8 ? 100 : 400);  // thousands of threads poll: keep the L2 free for the publishers
      if ((++polls & 1023u) == 0) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > 2000000000ull || (*reinterpret_cast<volatile unsigned long long*>(ctlw + 2) & 2ull)) {
          atomicOr(ctlw + 2, 2ull);
          return 0.0;
        }
      }
    }
    const double v = __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
    s = is_max ? fmax(s, v) : s + v;
  }
  return s;
}

// Reduced-system solve, one warp.  S: lower triangle [G][G] row-major followed by the right-hand side [G];
// the solution overwrites the right-hand side.  The right-hand side is carried through the factorisation as an
// extra row (so L y = rhs comes for free) and the pivots are kept as reciprocals.  Rows are spread over the
// lanes, column updates batched by four.  (Register-resident variants — lane r owning row r, columns broadcast
// with shuffles — were measured twice and were 1-2 us slower in this phase: the row array ends up in local memory.)
__device__ inline void mega_chol_solve_smem(double* S, int G, int lane, double* invd, int* bad) {
  double* rhs = S + G * G;
  for (int j = 0; j < G; ++j) {
    double d = S[j * G + j];
    if (!(d > 0.0)) { if (lane == 0) *bad = 1; d = 1.0; }
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

template <int MODEL>
__device__ __forceinline__ double mega_eval(const double* T, const double* cam, const double* Rc, const double* mask, V3 pw, double pcu,
                                            double pcv, double mult, double* slab, int lane) {
  return eval_obs_to_tile<MODEL, kSlabLd>(T, cam, Rc, mask, pw, pcu, pcv, mult, slab, lane, 32 + lane);
}

__global__ void __launch_bounds__(kMegaMaxThreads, 1) lm_mega_kernel(MegaArgs a) {
  extern __shared__ double smem[];
  cg::grid_group grid = cg::this_grid();
  constexpr int FD = 6;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int bid = blockIdx.x, nb = gridDim.x;
  const int n_cams = a.dp.n_cams;
  const int G = a.dp.G, M = G + 1, NS = G * G + G, PS = NS + kMegaPartExtra, nf = a.dp.n_frames;
  const int NP = n_cams * kCgStride, PC = NP + kMegaPartExtra;  // packed camera blocks
  const int64_t nfp = static_cast<int64_t>(nf) * FD;
  const int nk = bid < nf ? (nf - bid + nb - 1) / nb : 0;  // frames of this CTA: f = bid + k * nb, warp k mod nwarps

  double* wmem = smem;                                               // [nwarps][kWarpDoubles]; also phase scratch
  double* Cw = wmem + static_cast<size_t>(nwarps) * kWarpDoubles;    // [nwarps][NP] packed camera blocks per warp
  double* Cacc = Cw + static_cast<size_t>(nwarps) * NP;              // [2][NS]  C | gc of the two points
  double* Swork = Cacc + 2 * NS;                                     // [NS] Schur accumulator / reduced system
  double* dg = Swork + NS;                                           // [G] globals' step (scaled)
  double* smCam = dg + G;                                            // [kMaxCams][17] trial camera states
  double* smRc = smCam + kMaxCams * kCamStateStride;                 // [kMaxCams][9]
  double* sc = smRc + kMaxCams * 9;                                  // [kScCount]
  double* wred = sc + kScCount;                                      // [kMegaMaxWarps][8] per-warp scalars
  double* totc = wred + 8 * kMegaMaxWarps;                           // [NP + 8] packed camera blocks / scalars, grid (and rank) totals
  Ctl* ctl = reinterpret_cast<Ctl*>(totc + NP + kMegaPartExtra);
  __shared__ int bad;
  __shared__ int2 gtab[kGtabMax];          // (start, count) of the observations of this CTA's (frame slot, camera) pairs
  __shared__ unsigned char tri_lut[128];   // packed lower-triangle index -> (row << 4 | col)
  const double* scg = a.scale + nfp;
  unsigned long long* ctlw = a.xbuf[a.rank] + kXchgCtlOff;  // [0] S exchanges, [1] C exchanges, [2] exit word
  __shared__ unsigned long long exit_s;

  if (tid == 0) *ctl = *a.ctl;
  const bool use_gtab = nk * n_cams <= kGtabMax;
  if (use_gtab)  // frame ownership is static: look the groups up once per launch
    for (int e = tid; e < nk * n_cams; e += nthreads) {
      const int k = e / n_cams, c = e - k * n_cams;
      const int g = a.group_of[c * nf + bid + k * nb];
      gtab[e] = g < 0 ? make_int2(0, 0) : make_int2(a.grp_start[g], a.grp_count[g]);
    }
  for (int e = tid; e < 105; e += nthreads) {
    int p = 0;
    while ((p + 1) * (p + 2) / 2 <= e) ++p;
    tri_lut[e] = static_cast<unsigned char>((p << 4) | (e - p * (p + 1) / 2));
  }
  __syncthreads();
  if (ctl->done) return;
  if (bid == 0 && tid == 0 && !(ctlw[2] & 2ull)) {  // the stop bit of the previous launch; the abort bit is sticky
    ctlw[2] = 0;
    __threadfence();
  }
  {
    const Blocks& b0 = a.blk[ctl->cur];
    double* C0 = Cacc + ctl->cur * NS;
    for (int e = tid; e < NS; e += nthreads) C0[e] = e < G * G ? b0.C[e] : b0.gc[e - G * G];
  }
  // (start, count) of the observations of (frame slot k, camera c); count 0: the camera does not see the frame
  auto lookup = [&](int k, int c) -> int2 {
    if (use_gtab) return gtab[k * n_cams + c];
    const int g = a.group_of[c * nf + bid + k * nb];
    return g < 0 ? make_int2(0, 0) : make_int2(a.grp_start[g], a.grp_count[g]);
  };
  const bool sharded = a.nranks > 1;
  unsigned epS = static_cast<unsigned>(ctlw[0]), epC = static_cast<unsigned>(ctlw[1]);  // same on every CTA of every rank
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
    // A CTA never leaves the loop on its own: it idles through phase S and leaves with everybody else right
    // after the next grid barrier, on the word CTA 0 published (so a junk total can never split the grid).
    const bool skip = ctl->done != 0;
    const int cur = ctl->cur;
    const Blocks& bc = a.blk[cur];
    const Blocks& bt = a.blk[1 - cur];
    const double* x_cur = a.state[cur];
    double* x_new = a.state[1 - cur];
    const double rinv = 1.0 / ctl->radius;
    const double* Ccur = Cacc + cur * NS;
    double notpd = 0.0;

    // ------------------------------------------------------------ S: per-frame solves + Schur partial
    if (!skip) {
      double* Sacc = Swork;
      for (int k = tid; k < NS; k += nthreads) Sacc[k] = 0.0;
      for (int base = 0; base < nk; base += nwarps) {
        const int k = base + warp;
        __syncthreads();
        if (k < nk) {
          const int f = bid + k * nb;
          double* Esw = wmem + static_cast<size_t>(warp) * kWarpDoubles;  // [FD][M] scaled [E | g], then X [FD][M]
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
            const double inv = rsqrt(d);
            L[j][j] = d * inv;
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
            const double* Ew = wmem + static_cast<size_t>(w) * kWarpDoubles;
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
      const unsigned any = __ballot_sync(0xffffffffu, notpd > 0.0);
      if (lane == 0) wred[8 * warp] = any ? 1.0 : 0.0;
      __syncthreads();
      if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < nwarps; ++w) v = fmax(v, wred[8 * w]);
        out[NS + kPNotPD] = v;
      }
    }
    mark(kProfS);
    grid.sync();
    if (tid == 0) exit_s = *reinterpret_cast<volatile unsigned long long*>(ctlw + 2);
    __syncthreads();
    if (exit_s) break;
    mark(kProfSync);

    // ------------------------------------------------------------ G: reduced system, every CTA the same
    {
      double* S = Swork;  // [G*G] lower triangle, then rhs [G]
      ++epS;
      const int parS = static_cast<int>(epS & 1);
      // plain totals: in the exchange buffer itself on one GPU, next to the tagged words in a sharded run
      double* totS = reinterpret_cast<double*>(a.xbuf[a.rank] + (sharded ? kXchgPlainS : 0));
      if (sharded) {
        // every rank's CTA sums go to every rank (tagged words over NVLink); ONE thread of this grid per entry waits for
        // them and adds them up in rank order (every rank computes the same bits)
        mega_publish(a.partS, PS, nb, PS, a, 0, parS, epS, NS + kPNotPD, -1);
        for (int e = bid * nthreads + tid; e < PS; e += nb * nthreads) totS[e] = mega_total(a, 0, PS, parS, epS, e, e == NS + kPNotPD);
      } else {
        mega_reduce_stage1(a.partS, PS, nb, PS, totS, NS + kPNotPD, -1);
      }
      mark(kProfG);
      grid.sync();
      mark(kProfSync);
      for (int e = tid; e < NS; e += nthreads) {
        if (e < G * G && e % G > e / G) continue;  // lower triangle only
        const double p = __ldcg(totS + e);
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
      if (tid == 0) {
        sc[kScNotPD] = __ldcg(totS + NS + kPNotPD);
        bad = 0;
      }
      __syncthreads();
      if (warp == 0) {
        mega_chol_solve_smem(S, G, lane, dg, &bad);
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
      if (warp == nwarps - 1) {  // trial camera states (every CTA needs them; the last warp is the one most likely
                                 // to own no frame); CTA 0 also stores the globals
        if (lane < n_cams) {
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
        if (bid == 0 && a.rank == 0 && lane == 0) {  // the globals' share of the step statistics, counted once
          for (int c = 0; c < n_cams; ++c) {
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
      for (int e = tid; e < nwarps * NP; e += nthreads) Cw[e] = 0.0;
      __syncthreads();  // trial camera states and the trial poses are visible to every warp
    }
    mark(kProfU);

    // ------------------------------------------------------------ B: fused evaluate + build at the trial point
    // No block barrier in this phase: each warp runs its frames on its own slab, at its own pace.
    double cost = 0.0, gf2 = 0.0, gfmax = 0.0;
    {
      double* slab = wmem + static_cast<size_t>(warp) * kWarpDoubles;  // [16][kSlabLd]; reused as the 16 x 16 Gram matrix
      double* smB = slab + kSlabDoubles;                               // [36] | [6]
      double* smg = smB + 36;
      double* Cmine = Cw + static_cast<size_t>(warp) * NP;
      const double* frag = slab + (lane >> 2) * kSlabLd + (lane & 3);
      for (int k = warp; k < nk; k += nwarps) {
        const int f = bid + k * nb;
        const double* T = x_new + 7 * static_cast<int64_t>(f);
        __syncwarp();
        for (int q = lane; q < 42; q += 32) smB[q] = 0.0;
        double* Ef = bt.E + static_cast<int64_t>(f) * FD * G;
        bool zero_E = n_cams > 1;  // one camera: the expansion below writes every column
        for (int c = 0; c < n_cams; ++c) {
          const int2 grp = lookup(k, c);
          if (grp.y == 0) continue;
          if (zero_E) {
            for (int q = lane; q < FD * G; q += 32) Ef[q] = 0.0;
            zero_E = false;
          }
          const CamInfo& ci = a.dp.cams[c];
          const int K = ci.K, NG = 6 + K, model = ci.model;
          const int start = grp.x, cnt = grp.y;
          const double* cam = smCam + kCamStateStride * c;
          const double* Rc = smRc + 9 * c;
          const double* mask = a.mask + ci.goff;
          double acc[3][2];
#pragma unroll
          for (int b = 0; b < 3; ++b) acc[b][0] = acc[b][1] = 0.0;
          // software pipeline: the next slab's observation is loaded while this slab goes through the DMMAs
          double nx[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
          if (lane < cnt) {
            const int64_t i = start + lane;
            nx[0] = a.pw[3 * i]; nx[1] = a.pw[3 * i + 1]; nx[2] = a.pw[3 * i + 2]; nx[3] = a.pc[2 * i]; nx[4] = a.pc[2 * i + 1];
          }
          for (int s0 = 0; s0 < cnt; s0 += 32) {
            const int m = min(32, cnt - s0);
            const int m4 = (m + 3) & ~3;
            const V3 pw{nx[0], nx[1], nx[2]};
            const double pcu = nx[3], pcv = nx[4];
            if (s0 + 32 + lane < cnt) {
              const int64_t i = start + s0 + 32 + lane;
              nx[0] = a.pw[3 * i]; nx[1] = a.pw[3 * i + 1]; nx[2] = a.pw[3 * i + 2]; nx[3] = a.pc[2 * i]; nx[4] = a.pc[2 * i + 1];
            }
            __syncwarp();  // the previous slab's fragment loads are done
            if (lane < m) {
              switch (model) {
                case kLinear: cost += mega_eval<kLinear>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, slab, lane); break;
                case kFov: cost += mega_eval<kFov>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, slab, lane); break;
                case kPoly2: cost += mega_eval<kPoly2>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, slab, lane); break;
                case kPoly3: cost += mega_eval<kPoly3>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, slab, lane); break;
                default: cost += mega_eval<kKb4>(T, cam, Rc, mask, pw, pcu, pcv, a.dp.visual_mult, slab, lane); break;
              }
            } else if (lane < m4) {  // zero the padding rows of both halves
#pragma unroll
              for (int q = 0; q < kFusedCols; ++q) {
                slab[q * kSlabLd + lane] = 0.0;
                slab[q * kSlabLd + 32 + lane] = 0.0;
              }
            }
            __syncwarp();
            // SYRK over k: steps [0, m4/4) cover residual row 0, [m4/4, m4/2) residual row 1
            // SYRK over k: steps [0, m4/4) cover residual row 0, [m4/4, m4/2) residual row 1 (two accumulator sets
            // for six independent DMMA chains measured no faster: the phase is not DMMA-latency bound)
            const int ns = m4 >> 2;
            for (int s = 0; s < 2 * ns; ++s) {
              const int k0 = s < ns ? 4 * s : 32 + 4 * (s - ns);
              const double a0 = frag[k0], a1 = frag[8 * kSlabLd + k0];
              dmma_m8n8k4(acc[0][0], acc[0][1], a0, a0);  // (0,0)
              dmma_m8n8k4(acc[1][0], acc[1][1], a1, a0);  // (1,0)
              dmma_m8n8k4(acc[2][0], acc[2][1], a1, a1);  // (1,1)
            }
          }
          // the three accumulator blocks -> symmetric 16 x 16 Gram matrix (in the slab)
          __syncwarp();
          double* Gm = slab;
          {
            const int rr = lane >> 2, cc = 2 * (lane & 3);
            const double vm = a.dp.visual_mult;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const int i = (b == 0 ? 0 : 8) + rr, j = (b == 2 ? 8 : 0) + cc;
              const double v0 = acc[b][0] * vm, v1 = acc[b][1] * vm;
              // diagonal blocks: both triangles are computed (bitwise equal), keep j <= i and mirror
              if (b == 1 || j <= i) { Gm[i * 16 + j] = v0; Gm[j * 16 + i] = v0; }
              if (b == 1 || j + 1 <= i) { Gm[i * 16 + j + 1] = v1; Gm[(j + 1) * 16 + i] = v1; }
            }
          }
          __syncwarp();
          // expand: frame block, frame gradient, E (extrinsic columns through A), the camera's packed global block
          const double* Grf = Gm + (6 + K) * 16;
          const int nsym = NG * (NG + 1) / 2;
          const int n_out = 36 + 6 + 6 * NG + nsym + NG;
          double* Cc = Cmine + c * kCgStride;
          for (int e = lane; e < n_out; e += 32) {
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
              const int p = tri_lut[o] >> 4, q = tri_lut[o] & 15;
              double v;
              if (q >= 6) {
                v = Gm[p * 16 + q];                                        // intrinsics x intrinsics
              } else if (p >= 6) {
                v = mask[q] * times_A(Gm + p * 16, q, Rc);                // intrinsics x extrinsics
              } else {                                                     // extrinsics x extrinsics: (A^T Gff A)[p][q]
                if (p < 3) {
                  v = -times_A(Gm + (3 + p) * 16, q, Rc);
                } else {
                  const double* r = Rc + 3 * (p - 3);
                  v = -(r[0] * times_A(Gm, q, Rc) + r[1] * times_A(Gm + 16, q, Rc) + r[2] * times_A(Gm + 32, q, Rc));
                }
                v *= mask[p] * mask[q];
              }
              Cc[o] += v;  // the same lane owns the entry for every frame of this warp
              continue;
            }
            o -= nsym;
            Cc[105 + o] += o < 6 ? mask[o] * times_A(Grf, o, Rc) : Grf[o];
          }
        }
        if (zero_E)  // no camera saw the frame
          for (int q = lane; q < FD * G; q += 32) Ef[q] = 0.0;
        __syncwarp();
        double* Bf = bt.B + static_cast<int64_t>(f) * FD * FD;
        for (int q = lane; q < 36; q += 32) Bf[q] = smB[q];
        if (lane < 6) {
          const double gv = smg[lane];
          bt.gf[static_cast<int64_t>(f) * FD + lane] = gv;
          gf2 += gv * gv;
          gfmax = fmax(gfmax, fabs(gv));
        }
      }
    }
    // CTA partials: scalars, then the warps' packed camera blocks in fixed order
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cost += __shfl_xor_sync(0xffffffffu, cost, o);
        gf2 += __shfl_xor_sync(0xffffffffu, gf2, o);
        gfmax = fmax(gfmax, __shfl_xor_sync(0xffffffffu, gfmax, o));
      }
      if (lane == 0) {
        double* w = wred + 8 * warp;
        w[0] = cost; w[1] = gf2; w[2] = gfmax;
        w[3] = ustat[0]; w[4] = ustat[1]; w[5] = ustat[2]; w[6] = ustat[3];
      }
      __syncthreads();
      double* out = a.partC + static_cast<int64_t>(bid) * PC;
      for (int e = tid; e < NP; e += nthreads) {
        double v = 0.0;
        for (int w = 0; w < nwarps; ++w) v += Cw[static_cast<size_t>(w) * NP + e];
        out[e] = v;
      }
      if (tid == 0) {
        double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < nwarps; ++w) {
          const double* p = wred + 8 * w;
          t[0] += p[0]; t[1] += p[1]; t[2] = fmax(t[2], p[2]);
          t[3] += p[3]; t[4] += p[4]; t[5] += p[5]; t[6] += p[6];
        }
        out[NP + kPCost] = t[0]; out[NP + kPGf2] = t[1]; out[NP + kPGfMax] = t[2];
        out[NP + kPDotG] = t[3]; out[NP + kPDotD] = t[4]; out[NP + kPStep2] = t[5]; out[NP + kPXnorm2] = t[6];
        out[NP + kPNotPD] = 0.0;
      }
    }
    mark(kProfB);
    grid.sync();
    mark(kProfSync);

    // ------------------------------------------------------------ D: global block of the trial point, decision
    {
      double* Ctrial = Cacc + (1 - cur) * NS;
      ++epC;
      const int parC = static_cast<int>(epC & 1);
      double* totC = reinterpret_cast<double*>(a.xbuf[a.rank] + (sharded ? kXchgPlainC : kXchgCOff));
      if (sharded) {
        mega_publish(a.partC, PC, nb, PC, a, kXchgCOff, parC, epC, NP + kPGfMax, NP + kPNotPD);
        for (int e = bid * nthreads + tid; e < PC; e += nb * nthreads)
          totC[e] = mega_total(a, kXchgCOff, PC, parC, epC, e, e == NP + kPGfMax || e == NP + kPNotPD);
      } else {
        mega_reduce_stage1(a.partC, PC, nb, PC, totC, NP + kPGfMax, NP + kPNotPD);
      }
      mark(kProfD);
      grid.sync();
      mark(kProfSync);
      for (int e = tid; e < PC; e += nthreads) totc[e] = __ldcg(totC + e);
      __syncthreads();
      for (int e = tid; e < NS; e += nthreads) {  // packed per-camera blocks -> dense C | gc
        const int r = e < G * G ? e / G : e - G * G;
        const int cc = e < G * G ? e - r * G : -1;
        double v = 0.0;
        for (int c = 0; c < n_cams; ++c) {
          const CamInfo& ci = a.dp.cams[c];
          const int p = r - ci.goff;
          if (p < 0 || p >= 6 + ci.K) continue;
          if (cc < 0) {
            v = totc[c * kCgStride + 105 + p];
          } else {
            const int q = cc - ci.goff;
            if (q >= 0 && q < 6 + ci.K) {
              const int hi = max(p, q), lo = min(p, q);
              v = totc[c * kCgStride + hi * (hi + 1) / 2 + lo];
            }
          }
        }
        Ctrial[e] = v;
      }
      __syncthreads();
      if (warp == 0) {
        double v[7];
        v[0] = totc[NP + kPCost]; v[1] = totc[NP + kPGf2]; v[2] = totc[NP + kPDotG];
        v[3] = totc[NP + kPDotD]; v[4] = totc[NP + kPStep2]; v[5] = totc[NP + kPXnorm2];
        v[6] = totc[NP + kPGfMax];
        double g2 = 0.0, gm = 0.0;
        for (int q = lane; q < G; q += 32) {
          const double gv = Ctrial[G * G + q];
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
            if (ctl->done) atomicOr(ctlw + 2, 1ull);
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
  if (bid == 0 && tid == 0) {
    ctlw[0] = epS;
    ctlw[1] = epC;
    if (*reinterpret_cast<volatile unsigned long long*>(ctlw + 2) & 2ull) a.ctl->done = kMegaCommFailed;
  }
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
