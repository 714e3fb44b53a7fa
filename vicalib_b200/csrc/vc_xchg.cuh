// In-kernel all-gather between the frame shards of one node (one process per GPU, NVLink / NVSwitch peer stores).
//
// Every rank owns a totals buffer that all ranks of the node have mapped (CUDA IPC, vcgpu_comm_init).  An entry is
// published as two 64-bit words {low half | tag}, {high half | tag}: an 8-byte store is single-copy atomic, so a
// reader that sees both tags has the value — no fence, no flag, no barrier between publishing and reading.  tag = the
// number of the exchange (the host counts launches; every rank enqueues the same sequence), parity = tag & 1: a slot
// is rewritten two exchanges later, and nobody can be two exchanges ahead of a peer it still needs an entry from.
//
// The inertial persistent kernels use it the "one reader per entry" way: entry e of the result is read (polled) and
// summed over the ranks — in rank order, so every rank computes bit-identical totals — by ONE thread of the grid,
// written to local memory, and a grid barrier later everybody reads the local copy.  (The vision kernel's older
// scheme, vc_mega.cuh mega_total, has every CTA poll every entry of every rank.)
#pragma once
#include "vc_mega.cuh"

namespace vc {

// 64-bit word offsets inside the totals buffer: the vision kernel's regions end at kXchgCtlOff + a few words
constexpr size_t kXchgImuDenseOff = 524288;            // [2 parities][ranks][stride][2]
constexpr size_t kXchgImuEvalOff = 2 * 1048576;        // [2 parities][ranks][stride][2]
constexpr size_t kXchgWords = kXchgBytes / 8;

struct Xchg {
  int rank, nranks;
  unsigned long long* buf[kMaxRanks];   // totals buffer of every rank (peer pointers; [rank] is local)
  size_t off;                           // region (64-bit words)
  int stride;                           // entries per (parity, rank) slot
  unsigned tag;                         // exchange number (> 0)
  Ctl* ctl;                             // a reader that times out sets done = kMegaCommFailed
};

// The peer pointers move to shared memory once per launch: indexing the kernel-parameter array with a run-time rank
// would make the compiler copy the whole parameter block to local memory.
__device__ __forceinline__ void xchg_stage(const Xchg& x, unsigned long long** bufs_s) {
  if (threadIdx.x < kMaxRanks) {
    unsigned long long* p = x.buf[0];
#pragma unroll
    for (int r = 1; r < kMaxRanks; ++r)
      if (threadIdx.x == r) p = x.buf[r];
    bufs_s[threadIdx.x] = p;
  }
}

// entry e of this rank's slot -> every rank's buffer
__device__ __forceinline__ void xchg_put(const Xchg& x, unsigned long long* const* bufs, int e, double v) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
  const unsigned long long hi_tag = static_cast<unsigned long long>(x.tag) << 32;
  const unsigned long long w0 = (bits & 0xffffffffull) | hi_tag, w1 = (bits >> 32) | hi_tag;
  const size_t w = x.off + 2 * (static_cast<size_t>((x.tag & 1u) * x.nranks + x.rank) * x.stride + e);
  for (int r = 0; r < x.nranks; ++r) {
    const int rr = (x.rank + r) % x.nranks;  // start with the local copy, then walk the peers (spreads the links)
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(bufs[rr] + w), "l"(w0), "l"(w1) : "memory");
  }
}

// entry e of rank r's slot, from the local buffer; spins until the words carry this exchange's tag
__device__ __forceinline__ double xchg_get(const Xchg& x, unsigned long long* const* bufs, int r, int e) {
  const unsigned long long* w = bufs[x.rank] + x.off + 2 * (static_cast<size_t>((x.tag & 1u) * x.nranks + r) * x.stride + e);
  unsigned long long w0, w1;
  unsigned polls = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
    if (static_cast<unsigned>(w0 >> 32) == x.tag && static_cast<unsigned>(w1 >> 32) == x.tag) break;
    __nanosleep(polls < 8 ? 100 : 400);
    if ((++polls & 1023u) == 0) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull || *reinterpret_cast<volatile int*>(&x.ctl->done) == kMegaCommFailed) {
        *reinterpret_cast<volatile int*>(&x.ctl->done) = kMegaCommFailed;  // every later launch returns at once
        return 0.0;
      }
    }
  }
  return __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
}

}  // namespace vc
