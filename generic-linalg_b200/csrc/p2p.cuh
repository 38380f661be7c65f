// p2p.cuh -- device side of the NVLink peer-memory communication (see comm.cu).
//
// Every rank maps every peer's arena; a Mailbox sits at offset 0 of each.  A kernel whose last block
// has just finished this rank's partial sums can complete the sum over ranks ITSELF: lane g of one warp
// stores the values into slot [seq % 4][rank] of peer g's mailbox and polls the words rank g sent here,
// then the warp adds the G contributions in rank order.  All ranks obtain the bit-identical result; no
// library call, no extra kernel.  The words travel NCCL-LL style: every 8-byte store carries 32 bits of
// payload and a 32-bit tag derived from the sequence number, so data and flag arrive in one atomic
// store -- no fence, no second round trip, and the G peers are served in parallel by G lanes.
#pragma once
#include <cuda_runtime.h>

namespace glb {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_RED_SLOTS = 4;
constexpr int P2P_RED_WIDTH = 40;  // doubles per reduction (multi_dot of 16 complex vectors + slack)

struct Mailbox {  // at offset 0 of every arena
  // [slot][sending rank][2*t], [2*t+1] : low / high half of value t, each tagged in the upper 32 bits
  unsigned long long ll[P2P_RED_SLOTS][P2P_MAX_RANKS][2 * P2P_RED_WIDTH];
};

struct P2PRed {  // by-value kernel argument; seq == 0 means "not used"
  Mailbox* mb[P2P_MAX_RANKS];
  int rank, nranks;
  unsigned long long seq;
  long long budget;  // spin budget in clock64() ticks before the kernel gives up on a peer; 0 = wait for ever
};

struct HaloWait {  // by-value kernel argument; seq == 0 means "nothing to wait for"
  const unsigned long long* flag_lo;
  const unsigned long long* flag_hi;
  unsigned long long seq;
  long long budget;  // as P2PRed::budget
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Bounded spin: a peer that never shows up (crashed rank) must surface as a CUDA error, not as a hung GPU.  The
// budget comes from the communicator (GLB_P2P_TIMEOUT_S, default 600 s, 0 = never): ranks are only loosely coupled,
// so benign skew -- host-side set-up, file IO, a debugger -- must not be mistaken for a dead peer.
__device__ __forceinline__ void spin_until(const unsigned long long* flag, unsigned long long seq, long long budget) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq) {
    __nanosleep(64);
    if (budget > 0 && clock64() - t0 > budget) __trap();
  }
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Sum vals[0..n) over all ranks, in place.  Call from all 32 lanes of exactly ONE warp per rank; vals is
// read from lane 0 and valid in every lane on return.  `base` offsets the word index when one reduction
// (one seq) is fed through several calls.
__device__ __forceinline__ void p2p_allreduce_warp(const P2PRed& pr, double* vals, int n, int base = 0) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int slot = (int)(pr.seq % P2P_RED_SLOTS);
  const unsigned long long tag = ((pr.seq % 0xffffffffull) + 1ull) << 32;  // never 0: fresh mailboxes are zero
  Mailbox* mine = pr.mb[pr.rank];
  Mailbox* peer = pr.mb[lane < pr.nranks ? lane : pr.rank];
  for (int t = 0; t < n; t++) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(__shfl_sync(full, vals[t], 0));
    if (lane < pr.nranks) {
      st_relaxed_sys(&peer->ll[slot][pr.rank][2 * (base + t)], tag | (bits & 0xffffffffull));
      st_relaxed_sys(&peer->ll[slot][pr.rank][2 * (base + t) + 1], tag | (bits >> 32));
    }
  }
  for (int t = 0; t < n; t++) {
    double got = 0.0;
    if (lane < pr.nranks) {
      const unsigned long long* w = &mine->ll[slot][lane][2 * (base + t)];
      const long long t0 = clock64();
      unsigned long long lo, hi;
      for (;;) {
        lo = ld_relaxed_sys(w);
        hi = ld_relaxed_sys(w + 1);
        if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
        if (pr.budget > 0 && clock64() - t0 > pr.budget) __trap();  // a peer that never shows up must not hang the GPU
      }
      got = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    }
    double s = 0.0;
    for (int g = 0; g < pr.nranks; g++) s += __shfl_sync(full, got, g);  // rank order: same bits everywhere
    vals[t] = s;
  }
}
// block-wide wait for the neighbours' ghost rows (thread 0 spins, everybody syncs)
__device__ __forceinline__ void halo_wait_block(const HaloWait& hw) {
  if (hw.seq != 0) {
    if (threadIdx.x == 0) {
      spin_until(hw.flag_lo, hw.seq, hw.budget);
      spin_until(hw.flag_hi, hw.seq, hw.budget);
    }
    __syncthreads();
  }
}
#endif

}  // namespace glb
