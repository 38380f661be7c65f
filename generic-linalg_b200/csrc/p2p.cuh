// p2p.cuh -- device side of the NVLink peer-memory communication (see comm.cu).
//
// Every rank maps every peer's arena; a Mailbox sits at offset 0 of each.  A kernel whose last block
// has just finished this rank's partial sums can complete the sum over ranks ITSELF: one warp per peer
// stores the values into slot [seq % 4][rank] of that peer's mailbox and polls the words it sent here,
// then the block adds the G contributions in rank order.  All ranks obtain the bit-identical result; no
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
  int flush;         // 1: a system-scope fence after the sends pushes the posted 8-byte stores out at once (GLB_P2P_FLUSH)
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
// Sum vals[0..n) over all ranks, in place.  Call from EVERY thread of exactly one block per rank (block size a
// multiple of 32); vals is read from thread 0 and valid in thread 0 on return (n <= P2P_RED_WIDTH).
//
// One WARP per peer, one LANE per word: warp w serves ranks w, w + W, ...; its lanes 0 .. 2m-1 store the 2m tagged
// words of this rank's values into that peer's mailbox with ONE store instruction and then poll, again one lane per
// word, the words that peer sent here.  Measured on 8 B200s (tools/p2p_bench.py, profiles/r02_p2p_allreduce.md): a
// single warp that stores to all peers itself -- lane g sending 12 words to rank g one after the other -- takes
// 18 us per sum of six doubles (the stores to different GPUs leave the SM one destination after the other, about one
// NVLink round trip each), this shape 1.9 us (NCCL's all-reduce of the same six doubles: 22 us).
__device__ __forceinline__ void p2p_allreduce_block(const P2PRed& pr, double* vals, int n) {
  __shared__ double s_vals[P2P_RED_WIDTH];
  __shared__ double s_got[P2P_MAX_RANKS][P2P_RED_WIDTH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int slot = (int)(pr.seq % P2P_RED_SLOTS);
  const unsigned long long tag = ((pr.seq % 0xffffffffull) + 1ull) << 32;  // never 0: fresh mailboxes are zero
  Mailbox* mine = pr.mb[pr.rank];
  if (threadIdx.x == 0)
    for (int t = 0; t < n; t++) s_vals[t] = vals[t];
  __syncthreads();
  for (int c = 0; c < n; c += 16) {  // 16 values = 32 words = one warp-wide store
    const int m = (n - c < 16) ? n - c : 16;
    if (lane < 2 * m) {
      const unsigned mask = (m == 16) ? 0xffffffffu : ((1u << (2 * m)) - 1u);
      const unsigned long long bits = (unsigned long long)__double_as_longlong(s_vals[c + lane / 2]);
      const unsigned long long word = tag | ((lane & 1) ? (bits >> 32) : (bits & 0xffffffffull));
      for (int g = warp; g < pr.nranks; g += nwarp) st_relaxed_sys(&pr.mb[g]->ll[slot][pr.rank][2 * c + lane], word);
      for (int g = warp; g < pr.nranks; g += nwarp) {
        const unsigned long long* w = &mine->ll[slot][g][2 * c + lane];
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
          v = ld_relaxed_sys(w);
          if ((v & 0xffffffff00000000ull) == tag) break;
          if (pr.budget > 0 && clock64() - t0 > pr.budget) __trap();  // a peer that never shows up must not hang the GPU
        }
        const unsigned long long other = __shfl_xor_sync(mask, v, 1);
        if (!(lane & 1)) s_got[g][c + lane / 2] = __longlong_as_double((long long)((v & 0xffffffffull) | (other << 32)));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < n) {
    double s = 0.0;
    for (int g = 0; g < pr.nranks; g++) s += s_got[g][threadIdx.x];  // rank order: the same bits on every rank
    s_vals[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int t = 0; t < n; t++) vals[t] = s_vals[t];
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
