// p2p.cuh -- device side of the NVLink peer-memory communication (see comm.cu).
//
// Every rank maps every peer's arena; a Mailbox sits at offset 0 of each.  A kernel whose last block
// has just finished this rank's partial sums can complete the sum over ranks ITSELF: one thread
// stores the values into slot [seq % 4][rank] of every peer's mailbox, releases a sequence flag on
// each, waits for the G flags in its own mailbox and adds the G contributions in rank order.  All
// ranks obtain the bit-identical result; no library call, no extra kernel.
#pragma once
#include <cuda_runtime.h>

namespace glb {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_RED_SLOTS = 4;
constexpr int P2P_RED_WIDTH = 40;  // doubles per reduction (multi_dot of 16 complex vectors + slack)

struct Mailbox {  // at offset 0 of every arena
  double red[P2P_RED_SLOTS][P2P_MAX_RANKS][P2P_RED_WIDTH];
  unsigned long long red_seq[P2P_RED_SLOTS][P2P_MAX_RANKS];
};

struct P2PRed {  // by-value kernel argument; seq == 0 means "not used"
  Mailbox* mb[P2P_MAX_RANKS];
  int rank, nranks;
  unsigned long long seq;
};

struct HaloWait {  // by-value kernel argument; seq == 0 means "nothing to wait for"
  const unsigned long long* flag_lo;
  const unsigned long long* flag_hi;
  unsigned long long seq;
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
// Bounded spin: a peer that never shows up (crashed rank) must surface as a CUDA error, not as a hung GPU.
__device__ __forceinline__ void spin_until(const unsigned long long* flag, unsigned long long seq) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq) {
    __nanosleep(64);
    if (clock64() - t0 > 40000000000LL) __trap();  // ~20 s at 2 GHz
  }
}
// Sum vals[0..n) over all ranks, in place.  Call from exactly ONE thread of ONE block per rank.
__device__ __forceinline__ void p2p_allreduce_thread(const P2PRed& pr, double* vals, int n) {
  const int slot = (int)(pr.seq % P2P_RED_SLOTS);
  for (int g = 0; g < pr.nranks; g++)
    for (int t = 0; t < n; t++) pr.mb[g]->red[slot][pr.rank][t] = vals[t];
  __threadfence_system();
  for (int g = 0; g < pr.nranks; g++) st_release_sys(&pr.mb[g]->red_seq[slot][pr.rank], pr.seq);
  Mailbox* mine = pr.mb[pr.rank];
  for (int g = 0; g < pr.nranks; g++) spin_until(&mine->red_seq[slot][g], pr.seq);
  for (int t = 0; t < n; t++) {
    double s = 0.0;
    for (int g = 0; g < pr.nranks; g++) s += __ldcv(&mine->red[slot][g][t]);  // rank order: same bits everywhere
    vals[t] = s;
  }
}
// block-wide wait for the neighbours' ghost rows (thread 0 spins, everybody syncs)
__device__ __forceinline__ void halo_wait_block(const HaloWait& hw) {
  if (hw.seq != 0) {
    if (threadIdx.x == 0) {
      spin_until(hw.flag_lo, hw.seq);
      spin_until(hw.flag_hi, hw.seq);
    }
    __syncthreads();
  }
}
#endif

}  // namespace glb
