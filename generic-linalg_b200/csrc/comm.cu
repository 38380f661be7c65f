// comm.cu -- y-slab communicator: one process per GPU, ring of slabs (periodic lattice).
//
//  * halo exchange: the first / last `rows` of a slab-local vector go to the neighbouring ranks'
//    ghost buffers (2 x X x nc x 16 B per apply: 128 KiB per direction at 8192).
//  * reductions: k doubles summed over ranks.
// NCCL is bound lazily (dlopen) so that a single-GPU process never needs it; when torch is loaded
// first its bundled libnccl.so.2 is the one that gets used, otherwise the system one.
//
// Fast path over NVLink peer memory (default; GLB_P2P=0 turns it off): every rank owns an "arena" in
// device memory whose CUDA IPC handle is all-gathered once (through NCCL) at glb_comm_init, so each
// rank holds a mapped pointer to every peer's arena.  Ghost rows and reduction mailboxes live at
// IDENTICAL offsets in every arena:
//   * halo: a small kernel stores this rank's boundary rows straight into the neighbours' ghost rows
//     (remote 16-byte stores over NVLink), fences, and bumps a sequence flag in the neighbour's arena;
//     a one-thread kernel spins on the local flags before the stencil kernel runs.  Ghost rows are
//     double-buffered by exchange parity, so a rank may run one exchange ahead of its neighbour.
//   * allreduce: one kernel stores the k partial sums into slot [seq%4][rank] of EVERY peer's mailbox,
//     flags them, waits for the G local flags and adds the G contributions in rank order -- every
//     rank obtains the bit-identical sum, ~one NVLink store latency, no library launch.
// NCCL remains the bootstrap and the fallback when peer mapping is unavailable.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "p2p.cuh"
#include "runtime.hpp"

namespace glb {

// minimal NCCL surface (nccl.h: ncclUniqueId is 128 bytes; ncclFloat64 = 8; ncclSum = 0)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId_t;
enum { NCCL_SUCCESS = 0, NCCL_SUM = 0, NCCL_CHAR = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return GLB_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(GLB_ERR_COMM, std::string("cannot load libnccl: ") + dlerror());
#define BIND(field, sym)                                                      \
  *(void**)(&g_nccl.field) = dlsym(lib, sym);                                 \
  if (!g_nccl.field) return fail(GLB_ERR_COMM, std::string("libnccl lacks ") + sym);
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllReduce, "ncclAllReduce")
  BIND(Send, "ncclSend")
  BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_nccl.lib = lib;
  return GLB_OK;
}

#define GLB_NCCL(expr)                                                                                      \
  do {                                                                                                      \
    int _r = (expr);                                                                                        \
    if (_r != NCCL_SUCCESS) return fail(GLB_ERR_COMM, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

struct Comm {
  ncclComm_t nccl = nullptr;
  double* d_red = nullptr;  // device staging for host-value reductions
  double* h_red = nullptr;  // pinned
  // peer-memory fast path
  bool p2p = false;
  char* arena = nullptr;
  size_t arena_bytes = 0, arena_used = 0;
  char* peer[P2P_MAX_RANKS] = {nullptr};
  unsigned long long red_seq = 0;
  unsigned int* ticket = nullptr;
  struct FreeRegion {
    size_t off, bytes;
    unsigned long long seq;
  };
  std::vector<FreeRegion> arena_free;  // regions returned by destroyed operators, with their last exchange number
  long long spin_budget = 0;  // clock64() ticks a kernel waits for a peer (GLB_P2P_TIMEOUT_S; 0 = for ever)
  int flush = 1;              // GLB_P2P_FLUSH: fence.sys after the sends of a reduction
};

void comm_destroy(glb_context* ctx) {
  if (!ctx->comm) return;
  for (int g = 0; g < ctx->nranks; g++)
    if (ctx->comm->p2p && g != ctx->rank && ctx->comm->peer[g]) cudaIpcCloseMemHandle(ctx->comm->peer[g]);
  cudaFree(ctx->comm->arena);
  cudaFree(ctx->comm->ticket);
  if (ctx->comm->nccl) g_nccl.CommDestroy(ctx->comm->nccl);
  cudaFree(ctx->comm->d_red);
  cudaFreeHost(ctx->comm->h_red);
  delete ctx->comm;
  ctx->comm = nullptr;
}

// ------------------------------------------------------------------------------------------ peer-memory kernels
// boundary rows -> the neighbours' ghost rows (remote stores), then their flags
__global__ void __launch_bounds__(256) halo_push_kernel(const uint4* send_lo, const uint4* send_hi, uint4* dst_down_hi,
                                                         uint4* dst_up_lo, size_t n16, unsigned long long* flag_down_hi,
                                                         unsigned long long* flag_up_lo, unsigned long long seq,
                                                         unsigned int* ticket) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n16; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n16)
      dst_down_hi[i] = send_lo[i];
    else
      dst_up_lo[i - n16] = send_hi[i - n16];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    if (t == gridDim.x - 1) {  // every block's stores are fenced: publish
      __threadfence_system();
      st_release_sys(flag_down_hi, seq);
      st_release_sys(flag_up_lo, seq);
    }
  }
}
// spin until both neighbours have delivered exchange number `seq`
__global__ void halo_wait_kernel(const unsigned long long* flag_lo, const unsigned long long* flag_hi,
                                 unsigned long long seq, long long budget) {
  spin_until(flag_lo, seq, budget);
  spin_until(flag_hi, seq, budget);
}
// stand-alone one-shot allreduce (used when no producing kernel can finish the sum itself): one warp per peer
__global__ void p2p_allreduce_kernel(double* vals, int n, P2PRed pr) {
  __shared__ double s_v[P2P_RED_WIDTH];
  if (threadIdx.x == 0)
    for (int t = 0; t < n; t++) s_v[t] = vals[t];
  p2p_allreduce_block(pr, s_v, n);  // reads and writes through thread 0
  if (threadIdx.x == 0)
    for (int t = 0; t < n; t++) vals[t] = s_v[t];
}

// ---- measurement aid: how long does the rank-wide sum take when every rank arrives at (nearly) the same time?
// One warp per rank: [busy-wait busy_cycles] -> [sum of 6 doubles over ranks] -> record the time spent in the sum.
//   variant = 10 * send + recv
//   send 0: st.relaxed.sys words, lane g -> rank g (the first protocol)   1: + fence.sys after the sends   2: st.release.sys last word
//        3: atom.exch.sys words (not posted)
//   recv 0: ld.relaxed.sys poll   1: ld.acquire.sys   2: atom.add.sys(+0) poll (read at the home L2)   3: ld.cv
//        4: ld.relaxed.sys with __nanosleep(200) between polls   5: ld.volatile
__device__ __forceinline__ unsigned long long poll_word(const unsigned long long* w, int recv) {
  unsigned long long v;
  switch (recv) {
    case 1: return ld_acquire_sys(w);
    case 2: return atomicAdd_system((unsigned long long*)w, 0ull);
    case 3: asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory"); return v;
    case 5: asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory"); return v;
    default: return ld_relaxed_sys(w);
  }
}
__global__ void p2p_bench_kernel(P2PRed pr, int iters, long long busy_cycles, int variant, float* wait_us) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  __shared__ double s_got[8][6];
  if ((variant / 10) % 10 != 7 && warp > 0) return;  // only variant 7x uses the other warps
  const int send = (variant / 10) % 10, recv = variant % 10;
  if (variant >= 100) {  // only the first `active` ranks take part (the others leave at once)
    const int active = variant / 100;
    if (pr.rank >= active) return;
    pr.nranks = active;
  }
  Mailbox* mine = pr.mb[pr.rank];
  Mailbox* peer = pr.mb[lane < pr.nranks ? lane : pr.rank];
  for (int it = 0; it < iters; it++, pr.seq++) {
    const long long c0 = clock64();
    while (clock64() - c0 < busy_cycles) {
    }
    __syncwarp();
    unsigned long long ta, tb;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ta));
    const int slot = (int)(pr.seq % P2P_RED_SLOTS);
    const unsigned long long tag = ((pr.seq % 0xffffffffull) + 1ull) << 32;
    double vals[6];
    for (int t = 0; t < 6; t++) vals[t] = 1.0 + t + pr.rank;
    double sum[6];
    if (send == 7) {
      // one warp per peer: warp g sends this rank's 12 words to rank g with ONE store instruction (lane = word), then
      // polls the 12 words rank g sent here; the sums are formed through shared memory
      __syncthreads();
      if (warp < pr.nranks && lane < 12) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[lane / 2]);
        const unsigned long long word = tag | ((lane & 1) ? (bits >> 32) : (bits & 0xffffffffull));
        st_relaxed_sys(&pr.mb[warp]->ll[slot][pr.rank][lane], word);
        const unsigned long long* w = &mine->ll[slot][warp][lane];
        unsigned long long v;
        const long long t0 = clock64();
        for (;;) {
          v = ld_relaxed_sys(w);
          if ((v & 0xffffffff00000000ull) == tag) break;
          if (pr.budget > 0 && clock64() - t0 > pr.budget) __trap();
        }
        const unsigned long long other = __shfl_xor_sync(0xfffu, v, 1);
        if (!(lane & 1))
          s_got[warp][lane / 2] = __longlong_as_double((long long)((v & 0xffffffffull) | (other << 32)));
      }
      __syncthreads();
      for (int t = 0; t < 6; t++) {
        double sacc = 0.0;
        for (int r = 0; r < pr.nranks; r++) sacc += s_got[r][t];
        sum[t] = sacc;
      }
    } else if (send == 8) {
      // recursive doubling: log2(ranks) rounds, one partner per round; lane = word (12 words = 6 values).  Rows 8..11
      // of the mailbox (sending-rank index) serve as the rounds' receive buffers.
      double cur = (lane < 12) ? vals[lane / 2] : 0.0;
      int round = 0;
      for (int d = 1; d < pr.nranks; d <<= 1, round++) {
        const int partner = pr.rank ^ d;
        if (lane < 12) {
          const unsigned long long bits = (unsigned long long)__double_as_longlong(cur);
          const unsigned long long word = tag | ((lane & 1) ? (bits >> 32) : (bits & 0xffffffffull));
          st_relaxed_sys(&pr.mb[partner]->ll[slot][8 + round][lane], word);
          const unsigned long long* w = &mine->ll[slot][8 + round][lane];
          unsigned long long v;
          const long long t0 = clock64();
          for (;;) {
            v = ld_relaxed_sys(w);
            if ((v & 0xffffffff00000000ull) == tag) break;
            if (pr.budget > 0 && clock64() - t0 > pr.budget) __trap();
          }
          const unsigned long long other = __shfl_xor_sync(0xfffu, v, 1);
          const unsigned long long lo = (lane & 1) ? other : v, hi = (lane & 1) ? v : other;
          const double theirs = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
          // the lower rank's partial goes first: both partners form the same sum bit for bit
          cur = (pr.rank & d) ? (theirs + cur) : (cur + theirs);
        }
      }
      for (int t = 0; t < 6; t++) sum[t] = __shfl_sync(full, cur, 2 * t);
    } else if (send >= 5) {
      // spread over the warp: lane = part * 8 + peer (up to 8 ranks); every lane sends / polls only its share of the
      // six values, so no thread issues more than two or three system-scope stores in a row
      const int g = lane & 7, part = lane >> 3;  // 4 parts
      Mailbox* pg = pr.mb[g < pr.nranks ? g : pr.rank];
      if (g < pr.nranks) {
        for (int t = part; t < 6; t += 4) {
          const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[t]);
          const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
          unsigned long long* dst = &pg->ll[slot][pr.rank][2 * t];
          if (send == 6) {
            asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
          } else {
            st_relaxed_sys(dst, w0);
            st_relaxed_sys(dst + 1, w1);
          }
        }
      }
      double got[2] = {0.0, 0.0};
      if (g < pr.nranks) {
        int k = 0;
        for (int t = part; t < 6; t += 4, k++) {
          const unsigned long long* w = &mine->ll[slot][g][2 * t];
          unsigned long long lo, hi;
          const long long t0 = clock64();
          for (;;) {
            if (send == 6) {
              asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(w) : "memory");
            } else {
              lo = ld_relaxed_sys(w);
              hi = ld_relaxed_sys(w + 1);
            }
            if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
            if (pr.budget > 0 && clock64() - t0 > pr.budget) __trap();
          }
          got[k] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
        }
      }
      for (int t = 0; t < 6; t++) {
        double sacc = 0.0;
        for (int r = 0; r < pr.nranks; r++) sacc += __shfl_sync(full, got[t / 4], (t % 4) * 8 + r);  // rank order
        sum[t] = sacc;
      }
    } else {
    if (lane < pr.nranks) {
      for (int w = 0; w < 12; w++) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[w / 2]);
        const unsigned long long word = tag | ((w & 1) ? (bits >> 32) : (bits & 0xffffffffull));
        unsigned long long* dst = &peer->ll[slot][pr.rank][w];
        if (send == 3)
          atomicExch_system(dst, word);
        else if (send == 2 && w == 11)
          st_release_sys(dst, word);
        else
          st_relaxed_sys(dst, word);
      }
    }
    if (send == 1) __threadfence_system();
    for (int t = 0; t < 6; t++) {
      double got = 0.0;
      if (lane < pr.nranks) {
        const unsigned long long* w = &mine->ll[slot][lane][2 * t];
        unsigned long long lo, hi;
        const long long t0 = clock64();
        for (;;) {
          lo = poll_word(w, recv);
          hi = poll_word(w + 1, recv);
          if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
          if (recv == 4) __nanosleep(200);
          if (pr.budget > 0 && clock64() - t0 > pr.budget) __trap();
        }
        got = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
      }
      double s = 0.0;
      for (int g = 0; g < pr.nranks; g++) s += __shfl_sync(full, got, g);
      sum[t] = s;
    }
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tb));
    if (lane == 0) wait_us[it] = (float)(tb - ta) * 1e-3f + (sum[0] < 0.0 ? 1.0f : 0.0f);
  }
}

// ping-pong between rank 0 and rank `peer`: one 8-byte word each way per round trip; rtt_us[i] on rank 0
__global__ void p2p_pingpong_kernel(P2PRed pr, int peer, int iters, float* rtt_us) {
  if (pr.rank != 0 && pr.rank != peer) return;
  const int other = pr.rank == 0 ? peer : 0;
  unsigned long long* out = &pr.mb[other]->ll[0][pr.rank][0];
  const unsigned long long* in = &pr.mb[pr.rank]->ll[0][other][0];
  for (int it = 0; it < iters; it++) {
    const unsigned long long word = ((pr.seq + it) << 8) | 1ull;
    unsigned long long ta, tb;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ta));
    if (pr.rank == 0) {
      st_relaxed_sys(out, word);
      while (ld_relaxed_sys(in) != word) {
      }
    } else {
      while (ld_relaxed_sys(in) != word) {
      }
      st_relaxed_sys(out, word);
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tb));
    if (pr.rank == 0) rtt_us[it] = (float)(tb - ta) * 1e-3f;
    const long long c0 = clock64();
    while (clock64() - c0 < 20000) {
    }
  }
}

bool comm_p2p(const glb_context* ctx) { return ctx->comm && ctx->comm->p2p; }
char* comm_peer(glb_context* ctx, int g) { return ctx->comm->peer[g]; }
long long comm_spin_budget(const glb_context* ctx) { return ctx->comm ? ctx->comm->spin_budget : 0; }
unsigned int* comm_ticket(glb_context* ctx) { return ctx->comm->ticket; }

// descriptor of the NEXT rank-wide reduction (bumps the sequence number: one per reduction, same order on all ranks)
P2PRed comm_p2p_red(glb_context* ctx) {
  P2PRed pr;
  Comm* c = ctx->comm;
  for (int g = 0; g < P2P_MAX_RANKS; g++) pr.mb[g] = (g < ctx->nranks) ? (Mailbox*)c->peer[g] : nullptr;
  pr.rank = ctx->rank;
  pr.nranks = ctx->nranks;
  pr.seq = ++c->red_seq;
  pr.budget = c->spin_budget;
  pr.flush = c->flush;
  return pr;
}

// the same for a kernel that performs up to `count` reductions of its own (step s uses seq + s): reserves the range
P2PRed comm_p2p_red_range(glb_context* ctx, unsigned long long count) {
  P2PRed pr = comm_p2p_red(ctx);
  if (count > 1) ctx->comm->red_seq += count - 1;
  return pr;
}

// Start halo exchange number op->halo_seq+1 on the peer-memory path: where this rank's boundary rows go
// (remote pointers), which flags to raise there, which local flags to wait on.  Also points
// op->ghost_lo/hi at the buffers of this exchange's parity.
int halo_p2p_begin(glb_operator* op, int nrows, HaloTargets* t) {
  glb_context* ctx = op->ctx;
  Comm* c = ctx->comm;
  if (!c || !c->p2p || !op->ghost_p2p) return fail(GLB_ERR_STATE, "operator is not on the peer-memory path");
  const int G = ctx->nranks, g = ctx->rank;
  const int up = (g + 1) % G, down = (g + G - 1) % G;
  const size_t rowb = (size_t)op->X * op->nc * elem_bytes(op->dtype);
  const size_t gbytes = rowb * op->ghost_depth;
  const unsigned long long seq = ++op->halo_seq;
  const size_t par = (size_t)(seq & 1);
  // arena layout of this operator: [parity 0: lo | hi][parity 1: lo | hi][flag_lo][flag_hi]
  const size_t off_lo = op->ghost_off + par * 2 * gbytes, off_hi = off_lo + gbytes;
  const size_t off_flag = op->ghost_off + 4 * gbytes;
  t->dst_down_hi = c->peer[down] + off_hi;                                 // its rows Yloc ..
  t->dst_up_lo = c->peer[up] + off_lo + rowb * (op->ghost_depth - nrows);  // its rows -nrows .. -1
  t->flag_down_hi = (unsigned long long*)(c->peer[down] + off_flag + 8);
  t->flag_up_lo = (unsigned long long*)(c->peer[up] + off_flag);
  t->wait.flag_lo = (const unsigned long long*)(c->arena + off_flag);
  t->wait.flag_hi = (const unsigned long long*)(c->arena + off_flag + 8);
  t->wait.seq = seq;
  t->wait.budget = c->spin_budget;
  t->ticket = c->ticket;
  t->bytes = rowb * nrows;
  op->ghost_lo = c->arena + off_lo;
  op->ghost_hi = c->arena + off_hi;
  return GLB_OK;
}

// carve `bytes` (256-byte aligned) out of the arena; identical call sequences on all ranks give
// identical offsets.  Regions of destroyed operators are reused (exact size match: operators of one lattice
// size come and go, e.g. a new gauge field per configuration or a multigrid set-up per solve).  A reused region is
// NOT cleared -- ranks are not synchronised here, a slower peer may still be storing the old operator's last rows
// into it -- instead the new operator continues the old one's exchange numbering (*seq_start): flags only ever
// grow, and ghost rows are read only after the flag of their own exchange has arrived.  Returns nullptr when the
// arena is exhausted (the caller falls back to NCCL buffers).
void* comm_arena_alloc(glb_context* ctx, size_t bytes, size_t* offset, unsigned long long* seq_start) {
  Comm* c = ctx->comm;
  if (seq_start) *seq_start = 0;
  if (!c || !c->p2p) return nullptr;
  const size_t need = (bytes + 255) & ~(size_t)255;
  for (size_t i = 0; i < c->arena_free.size(); i++) {
    if (c->arena_free[i].bytes == need) {
      *offset = c->arena_free[i].off;
      if (seq_start) *seq_start = c->arena_free[i].seq;
      c->arena_free.erase(c->arena_free.begin() + i);
      return c->arena + *offset;
    }
  }
  if (c->arena_used + need > c->arena_bytes) {
    if (getenv("GLB_VERBOSE"))
      fprintf(stderr, "[glb200] rank %d: peer-memory arena exhausted (%zu MB, GLB_P2P_ARENA_MB): this operator's halos "
                      "go through NCCL\n", ctx->rank, c->arena_bytes >> 20);
    return nullptr;
  }
  *offset = c->arena_used;
  c->arena_used += need;
  return c->arena + *offset;
}
// give a region back (same order on every rank, so the free lists stay identical); seq = its last exchange number
void comm_arena_free(glb_context* ctx, size_t offset, size_t bytes, unsigned long long seq) {
  Comm* c = ctx->comm;
  if (!c || !c->p2p) return;
  Comm::FreeRegion f = {offset, (bytes + 255) & ~(size_t)255, seq};
  c->arena_free.push_back(f);
}

static int halo_exchange_p2p(glb_operator* op, const void* send_lo, const void* send_hi, int nrows, HaloWait* defer) {
  glb_context* ctx = op->ctx;
  HaloTargets t;
  int rc = halo_p2p_begin(op, nrows, &t);
  if (rc) return rc;
  const size_t n16 = t.bytes / 16;
  int grid = (int)((2 * n16 + 255) / 256);
  if (grid > 64) grid = 64;
  halo_push_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4*)send_lo, (const uint4*)send_hi, (uint4*)t.dst_down_hi,
                                                  (uint4*)t.dst_up_lo, n16, t.flag_down_hi, t.flag_up_lo, t.wait.seq,
                                                  t.ticket);
  GLB_LAUNCH_CHECK();
  if (defer != nullptr) {  // the consuming kernel waits itself, and only where it reads a ghost row
    *defer = t.wait;
    return GLB_OK;
  }
  halo_wait_kernel<<<1, 1, 0, ctx->stream>>>(t.wait.flag_lo, t.wait.flag_hi, t.wait.seq, t.wait.budget);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int halo_exchange_ptrs(glb_operator* op, const void* send_lo, const void* send_hi, int nrows, HaloWait* defer) {
  glb_context* ctx = op->ctx;
  if (defer != nullptr) *defer = HaloWait{};
  if (ctx->nranks == 1) return GLB_OK;
  if (!ctx->comm) return fail(GLB_ERR_STATE, "slab operator used before glb_comm_init");
  if (nrows < 1 || nrows > op->ghost_depth) return fail(GLB_ERR_ARG, "halo deeper than the operator's ghost rows");
  if (op->ghost_p2p) return halo_exchange_p2p(op, send_lo, send_hi, nrows, defer);
  const int G = ctx->nranks, g = ctx->rank;
  const int up = (g + 1) % G, down = (g + G - 1) % G;
  const size_t rowb = (size_t)op->X * op->nc * elem_bytes(op->dtype);
  const size_t bytes = rowb * nrows;
  // my lowest rows -> `down`'s ghost_hi (its rows Yloc ..) ; my highest rows -> `up`'s ghost_lo (its rows -nrows .. -1)
  char* recv_lo = (char*)op->ghost_lo + rowb * (op->ghost_depth - nrows);
  char* recv_hi = (char*)op->ghost_hi;
  // Order matters on 2 ranks, where `up` and `down` are the same peer and NCCL pairs the k-th send
  // with the peer's k-th receive: (hi -> up) must meet the peer's (ghost_lo <- down).
  GLB_NCCL(g_nccl.GroupStart());
  GLB_NCCL(g_nccl.Send(send_hi, bytes, NCCL_CHAR, up, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.Send(send_lo, bytes, NCCL_CHAR, down, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.Recv(recv_lo, bytes, NCCL_CHAR, down, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.Recv(recv_hi, bytes, NCCL_CHAR, up, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.GroupEnd());
  return GLB_OK;
}

int halo_exchange(glb_operator* op, const void* in, int nrows, HaloWait* defer) {
  if (defer != nullptr) *defer = HaloWait{};
  if (op->ctx->nranks == 1) return GLB_OK;
  if (nrows > op->Yloc) return fail(GLB_ERR_ARG, "slab thinner than the halo");
  const size_t rowb = (size_t)op->X * op->nc * elem_bytes(op->dtype);
  const char* base = (const char*)in;
  return halo_exchange_ptrs(op, base, base + rowb * (op->Yloc - nrows), nrows, defer);
}

int allreduce_device(glb_context* ctx, double* d_vals, int n);

int allreduce_sum(glb_context* ctx, double* vals, int n) {
  if (ctx->nranks == 1) return GLB_OK;
  if (!ctx->comm) return fail(GLB_ERR_STATE, "reduction before glb_comm_init");
  if (n > 64) return fail(GLB_ERR_ARG, "allreduce_sum: too many values");
  std::memcpy(ctx->comm->h_red, vals, sizeof(double) * n);
  GLB_CUDA(cudaMemcpyAsync(ctx->comm->d_red, ctx->comm->h_red, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  {
    int rc = allreduce_device(ctx, ctx->comm->d_red, n);
    if (rc) return rc;
  }
  GLB_CUDA(cudaMemcpyAsync(ctx->comm->h_red, ctx->comm->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(vals, ctx->comm->h_red, sizeof(double) * n);
  return GLB_OK;
}

// in-stream sum over ranks of n doubles living in device memory (device-resident CG)
int allreduce_device(glb_context* ctx, double* d_vals, int n) {
  if (ctx->nranks == 1) return GLB_OK;
  if (!ctx->comm) return fail(GLB_ERR_STATE, "reduction before glb_comm_init");
  Comm* c = ctx->comm;
  if (c->p2p && n <= P2P_RED_WIDTH) {
    p2p_allreduce_kernel<<<1, 32 * (ctx->nranks < 16 ? ctx->nranks : 16), 0, ctx->stream>>>(d_vals, n, comm_p2p_red(ctx));
    GLB_LAUNCH_CHECK();
    return GLB_OK;
  }
  GLB_NCCL(g_nccl.AllReduce(d_vals, d_vals, n, NCCL_FLOAT64, NCCL_SUM, ctx->comm->nccl, ctx->stream));
  return GLB_OK;
}

}  // namespace glb

using namespace glb;

extern "C" {

int glb_comm_unique_id(char id[GLB_COMM_ID_BYTES]) {
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t u;
  GLB_NCCL(g_nccl.GetUniqueId(&u));
  std::memcpy(id, u.internal, GLB_COMM_ID_BYTES);
  return GLB_OK;
}

int glb_comm_init(glb_context* ctx, int rank, int nranks, const char id[GLB_COMM_ID_BYTES]) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GLB_ERR_ARG, "glb_comm_init: bad rank/nranks");
  if (ctx->comm) return fail(GLB_ERR_STATE, "communicator already initialised");
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return GLB_OK;
  int rc = load_nccl();
  if (rc) return rc;
  GLB_CUDA(cudaSetDevice(ctx->device));
  Comm* c = new Comm();
  {
    // how long a kernel spins for a peer before it traps: seconds -> clock64() ticks at the SM's maximum clock
    const char* et = getenv("GLB_P2P_TIMEOUT_S");
    const double secs = et ? atof(et) : 600.0;
    int khz = 2000000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    c->spin_budget = secs > 0.0 ? (long long)(secs * 1e3 * (double)khz) : 0;
    const char* ef = getenv("GLB_P2P_FLUSH");
    c->flush = ef ? atoi(ef) : 1;
  }
  ncclUniqueId_t u;
  std::memcpy(u.internal, id, GLB_COMM_ID_BYTES);
  GLB_NCCL(g_nccl.CommInitRank(&c->nccl, nranks, u, rank));
  GLB_CUDA(cudaMalloc(&c->d_red, sizeof(double) * 64));
  GLB_CUDA(cudaHostAlloc((void**)&c->h_red, sizeof(double) * 64, cudaHostAllocDefault));
  ctx->comm = c;
  // ---- peer-memory arena: export, all-gather the IPC handles through NCCL, map every peer
  const char* e = getenv("GLB_P2P");
  const bool want = !(e && atoi(e) == 0) && nranks <= P2P_MAX_RANKS;
  int ok = 0;
  if (want) {
    const char* es = getenv("GLB_P2P_ARENA_MB");
    c->arena_bytes = (size_t)(es ? atoi(es) : 96) << 20;
    ok = (cudaMalloc((void**)&c->arena, c->arena_bytes) == cudaSuccess) &&
         (cudaMemset(c->arena, 0, c->arena_bytes) == cudaSuccess) &&
         (cudaMalloc((void**)&c->ticket, sizeof(unsigned int)) == cudaSuccess) &&
         (cudaMemset(c->ticket, 0, sizeof(unsigned int)) == cudaSuccess);
  }
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof mine);
  if (ok) ok = (cudaIpcGetMemHandle(&mine, c->arena) == cudaSuccess);
  // gather [ok flag | handle] from everybody (fixed-size record so a failing rank cannot desynchronise)
  const size_t rec = 8 + sizeof(cudaIpcMemHandle_t);
  char* d_all = nullptr;
  GLB_CUDA(cudaMalloc((void**)&d_all, rec * nranks));
  std::vector<char> h_all(rec * nranks, 0);
  long long okl = ok;
  std::memcpy(&h_all[rec * rank], &okl, 8);
  std::memcpy(&h_all[rec * rank + 8], &mine, sizeof mine);
  GLB_CUDA(cudaMemcpy(d_all + rec * rank, &h_all[rec * rank], rec, cudaMemcpyHostToDevice));
  typedef int (*AllGatherFn)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  AllGatherFn all_gather = (AllGatherFn)dlsym(g_nccl.lib, "ncclAllGather");
  if (!all_gather) return fail(GLB_ERR_COMM, "libnccl lacks ncclAllGather");
  GLB_NCCL(all_gather(d_all + rec * rank, d_all, rec, NCCL_CHAR, c->nccl, ctx->stream));
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  GLB_CUDA(cudaMemcpy(h_all.data(), d_all, rec * nranks, cudaMemcpyDeviceToHost));
  GLB_CUDA(cudaFree(d_all));
  bool all_ok = want;
  for (int g = 0; g < nranks && all_ok; g++) {
    long long v;
    std::memcpy(&v, &h_all[rec * g], 8);
    all_ok = (v != 0);
  }
  int mapped = all_ok ? 1 : 0;
  if (all_ok) {
    for (int g = 0; g < nranks && mapped; g++) {
      if (g == rank) {
        c->peer[g] = c->arena;
        continue;
      }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, &h_all[rec * g + 8], sizeof h);
      if (cudaIpcOpenMemHandle((void**)&c->peer[g], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mapped = 0;
      }
    }
  }
  // everybody must agree, otherwise nobody uses the fast path
  double agree = mapped ? 0.0 : 1.0;
  c->p2p = false;
  int rc2 = allreduce_sum(ctx, &agree, 1);
  if (rc2) return rc2;
  c->p2p = (agree == 0.0);
  c->arena_used = (sizeof(Mailbox) + 255) & ~(size_t)255;
  if (!c->p2p && getenv("GLB_VERBOSE")) fprintf(stderr, "[glb200] rank %d: peer-memory path unavailable, using NCCL\n", rank);
  return GLB_OK;
}

int glb_comm_p2p_enabled(glb_context* ctx) { return comm_p2p(ctx) ? 1 : 0; }

// measurement aid (tools/p2p_bench.py): see p2p_bench_kernel.  Collective: every rank calls it with equal arguments.
int glb_dbg_p2p_pingpong(glb_context* ctx, int peer, int iters, float* rtt_us_host) {
  if (!comm_p2p(ctx)) return fail(GLB_ERR_STATE, "glb_dbg_p2p_pingpong needs the peer-memory communicator");
  float* d = nullptr;
  GLB_CUDA(cudaMalloc((void**)&d, sizeof(float) * iters));
  GLB_CUDA(cudaMemset(d, 0, sizeof(float) * iters));
  P2PRed pr = comm_p2p_red_range(ctx, (unsigned long long)iters);
  p2p_pingpong_kernel<<<1, 1, 0, ctx->stream>>>(pr, peer, iters, d);
  GLB_LAUNCH_CHECK();
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  GLB_CUDA(cudaMemcpy(rtt_us_host, d, sizeof(float) * iters, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return GLB_OK;
}

int glb_dbg_p2p_bench(glb_context* ctx, int variant, int iters, double busy_us, float* wait_us_host) {
  if (!comm_p2p(ctx)) return fail(GLB_ERR_STATE, "glb_dbg_p2p_bench needs the peer-memory communicator");
  float* d = nullptr;
  GLB_CUDA(cudaMalloc((void**)&d, sizeof(float) * iters));
  int khz = 2000000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
  P2PRed pr = comm_p2p_red_range(ctx, (unsigned long long)iters);
  p2p_bench_kernel<<<1, 256, 0, ctx->stream>>>(pr, iters, (long long)(busy_us * 1e-3 * khz), variant, d);
  GLB_LAUNCH_CHECK();
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  GLB_CUDA(cudaMemcpy(wait_us_host, d, sizeof(float) * iters, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return GLB_OK;
}
int glb_comm_rank(glb_context* ctx) { return ctx->rank; }
int glb_comm_size(glb_context* ctx) { return ctx->nranks; }

int glb_comm_barrier(glb_context* ctx) {
  double z = 0.0;
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  return allreduce_sum(ctx, &z, 1);
}

}  // extern "C"
