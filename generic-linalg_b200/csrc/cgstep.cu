// cgstep.cu -- one WHOLE iteration of minv_vector_cg on D^dag D (generic_cg.cpp:324-354 with
// square_staggered_normal_u1, operators.cpp:444) as ONE kernel and ONE rank-wide reduction.
//
// The reference iteration is   alpha = rsq/<p,Ap>; x += alpha p; r -= alpha Ap; rsqNew = |r|^2;
//                              beta = rsqNew/rsq; p = r + beta p; Ap = D^dag D p
// i.e. two global synchronisation points (|r|^2 before beta, <p,Ap> before alpha) and, fused as far as
// those allow, two kernels moving 96 + 96 B/site.  Here step i does, in one pass over the lattice,
//     r_i = r_{i-1} - alpha_{i-1} q_{i-1} ;  x += alpha_{i-1} p_{i-1} ;  p_i = r_i + beta_i p_{i-1} ;
//     q_i = D^dag D p_i ;   sums  |r_i|^2, <p_i,q_i>, <r_i,q_i>, |q_i|^2
// reading r, q, p, x, U_x, U_y (96 B/site) and writing r, p, q, x (64 B/site): 160 B/site per iteration
// instead of 192, one launch, one reduction.  beta_i = |r_i|^2/|r_{i-1}|^2 is needed while r_i is being
// formed, so its numerator is PREDICTED by the previous step from the sums it already has:
//     |r - alpha q|^2 = |r|^2 - 2 Re(alpha <r,q>) + |alpha|^2 |q|^2 .
// Every step also sums the exact |r_i|^2, which feeds alpha_i = |r_i|^2/<p_i,q_i>, the stopping test
// (generic_cg.cpp:339) and the next prediction: prediction errors (a few ulp) never accumulate, the
// iteration count and the final residual are the reference's (tests: equal counts 64^2 .. 4096^2).
//
// Shape (B200): a CTA = 4 consumer warps + 1 producer warp sweeps a (112+4)-site wide strip up the rows
// of its row block.  The producer's elected lane streams whole tile rows of the six arrays into a ring
// of shared-memory stages with cp.async.bulk (TMA, SASS UBLKCP) completing on mbarriers ("full"); the
// consumers read their sites and x neighbours from the stage, release it ("empty") and keep p, t = D p,
// r and the links of three consecutive rows in registers (three-slot windows renamed by unrolling the
// row loop by three, as normal1.cu).  Consumers issue no global loads at all; only the 2-site tile halo
// is fetched twice from L2 (3.6 %).  One site per thread, the reference's expression order without FMA
// contraction: given equal scalars every vector update is bit-identical to the CPU code.
//
// Slabs (peer memory): the warps that own the first / last two rows of the slab store r_i, p_i, q_i of
// those rows straight into the neighbours' ghost rows (remote stores over NVLink) while the rest of the
// kernel runs; the last of them raises the neighbour's flag.  The next step's producers of boundary
// row blocks wait for that flag before they read ghost rows -- interior row blocks never wait.  The
// kernel's last block completes the sum of the six scalars over ranks itself (p2p_allreduce_block):
// one rank-wide synchronisation per CG iteration, no separate halo or reduction launch.
#include <climits>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "cg_state.cuh"
#include "cgstep.cuh"
#include "runtime.hpp"

namespace glb {

// ------------------------------------------------------------------ mbarrier / bulk-copy primitives
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy (TMA unit, no registers), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// hopping term at this lane's site, reference order (operators.cpp:215-224); eta = -1 on odd x
__device__ __forceinline__ cplx cs_hop(bool eta_neg, cplx ux, cplx ux_m, cplx uy, cplx uy_m, cplx psi_xp, cplx psi_xm,
                                       cplx psi_yp, cplx psi_ym) {
  cplx h = mk(0.0, 0.0);
  h = fsub(h, fmul(ux, psi_xp));
  h = fadd(h, fcmul(ux_m, psi_xm));
  cplx t3 = fmul(uy, psi_yp);
  if (eta_neg) t3 = fneg(t3);  // h - (-t3) == h + t3 exactly
  h = fsub(h, t3);
  cplx t4 = fcmul(uy_m, psi_ym);
  if (eta_neg) t4 = fneg(t4);
  h = fadd(h, t4);
  return h;
}
template <bool DAGGER>
__device__ __forceinline__ cplx cs_row(bool eta_neg, cplx below, cplx centre, cplx above, cplx ux, cplx ux_left, cplx uy,
                                       cplx uy_below, double mass) {
  const cplx left = shfl_up_c(centre, 1);
  const cplx right = shfl_down_c(centre, 1);
  cplx h = cs_hop(eta_neg, ux, ux_left, uy, uy_below, right, left, above, below);
  if (DAGGER) h = fneg(h);
  return fadd(fscale(0.5, h), fscale(mass, centre));
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// arrays of one stage, in this order
enum { CS_R = 0, CS_Q = 1, CS_P = 2, CS_XV = 3, CS_UX = 4, CS_UY = 5, CS_NARR = 6 };

__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int CW, int STAGES, int MINB>
__global__ void __launch_bounds__((CW + 1) * 32, MINB) cg_step_kernel(const CgStepArgs a) {
  constexpr int OUT_W = 28;           // sites a consumer warp produces per row (32 loaded - 2 halo sites per side)
  constexpr int OUT = CW * OUT_W;     // per CTA
  constexpr int TW = OUT + 4;         // tile width in sites
  constexpr unsigned ROW_BYTES = TW * sizeof(cplx);
  constexpr unsigned STAGE_BYTES = CS_NARR * ROW_BYTES;
  extern __shared__ __align__(128) unsigned char cs_smem[];
  cplx* const tiles = reinterpret_cast<cplx*>(cs_smem);
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(cs_smem + (size_t)STAGES * STAGE_BYTES);
  unsigned long long* const empty = full + STAGES;
  // Stage headers: the producer is also the scheduler.  Every stage carries {row L, ya, yb, x_lo} of the item it
  // belongs to (L = INT_MIN: end of this CG step), so the consumers are a uniform stream processor whatever item
  // the rows come from -- items follow each other in the ring without draining it.
  int4* const hdr = reinterpret_cast<int4*>(empty + STAGES);
  __shared__ double s_scal[3];  // alpha.re, alpha.im, beta of the current step
  __shared__ int s_ctl[3];      // step number, done, "this block finishes the step"
  __shared__ double s_wpart[CW][6];   // the consumer warps' sums of the step
  __shared__ double s_red[6 * 32];
  __shared__ unsigned long long s_rbar;  // mbarrier: all consumer warps have posted their sums
  unsigned long long* const rbar = &s_rbar;

  CgState* const st = a.st;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CW);
    }
    mbar_init(rbar, CW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int X = a.X, Y = a.Y;
  const int nstrips = a.nstrips, nrb = a.nrb;
  const int nitems = nstrips * nrb;
  const bool slab = (a.ghost != nullptr);
  const size_t gvec = (size_t)2 * X;  // one vector's two ghost rows
  unsigned j = 0;  // running stage counter: the same sequence in the producer and in every consumer
  int have_step = -1;

  // consumer state that lives across steps: three rotating register windows.  At a row step of phase K slot K2
  // takes the new row L, K1 holds row L-1, K0 row L-2.  A new item (or CG step) may start at any phase: its first
  // four rows only feed results that are predicated off (rows below ya), so nothing is carried over.
  const bool eta_neg = (lane & 1);  // tiles and warp windows start on even x
  const int idx = warp * OUT_W + lane;              // this lane's site inside the tile (consumers)
  const int idx_l = idx > 0 ? idx - 1 : 0;          // its left neighbour (lane 0 of warp 0 never uses it)
  const bool lane_in = (lane >= 2) && (lane < 30);
  cplx p[3], r[3], t[3], ux[3], uy[3], uxl[3];
#pragma unroll
  for (int k = 0; k < 3; k++) p[k] = r[k] = t[k] = ux[k] = uy[k] = uxl[k] = mk(0.0, 0.0);

#pragma unroll 1
  for (int n = 0; n < a.nsteps; n++) {
    // ---- the scalars of this step, published by the last block of the previous one (or by the host)
    if (threadIdx.x == 0) {
      int step = ld_acquire_gpu_s32(&st->step);
      if (n > 0) {
        const long long t0 = clock64();
        while (step <= have_step) {
          __nanosleep(32);
          step = ld_acquire_gpu_s32(&st->step);
          if (a.wait.budget > 0 && clock64() - t0 > a.wait.budget) __trap();
        }
      }
      s_ctl[0] = step;
      s_ctl[1] = __ldcg(&st->done);
      s_scal[0] = __ldcg(&st->alpha_re);
      s_scal[1] = __ldcg(&st->alpha_im);
      s_scal[2] = __ldcg(&st->beta);
    }
    __syncthreads();
    const int step = s_ctl[0];
    if (s_ctl[1]) break;  // uniform over the grid: rank-summed scalars are equal on every block and every rank
    have_step = step;
    const cplx alpha = mk(s_scal[0], s_scal[1]);
    const double beta = s_scal[2];
    const int cur = step & 1;  // ping-pong buffers of the recurrence
    const cplx* const r_in = a.r[cur];
    const cplx* const q_in = a.q[cur];
    const cplx* const p_in = a.p[cur];
    cplx* const r_out = a.r[cur ^ 1];
    cplx* const q_out = a.q[cur ^ 1];
    cplx* const p_out = a.p[cur ^ 1];
    // slabs: this step has exchange number seq; it reads the ghost rows of parity (seq-1)&1, stores its own boundary
    // rows into the neighbours' buffers of parity seq&1 and raises their flags to seq
    const unsigned long long seq = a.seq_base + (unsigned long long)step;
    const cplx* g_lo = nullptr;
    const cplx* g_hi = nullptr;
    cplx* push_down = nullptr;
    cplx* push_up = nullptr;
    if (slab) {
      const size_t rd = (size_t)((seq - 1) & 1) * 2 * (3 * gvec), wr = (size_t)(seq & 1) * 2 * (3 * gvec);
      g_lo = a.ghost + rd;
      g_hi = a.ghost + rd + 3 * gvec;
      push_down = a.peer_down + wr + 3 * gvec;  // my rows 0, 1 are rows Y, Y+1 of the rank below
      push_up = a.peer_up + wr;                 // my rows Y-2, Y-1 are rows -2, -1 of the rank above
    }
    double acc[6];  // |r|^2, <p,q>.re, <p,q>.im, <r,q>.re, <r,q>.im, |q|^2
#pragma unroll
    for (int i = 0; i < 6; i++) acc[i] = 0.0;
    const bool tracing = (a.trace != nullptr) && (step == a.trace_step);
    unsigned long long t_start = 0;
    int items_done = 0;
    if (a.trace != nullptr && threadIdx.x == 0) {
      t_start = gtime();
      if (step == a.trace_step + 1) a.trace[24 * (size_t)blockIdx.x + 13] = t_start;  // when the next step began here
    }

    if (warp == CW) {
      // ================================================================== producer (one elected lane)
      if (lane == 0) {
        // data written by the previous step's generic stores (any SM) is read by the TMA unit from here on
        asm volatile("fence.proxy.async;" ::: "memory");
        bool waited = false;
        int item = blockIdx.x;
        for (;;) {
          if (a.queue != nullptr) item = (int)atomicAdd(a.queue, 1u);  // dynamic schedule: first come, first served
          if (item >= nitems) break;
          // item -> (strip, row block); on slabs the two boundary row blocks come first so that their rows are on
          // their way to the neighbours while the interior is still being worked on
          const int strip = item % nstrips;
          int rb = item / nstrips;
          if (slab && nrb > 2) rb = (rb == 0) ? 0 : (rb == 1 ? nrb - 1 : rb - 1);
          if (a.queue == nullptr) item += gridDim.x;
          const int ya = a.rb_rows ? a.rb_rows[rb] : (int)((long long)Y * rb / nrb);
          const int yb = a.rb_rows ? a.rb_rows[rb + 1] : (int)((long long)Y * (rb + 1) / nrb);
          if (ya >= yb) continue;
          items_done++;
          const int x_lo = strip * OUT - 2;
          const int pos0 = ((x_lo % X) + X) % X;
          if (slab && !waited && (ya - 2 < 0 || yb + 2 > Y)) {
            // the neighbours' rows of the previous step must have landed before the TMA unit reads them
            spin_until(a.wait.flag_lo, seq - 1, a.wait.budget);
            spin_until(a.wait.flag_hi, seq - 1, a.wait.budget);
            asm volatile("fence.proxy.async;" ::: "memory");
            waited = true;
          }
          for (int L = ya - 2; L < yb + 2; L++, j++) {
            const int s = j % STAGES;
            mbar_wait(&empty[s], ((j / STAGES) & 1) ^ 1);
            hdr[s] = make_int4(L, ya, yb, x_lo);
            mbar_expect_tx(&full[s], STAGE_BYTES);
            cplx* const dst = tiles + (size_t)s * (CS_NARR * TW);
            // source rows: vectors at row L (periodic on one rank, ghost rows on slabs), links at row L-1
            const cplx* src[CS_NARR];
            if (slab && L < 0) {
              const cplx* g = g_lo + (size_t)(L + 2) * X;
              src[CS_R] = g;
              src[CS_Q] = g + gvec;
              src[CS_P] = g + 2 * gvec;
              src[CS_XV] = a.x;  // never used on ghost rows
            } else if (slab && L >= Y) {
              const cplx* g = g_hi + (size_t)(L - Y) * X;
              src[CS_R] = g;
              src[CS_Q] = g + gvec;
              src[CS_P] = g + 2 * gvec;
              src[CS_XV] = a.x;
            } else {
              const size_t o = (size_t)(((L % Y) + Y) % Y) * X;
              src[CS_R] = r_in + o;
              src[CS_Q] = q_in + o;
              src[CS_P] = p_in + o;
              src[CS_XV] = a.x + o;
            }
            const int lrow = (L - 1 < -LINK_GHOST) ? -LINK_GHOST : L - 1;  // row ya-3 is never used
            src[CS_UX] = a.Ux + (ptrdiff_t)lrow * X;
            src[CS_UY] = a.Uy + (ptrdiff_t)lrow * X;
            // the tile is periodic in x: split at the seam (several times when the lattice is narrower than the tile)
            int pos = pos0, rem = TW, d = 0;
            while (rem > 0) {
              const int seg = (rem < X - pos) ? rem : X - pos;
#pragma unroll
              for (int arr = 0; arr < CS_NARR; arr++)
                bulk_g2s(dst + arr * TW + d, src[arr] + pos, (unsigned)seg * (unsigned)sizeof(cplx), &full[s]);
              d += seg;
              rem -= seg;
              pos = 0;
            }
          }
        }
        // end of this step: one header-only stage sends the consumers to the reduction
        const int s = j % STAGES;
        mbar_wait(&empty[s], ((j / STAGES) & 1) ^ 1);
        hdr[s] = make_int4(INT_MIN, 0, 0, 0);
        mbar_arrive(&full[s]);
        j++;
      }
    } else {
      // ================================================================== consumers
      auto row_step = [&](auto Kc) -> bool {
        constexpr int K0 = decltype(Kc)::value % 3, K1 = (K0 + 1) % 3, K2 = (K0 + 2) % 3;
        const int s = j % STAGES;
        mbar_wait(&full[s], (j / STAGES) & 1);
        const int4 h = hdr[s];
        const int L = h.x, ya = h.y, yb = h.z;
        if (L == INT_MIN) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
          j++;
          return false;
        }
        const cplx* const tile = tiles + (size_t)s * (CS_NARR * TW);
        const cplx ro = tile[CS_R * TW + idx];
        const cplx qo = tile[CS_Q * TW + idx];
        const cplx po = tile[CS_P * TW + idx];
        const cplx xo = tile[CS_XV * TW + idx];
        ux[K2] = tile[CS_UX * TW + idx];   // U_x(L-1)
        uxl[K2] = tile[CS_UX * TW + idx_l];
        uy[K2] = tile[CS_UY * TW + idx];   // U_y(L-1)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        j++;
        const int x = h.w + idx;
        const bool active = lane_in && (x < X);
        // the streaming part of the iteration at row L (generic_cg.cpp:328-351)
        const cplx rn = fsub(ro, fmul(alpha, qo));        // r = r - alpha*Ap
        const cplx pn = fadd(rn, fscale(beta, po));       // p = r + beta*p
        r[K2] = rn;
        p[K2] = pn;
        const bool own = active && (L >= ya) && (L < yb);
        if (own) {
          const size_t o = (size_t)L * X + x;
          r_out[o] = rn;
          p_out[o] = pn;
          a.x[o] = fadd(xo, fmul(alpha, po));             // phi = phi + alpha*p
          acc[0] += fnorm(rn);
          if (slab) {
            if (L < 2) {
              cplx* g = push_down + (size_t)L * X + x;
              g[0] = rn;
              g[2 * gvec] = pn;
            }
            if (L >= Y - 2) {
              cplx* g = push_up + (size_t)(L - (Y - 2)) * X + x;
              g[0] = rn;
              g[2 * gvec] = pn;
            }
          }
        }
        // t(L-1) = D p at row L-1 ; q(L-2) = D^dag t at row L-2   (operators.cpp:444-453, one pass)
        t[K2] = cs_row<false>(eta_neg, p[K0], p[K1], p[K2], ux[K2], uxl[K2], uy[K2], uy[K1], a.mass);
        const cplx res = cs_row<true>(eta_neg, t[K0], t[K1], t[K2], ux[K1], uxl[K1], uy[K1], uy[K0], a.mass);
        const int y = L - 2;
        if (active && y >= ya) {
          const size_t o = (size_t)y * X + x;
          q_out[o] = res;
          Field<cplx>::dot_acc(acc + 1, p[K0], res);
          Field<cplx>::dot_acc(acc + 3, r[K0], res);
          acc[5] += fnorm(res);
          if (slab) {
            if (y < 2) push_down[gvec + (size_t)y * X + x] = res;
            if (y >= Y - 2) push_up[gvec + (size_t)(y - (Y - 2)) * X + x] = res;
          }
        }
        if (slab && L == yb + 1 && (ya < 2 || yb > Y - 2)) {
          // last row of an item that owns boundary rows of the slab: this warp's share of them is on its way to the
          // neighbours -- make it visible system-wide, then the last warp to get here raises the neighbours' flags
          __threadfence_system();
          __syncwarp();
          if (lane == 0) {
            const unsigned need = (unsigned)(nstrips * CW);
            if (ya < 2) {
              if (atomicAdd(&a.push_count[0], 1u) == need - 1) {
                a.push_count[0] = 0;
                __threadfence_system();
                st_release_sys(a.flag_down, seq);
              }
            }
            if (yb > Y - 2) {
              if (atomicAdd(&a.push_count[1], 1u) == need - 1) {
                a.push_count[1] = 0;
                __threadfence_system();
                st_release_sys(a.flag_up, seq);
              }
            }
          }
        }
        return true;
      };
#pragma unroll 1
      for (;;) {
        if (!row_step(std::integral_constant<int, 0>())) break;
        if (!row_step(std::integral_constant<int, 1>())) break;
        if (!row_step(std::integral_constant<int, 2>())) break;
      }
    }
    if (tracing) {
      // per CTA: [0] SM, [1] producer out of work, [2] items, [3] stages, [4] start, [5] consumers done, [6] after the
      // grid-wide sum (last block only: the step's effective end)
      unsigned long long* tr = a.trace + 24 * (size_t)blockIdx.x;
      unsigned long long tn;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tn));
      if (threadIdx.x == CW * 32) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        tr[0] = smid;
        tr[1] = tn;
        tr[2] = (unsigned long long)items_done;
        tr[3] = (unsigned long long)j;
      }
      if (threadIdx.x == 0) {
        tr[4] = t_start;
        tr[5] = tn;
      }
      if (lane == 0) tr[16 + warp] = tn;  // when each warp left its loop
    }

    // ---- end of the step.  Two things have to happen before anybody starts the next one: the six sums must be
    // known on every rank, and every CTA's stores must have landed.  A block barrier waits for the CTA's outstanding
    // stores (~6 us at the end of a step: the memory system is draining at full write bandwidth), so the sums do NOT
    // go through one: consumer warps reduce by shuffles, hand their six numbers to the producer's lane through shared
    // memory and an mbarrier, and that lane -- which has no stores of its own in flight -- posts the block's partials
    // and takes a ticket.  The block that took the step's FIRST ticket (long finished, its stores long drained) collects
    // all partials, completes the sum over ranks and evaluates the recurrence while the late blocks drain; it publishes
    // once both counters are full.  Fixed shuffle tree, fixed warp / block / rank order: run-to-run reproducible.
    const unsigned tgt = (unsigned)(n + 1) * gridDim.x;  // both counters are zeroed by the host before the launch
    if (warp < CW) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc[i] += shfl_xor_d(acc[i], m);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) s_wpart[warp][i] = acc[i];
        mbar_arrive(rbar);
      }
    } else if (lane == 0) {
      mbar_wait(rbar, n & 1);
      double bsum[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {
        bsum[i] = 0.0;
        for (int w = 0; w < CW; w++) bsum[i] += s_wpart[w][i];
      }
#pragma unroll
      for (int i = 0; i < 6; i++) a.red.partials[i * MAX_PARTIAL_BLOCKS + blockIdx.x] = bsum[i];
      __threadfence();
      const unsigned tA = atomicAdd(a.counters + 1, 1u);
      s_ctl[2] = (tA == (unsigned)n * gridDim.x) ? 1 : 0;  // the first block to get here finishes the step
    }
    __syncthreads();  // every thread of this CTA is past its last store of the step
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.counters + 2, 1u);  // ... and they are visible device-wide
    }
    if (s_ctl[2]) {
      if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_gpu_u32(a.counters + 1) < tgt) {
          __nanosleep(20);
          if (a.wait.budget > 0 && clock64() - t0 > a.wait.budget) __trap();
        }
      }
      __syncthreads();
      double fin[6];
#pragma unroll
      for (int i = 0; i < 6; i++) fin[i] = 0.0;
      for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += blockDim.x) {
#pragma unroll
        for (int i = 0; i < 6; i++) fin[i] += __ldcg(&a.red.partials[i * MAX_PARTIAL_BLOCKS + bb]);
      }
      block_sum<6>(fin, s_red);
      double total[6];
      if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) total[i] = fin[i];
        if (a.step_ns != nullptr && 2 * step + 1 < a.step_ns_cap) a.step_ns[2 * step] = gtime();
      }
      if (a.pr.seq != 0) {  // every thread of this block: one warp per peer
        P2PRed pr = a.pr;
        pr.seq += (unsigned long long)step;
        p2p_allreduce_block(pr, total, 6);
      }
      if (threadIdx.x == 0) {
        if (a.step_ns != nullptr && 2 * step + 1 < a.step_ns_cap) a.step_ns[2 * step + 1] = gtime();  // after the rank sum
        const double rr = total[0];
        // step 0: the set-up pass (alpha = beta = 0), step i >= 1: reference iteration k = i-1
        if (step > 0) {
          const int k = step - 1;
          st->rsq_new = rr;
          st->iter = k + 1;
          if (a.hist != nullptr && k < __ldcg(&st->hist_cap)) a.hist[k] = rr;
          const double want = __ldcg(&st->rsq_pred);  // how good was the prediction that went into beta (diagnostic)
          if (rr > 0.0) {
            const double e = fabs(want - rr) / rr;
            if (e > __ldcg(&st->pred_err)) st->pred_err = e;
          }
          const bool conv = sqrt(rr) < __ldcg(&st->eps) * __ldcg(&st->bnorm);  // generic_cg.cpp:339
          const bool last = (k == __ldcg(&st->max_iter) - 1);
          if (conv || last) {
            st->done = 1;
            st->hit_max = last ? 1 : 0;  // generic_cg.cpp:356 tests k alone
          }
        }
        // scalars of the next step: alpha = rsq/<p,Ap> (generic_cg.cpp:326), beta = rsqNew/rsq (:344) with the
        // numerator predicted:  |r - alpha q|^2 = |r|^2 - 2 Re(alpha <r,q>) + |alpha|^2 |q|^2
        const cplx al = cdiv(mk(rr, 0.0), mk(total[1], total[2]));
        const cplx arq = fmul(al, mk(total[3], total[4]));
        double pred = rr - 2.0 * arq.x + (al.x * al.x + al.y * al.y) * total[5];
        if (!(pred > 0.0)) pred = 0.0;  // converged to rounding: restart the direction (beta = 0)
        st->alpha_re = al.x;
        st->alpha_im = al.y;
        st->beta = (rr > 0.0) ? pred / rr : 0.0;
        st->rsq_pred = pred;
        st->rsq_old = rr;
        st->pAp_re = total[1];
        st->pAp_im = total[2];
        // every block's stores of this step must be in before anybody reads them
        const long long t0 = clock64();
        while (ld_acquire_gpu_u32(a.counters + 2) < tgt) {
          if (a.wait.budget > 0 && clock64() - t0 > a.wait.budget) __trap();
        }
        if (a.queue != nullptr) *a.queue = 0;  // every block has claimed its last item: ready for the next step
        if (tracing) a.trace[24 * (size_t)blockIdx.x + 6] = gtime();
        __threadfence();
        st_release_gpu_s32(&st->step, step + 1);  // every block of this launch is waiting for this word
      }
    }
  }
}

// boundary rows of (r, q, p) -> the neighbours' ghost rows, before the first step of a solve
__global__ void __launch_bounds__(256) cg_step_halo_init_kernel(const cplx* r, const cplx* q, const cplx* p, int X, int Y,
                                                                cplx* push_down, cplx* push_up,
                                                                unsigned long long* flag_down, unsigned long long* flag_up,
                                                                unsigned long long seq, unsigned int* ticket) {
  const size_t n = (size_t)2 * X;  // two rows
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 6 * n; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i / (2 * n));          // 0 r, 1 q, 2 p
    const size_t w = i % (2 * n);
    const bool up = w >= n;
    const size_t e = up ? w - n : w;           // element inside the two rows
    const cplx* src = (v == 0 ? r : (v == 1 ? q : p)) + (up ? (size_t)(Y - 2) * X : 0) + e;
    (up ? push_up : push_down)[(size_t)v * n + e] = *src;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    if (t == gridDim.x - 1) {
      __threadfence_system();
      st_release_sys(flag_down, seq);
      st_release_sys(flag_up, seq);
    }
  }
}

static unsigned int* g_cs_counters = nullptr;  // [0] item counter of the dynamic schedule, [1] blocks whose sums are
                                               // posted, [2] blocks whose stores are in (one set per process / device)
static int g_cgstep_enabled = -1;
static int g_cgstep_variant = -1;   // rows_per_item * 1000 + 100 * CW + 10 * STAGES + MINB
static int g_cgstep_persist = 1;     // GLB_CGSTEP_PERSIST=0: one launch per CG iteration instead of one per solve
static const char* g_trace_path = nullptr;
static int g_trace_launch = 0;

static double g_gfac = 0.0;  // GLB_CGSTEP_GFAC: guided schedule, first item height = rows / (CTAs per strip * gfac); 0 = by slab height
struct RowTable {
  int Y, hmin;
  double per_strip;
  int nrb;
  int* d_rows;
};
static std::vector<RowTable> g_row_tables;
static const RowTable* row_table(int Y, int hmin, double per_strip) {
  for (const RowTable& t : g_row_tables)
    if (t.Y == Y && t.hmin == hmin && t.per_strip == per_strip) return &t;
  std::vector<int> rows;
  rows.push_back(0);
  int y = 0;
  while (y < Y) {
    int h = (int)((Y - y) / per_strip);
    if (h < hmin) h = hmin;
    if (Y - y - h < hmin / 2) h = Y - y;  // do not leave a sliver
    if (h > Y - y) h = Y - y;
    y += h;
    rows.push_back(y);
  }
  RowTable t{Y, hmin, per_strip, (int)rows.size() - 1, nullptr};
  if (cudaMalloc((void**)&t.d_rows, sizeof(int) * rows.size()) != cudaSuccess) return nullptr;
  if (cudaMemcpy(t.d_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  g_row_tables.push_back(t);
  return &g_row_tables.back();
}

template <int CW, int STAGES, int MINB>
static int launch_cg_step_t(glb_operator* op, CgStepArgs a, int rows_per_item) {
  glb_context* ctx = op->ctx;
  constexpr int OUT = CW * 28, TW = OUT + 4;
  auto kern = cg_step_kernel<CW, STAGES, MINB>;
  const size_t smem = (size_t)STAGES * CS_NARR * TW * sizeof(cplx) + 2 * STAGES * sizeof(unsigned long long) +
                      STAGES * sizeof(int4);
  static int per_sm = 0;
  if (per_sm == 0) {
    if (smem + 4096 > 48 * 1024)
      GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (CW + 1) * 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const long long nstrips = (a.X + OUT - 1) / OUT;
  const long long max_ctas = (long long)ctx->sm_count * per_sm;
  long long nrb;
  a.rb_rows = nullptr;
  if (rows_per_item >= 500) {
    // guided schedule: row blocks shrink from (rows left) / (CTAs per strip * gfac) down to hmin as the sweep nears
    // the top, so the last items are short (small tail) while most rows are in tall items (little halo re-reading)
    const int hmin = rows_per_item - 500 > 4 ? rows_per_item - 500 : 4;
    // measured on the B200 (profiles/r02_tune_cgstep.md): gfac 1.0 is best up to 2048 rows, 1.25 - 1.5 at 4096
    double gfac = g_gfac;
    if (gfac <= 0.0) gfac = a.Y <= 2048 ? 1.0 : (a.Y >= 4096 ? 1.25 : 1.0 + 0.25 * (a.Y - 2048) / 2048.0);
    const double per_strip = (double)max_ctas / (double)nstrips * gfac;
    const RowTable* t = row_table(a.Y, hmin, per_strip);
    if (!t) return fail(GLB_ERR_CUDA, "cg_step: row table allocation failed");
    a.rb_rows = t->d_rows;
    nrb = t->nrb;
  } else if (rows_per_item > 0) {
    // dynamic schedule: many more items than CTAs, claimed from a counter as CTAs run out of work
    nrb = (a.Y + rows_per_item - 1) / rows_per_item;
  } else {
    nrb = 0;
  }
  if (rows_per_item > 0) {
    a.queue = (unsigned int*)1;  // set below
  }
  if (rows_per_item <= 0) {
    nrb = max_ctas / nstrips;  // static: one item per resident CTA
    const long long nrb_cap = a.Y >= 16 ? a.Y / 8 : 1;  // at least 8 rows per item (4 halo rows each)
    if (nrb > nrb_cap) nrb = nrb_cap;
    a.queue = nullptr;
  }
  if (nrb < 1) nrb = 1;
  if (g_cs_counters == nullptr) GLB_CUDA(cudaMalloc((void**)&g_cs_counters, 4 * sizeof(unsigned int)));
  GLB_CUDA(cudaMemsetAsync(g_cs_counters, 0, 4 * sizeof(unsigned int), ctx->stream));  // counted up from 0 by every launch
  a.counters = g_cs_counters;
  if (a.queue != nullptr) a.queue = g_cs_counters;
  a.nstrips = (int)nstrips;
  a.nrb = (int)nrb;
  long long blocks = nstrips * nrb;
  if (blocks > max_ctas) blocks = max_ctas;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  // GLB_CGSTEP_TRACE=<file>: per-CTA record (SM, start, end, items) of step 20 -- where does the tail go?
  unsigned long long* d_trace = nullptr;
  a.trace = nullptr;
  a.trace_step = -1;
  if (g_trace_path != nullptr && (a.nsteps > 1 ? g_trace_launch++ == 0 : ++g_trace_launch == 20)) {
    GLB_CUDA(cudaMalloc((void**)&d_trace, sizeof(unsigned long long) * 24 * blocks));
    GLB_CUDA(cudaMemset(d_trace, 0, sizeof(unsigned long long) * 24 * blocks));
    a.trace = d_trace;
    a.trace_step = 20;
    if (a.nsteps > 1) {  // and when every step of this launch ended
      a.step_ns_cap = 8192;  // two stamps per step: local sums done, sum over ranks done
      GLB_CUDA(cudaMalloc((void**)&a.step_ns, sizeof(unsigned long long) * a.step_ns_cap));
      GLB_CUDA(cudaMemset(a.step_ns, 0, sizeof(unsigned long long) * a.step_ns_cap));
    }
  }
  {
    ProfScope prof(ctx, PROF_CG_STEP);
    if (a.nsteps > 1) {
      // persistent: every CTA waits for the last block of each step, so all of them must be resident
      void* kargs[] = {(void*)&a};
      GLB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)blocks), dim3((CW + 1) * 32), kargs, smem,
                                           ctx->stream));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
      kern<<<(unsigned)blocks, (CW + 1) * 32, smem, ctx->stream>>>(a);
      GLB_LAUNCH_CHECK();
    }
  }
  if (d_trace != nullptr) {
    std::vector<unsigned long long> h(24 * blocks);
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    GLB_CUDA(cudaMemcpy(h.data(), d_trace, sizeof(unsigned long long) * 24 * blocks, cudaMemcpyDeviceToHost));
    cudaFree(d_trace);
    if (a.step_ns != nullptr) {
      std::vector<unsigned long long> hs(a.step_ns_cap);
      GLB_CUDA(cudaMemcpy(hs.data(), a.step_ns, sizeof(unsigned long long) * a.step_ns_cap, cudaMemcpyDeviceToHost));
      cudaFree(a.step_ns);
      if (FILE* f = fopen((std::string(g_trace_path) + ".steps").c_str(), "w")) {
        fprintf(f, "# step t_local_sums_done_ns t_rank_sum_done_ns\n");
        for (int i = 0; 2 * i + 1 < a.step_ns_cap && hs[2 * i + 1] != 0; i++)
          fprintf(f, "%d %llu %llu\n", i, hs[2 * i], hs[2 * i + 1]);
        fclose(f);
      }
    }
    if (FILE* f = fopen(g_trace_path, "w")) {
      fprintf(f, "# block smid t_start t_producer_end items stages t_consumers_end t_after_gridsum t_blocksum t_fence "
                 "t_ticket t_partials t_updated t_next_start t_published t_warp0..4_left_loop   (ns; X=%d Y=%d nstrips=%d nrb=%d blocks=%lld)\n",
              a.X, a.Y, a.nstrips, a.nrb, blocks);
      for (long long b = 0; b < blocks; b++) {
        const unsigned long long* t = &h[24 * b];
        fprintf(f, "%lld %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu\n", b,
                t[0], t[4], t[1], t[2], t[3], t[5], t[6], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[16], t[17],
                t[18], t[19], t[20]);
      }
      fclose(f);
    }
  }
  return GLB_OK;
}

static void cg_step_env() {
  if (g_cgstep_enabled < 0) {
    const char* e = getenv("GLB_CGSTEP");
    g_cgstep_enabled = (e && atoi(e) == 0) ? 0 : 1;
    const char* v = getenv("GLB_CGSTEP_VARIANT");
    g_cgstep_variant = v ? atoi(v) : 508433;  // guided schedule down to 8-row items, 4 consumer warps, 3 stages, 3 CTAs/SM
    g_trace_path = getenv("GLB_CGSTEP_TRACE");  // on slabs give every rank its own file (the launcher's job)
    const char* pe = getenv("GLB_CGSTEP_PERSIST");
    g_cgstep_persist = (pe && atoi(pe) == 0) ? 0 : 1;
    if (const char* gf = getenv("GLB_CGSTEP_GFAC")) g_gfac = atof(gf) > 0.0 ? atof(gf) : 0.0;
  }
}

// GLB_CGSTEP=0 switches the single-kernel iteration off (the two-kernel loop of cg.cu runs instead);
// GLB_CGSTEP_VARIANT = 1000*rows_per_item + 100*CW + 10*STAGES + MINB picks the schedule (rows_per_item = 0: static,
// one item per resident CTA) and an instantiated shape.
bool cg_step_ok(const glb_operator* op) {
  cg_step_env();
  if (!g_cgstep_enabled || !normal_fused_ok(op) || op->dtype != GLB_COMPLEX) return false;
  if (op->ctx->nranks > 1 && (!comm_p2p(op->ctx) || op->Yloc < 4)) return false;
  return true;
}

bool cg_step_persistent() {
  cg_step_env();
  return g_cgstep_persist != 0;
}

int launch_cg_step(glb_operator* op, const CgStepArgs& a) {
  cg_step_env();
  const int rows = g_cgstep_variant / 1000;
  switch (g_cgstep_variant % 1000) {
    case 423: return launch_cg_step_t<4, 2, 3>(op, a, rows);
    case 443: return launch_cg_step_t<4, 4, 3>(op, a, rows);
    case 632: return launch_cg_step_t<6, 3, 2>(op, a, rows);
    case 642: return launch_cg_step_t<6, 4, 2>(op, a, rows);
    case 532: return launch_cg_step_t<5, 3, 2>(op, a, rows);
    case 732: return launch_cg_step_t<7, 3, 2>(op, a, rows);
    case 333: return launch_cg_step_t<3, 3, 3>(op, a, rows);
    case 334: return launch_cg_step_t<3, 3, 4>(op, a, rows);
    default: return launch_cg_step_t<4, 3, 3>(op, a, rows);
  }
}

int launch_cg_step_halo_init(glb_operator* op, const void* r, const void* q, const void* p, const CgStepArgs& a,
                             unsigned long long seq, unsigned int* ticket) {
  glb_context* ctx = op->ctx;
  const size_t gvec = (size_t)2 * op->X;
  const size_t wr = (size_t)(seq & 1) * 2 * (3 * gvec);
  const int grid = 16;
  cg_step_halo_init_kernel<<<grid, 256, 0, ctx->stream>>>((const cplx*)r, (const cplx*)q, (const cplx*)p, op->X, op->Yloc,
                                                          a.peer_down + wr + 3 * gvec, a.peer_up + wr, a.flag_down,
                                                          a.flag_up, seq, ticket);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

}  // namespace glb

// Select the CG loop of glb_cg_solve on the one-pass staggered D^dag D operator: on != 0 the single-kernel
// iteration (default), 0 the two-kernel loop; variant > 0 picks an instantiated kernel shape (100*consumer warps +
// 10*stages + blocks per SM), 0 keeps the current one.  Returns the previous on/off setting.
extern "C" int glb_cg_step_mode(int on, int variant) {
  glb::cg_step_env();
  const int prev = glb::g_cgstep_enabled;
  glb::g_cgstep_enabled = on ? 1 : 0;
  glb::g_cgstep_persist = (on == 2) ? 0 : 1;  // on = 2: single-kernel iteration, one launch per iteration
  if (variant > 0) glb::g_cgstep_variant = variant;
  return prev;
}
