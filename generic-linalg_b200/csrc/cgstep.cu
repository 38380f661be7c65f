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
// kernel's last block completes the sum of the six scalars over ranks itself (p2p_allreduce_warp):
// one rank-wide synchronisation per CG iteration, no separate halo or reduction launch.
#include <cstdlib>
#include <type_traits>

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

// arrays of one stage, in this order
enum { CS_R = 0, CS_Q = 1, CS_P = 2, CS_XV = 3, CS_UX = 4, CS_UY = 5, CS_NARR = 6 };

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

  CgState* const st = a.st;
  if (st->done) return;  // (uniform: every thread reads the same word; rank-summed scalars are equal on all ranks)
  const cplx alpha = mk(st->alpha_re, st->alpha_im);
  const cplx nalpha = fneg(alpha);
  const double beta = st->beta;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int X = a.X, Y = a.Y;
  const int nstrips = a.nstrips, nrb = a.nrb;
  const int nitems = nstrips * nrb;
  const bool slab = (a.g_lo != nullptr);
  double acc[6];  // |r|^2, <p,q>.re, <p,q>.im, <r,q>.re, <r,q>.im, |q|^2
#pragma unroll
  for (int i = 0; i < 6; i++) acc[i] = 0.0;

  unsigned j = 0;  // running stage counter: the same sequence in the producer and in every consumer
  if (warp == CW) {
    // ================================================================== producer (one elected lane)
    if (lane == 0) {
      bool waited = false;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int strip = item % nstrips, rb = item / nstrips;
        const int ya = (int)((long long)Y * rb / nrb), yb = (int)((long long)Y * (rb + 1) / nrb);
        if (ya >= yb) continue;
        const int x_lo = strip * OUT - 2;
        const int pos0 = ((x_lo % X) + X) % X;
        if (slab && !waited && (ya - 2 < 0 || yb + 2 > Y) && a.wait.seq != 0) {
          // the neighbours' rows of the previous step must have landed before the TMA unit reads them
          spin_until(a.wait.flag_lo, a.wait.seq, a.wait.budget);
          spin_until(a.wait.flag_hi, a.wait.seq, a.wait.budget);
          asm volatile("fence.proxy.async;" ::: "memory");
          waited = true;
        }
        for (int L = ya - 2; L < yb + 2; L++, j++) {
          const int s = j % STAGES;
          mbar_wait(&empty[s], ((j / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          cplx* const dst = tiles + (size_t)s * (CS_NARR * TW);
          // source rows: vectors at row L (periodic on one rank, ghost rows on slabs), links at row L-1
          const cplx* src[CS_NARR];
          if (slab && L < 0) {
            const cplx* g = a.g_lo + (size_t)(L + 2) * X;
            src[CS_R] = g;
            src[CS_Q] = g + (size_t)2 * X;
            src[CS_P] = g + (size_t)4 * X;
            src[CS_XV] = a.x;  // never used on ghost rows
          } else if (slab && L >= Y) {
            const cplx* g = a.g_hi + (size_t)(L - Y) * X;
            src[CS_R] = g;
            src[CS_Q] = g + (size_t)2 * X;
            src[CS_P] = g + (size_t)4 * X;
            src[CS_XV] = a.x;
          } else {
            const size_t o = (size_t)(((L % Y) + Y) % Y) * X;
            src[CS_R] = a.r_in + o;
            src[CS_Q] = a.q_in + o;
            src[CS_P] = a.p_in + o;
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
    }
  } else {
    // ================================================================== consumers
    const bool eta_neg = (lane & 1);  // tiles and warp windows start on even x
    const int idx = warp * OUT_W + lane;              // this lane's site inside the tile
    const int idx_l = idx > 0 ? idx - 1 : 0;          // its left neighbour (lane 0 of warp 0 never uses it)
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int strip = item % nstrips, rb = item / nstrips;
      const int ya = (int)((long long)Y * rb / nrb), yb = (int)((long long)Y * (rb + 1) / nrb);
      if (ya >= yb) continue;
      const int x = strip * OUT - 2 + idx;
      const bool active = (lane >= 2) && (lane < 30) && (x < X);
      // register windows, slot = (L - (ya-2)) % 3 : at step L slot K2 takes row L, K1 holds L-1, K0 holds L-2
      cplx p[3], r[3], t[3], ux[3], uy[3], uxl[3];
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = r[k] = t[k] = ux[k] = uy[k] = uxl[k] = mk(0.0, 0.0);

      auto row_step = [&](auto Kc, const int L) {
        constexpr int K0 = decltype(Kc)::value % 3, K1 = (K0 + 1) % 3, K2 = (K0 + 2) % 3;
        const int s = j % STAGES;
        mbar_wait(&full[s], (j / STAGES) & 1);
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
        // the streaming part of the iteration at row L (generic_cg.cpp:328-351)
        const cplx rn = fadd(ro, fmul(nalpha, qo));       // r = r - alpha*Ap
        const cplx pn = fadd(rn, fscale(beta, po));       // p = r + beta*p
        r[K2] = rn;
        p[K2] = pn;
        const bool own = active && (L >= ya) && (L < yb);
        if (own) {
          const size_t o = (size_t)L * X + x;
          a.r_out[o] = rn;
          a.p_out[o] = pn;
          a.x[o] = fadd(xo, fmul(alpha, po));             // phi = phi + alpha*p
          acc[0] += fnorm(rn);
          if (slab) {
            if (L < 2) {
              cplx* g = a.push_down + (size_t)L * X + x;
              g[0] = rn;
              g[(size_t)4 * X] = pn;
            }
            if (L >= Y - 2) {
              cplx* g = a.push_up + (size_t)(L - (Y - 2)) * X + x;
              g[0] = rn;
              g[(size_t)4 * X] = pn;
            }
          }
        }
        // t(L-1) = D p at row L-1 ; q(L-2) = D^dag t at row L-2   (operators.cpp:444-453, one pass)
        t[K2] = cs_row<false>(eta_neg, p[K0], p[K1], p[K2], ux[K2], uxl[K2], uy[K2], uy[K1], a.mass);
        const cplx res = cs_row<true>(eta_neg, t[K0], t[K1], t[K2], ux[K1], uxl[K1], uy[K1], uy[K0], a.mass);
        const int y = L - 2;
        if (active && y >= ya) {
          const size_t o = (size_t)y * X + x;
          a.q_out[o] = res;
          Field<cplx>::dot_acc(acc + 1, p[K0], res);
          Field<cplx>::dot_acc(acc + 3, r[K0], res);
          acc[5] += fnorm(res);
          if (slab) {
            if (y < 2) a.push_down[(size_t)(2 + y) * X + x] = res;
            if (y >= Y - 2) a.push_up[(size_t)(2 + y - (Y - 2)) * X + x] = res;
          }
        }
      };
      int L = ya - 2;
#pragma unroll 1
      for (; L + 3 <= yb + 2; L += 3) {
        row_step(std::integral_constant<int, 0>(), L);
        row_step(std::integral_constant<int, 1>(), L + 1);
        row_step(std::integral_constant<int, 2>(), L + 2);
      }
      if (L < yb + 2) row_step(std::integral_constant<int, 0>(), L);
      if (L + 1 < yb + 2) row_step(std::integral_constant<int, 1>(), L + 1);

      if (slab && (ya < 2 || yb > Y - 2)) {
        // this warp's share of the slab's boundary rows is on its way to the neighbours: make it visible
        // system-wide, then the last warp to get here raises the neighbours' flags
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
          const unsigned need = (unsigned)(nstrips * CW);
          if (ya < 2) {
            if (atomicAdd(&a.push_count[0], 1u) == need - 1) {
              a.push_count[0] = 0;
              __threadfence_system();
              st_release_sys(a.flag_down, a.push_seq);
            }
          }
          if (yb > Y - 2) {
            if (atomicAdd(&a.push_count[1], 1u) == need - 1) {
              a.push_count[1] = 0;
              __threadfence_system();
              st_release_sys(a.flag_up, a.push_seq);
            }
          }
        }
      }
    }
  }

  // ---- the six sums: block -> grid (last block) -> ranks (its first warp, peer memory) -> recurrence
  double total[6];
  if (!grid_sum<6>(acc, a.red, total)) return;
  if (a.pr.seq != 0 && threadIdx.x < 32) p2p_allreduce_warp(a.pr, total, 6);
  if (threadIdx.x == 0) {
    const double rr = total[0];
    const int step = st->step;  // 0: the set-up pass (alpha = beta = 0), i >= 1: reference iteration k = i-1
    if (step > 0) {
      const int k = step - 1;
      st->rsq_new = rr;
      st->iter = k + 1;
      if (a.hist != nullptr && k < st->hist_cap) a.hist[k] = rr;
      const double want = st->rsq_pred;  // how good was the prediction that went into beta (diagnostic)
      if (rr > 0.0) {
        const double e = fabs(want - rr) / rr;
        if (e > st->pred_err) st->pred_err = e;
      }
      const bool conv = sqrt(rr) < st->eps * st->bnorm;  // generic_cg.cpp:339
      const bool last = (k == st->max_iter - 1);
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
    st->step = step + 1;
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

template <int CW, int STAGES, int MINB>
static int launch_cg_step_t(glb_operator* op, CgStepArgs a) {
  glb_context* ctx = op->ctx;
  constexpr int OUT = CW * 28, TW = OUT + 4;
  auto kern = cg_step_kernel<CW, STAGES, MINB>;
  const size_t smem = (size_t)STAGES * CS_NARR * TW * sizeof(cplx) + 2 * STAGES * sizeof(unsigned long long);
  static int per_sm = 0;
  if (per_sm == 0) {
    if (smem + 4096 > 48 * 1024)
      GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (CW + 1) * 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const long long nstrips = (a.X + OUT - 1) / OUT;
  const long long max_ctas = (long long)ctx->sm_count * per_sm;
  long long nrb = max_ctas / nstrips;
  const long long nrb_cap = a.Y >= 16 ? a.Y / 8 : 1;  // at least 8 rows per item (4 halo rows each)
  if (nrb > nrb_cap) nrb = nrb_cap;
  if (nrb < 1) nrb = 1;
  a.nstrips = (int)nstrips;
  a.nrb = (int)nrb;
  long long blocks = nstrips * nrb;
  if (blocks > max_ctas) blocks = max_ctas;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  ProfScope prof(ctx, PROF_CG_STEP);
  kern<<<(unsigned)blocks, (CW + 1) * 32, smem, ctx->stream>>>(a);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

// GLB_CGSTEP=0 switches the single-kernel iteration off (the two-kernel loop of cg.cu runs instead);
// GLB_CGSTEP_VARIANT = 100*CW + 10*STAGES + MINB picks an instantiated shape (default 443).
static int g_cgstep_enabled = -1;
static int g_cgstep_variant = -1;
bool cg_step_ok(const glb_operator* op) {
  if (g_cgstep_enabled < 0) {
    const char* e = getenv("GLB_CGSTEP");
    g_cgstep_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!g_cgstep_enabled || !normal_fused_ok(op) || op->dtype != GLB_COMPLEX) return false;
  if (op->ctx->nranks > 1 && (!comm_p2p(op->ctx) || op->Yloc < 4)) return false;
  return true;
}

int launch_cg_step(glb_operator* op, const CgStepArgs& a) {
  if (g_cgstep_variant < 0) {
    const char* e = getenv("GLB_CGSTEP_VARIANT");
    g_cgstep_variant = e ? atoi(e) : 443;
  }
  switch (g_cgstep_variant) {
    case 433: return launch_cg_step_t<4, 3, 3>(op, a);
    case 453: return launch_cg_step_t<4, 5, 3>(op, a);
    case 463: return launch_cg_step_t<4, 6, 3>(op, a);
    case 444: return launch_cg_step_t<4, 4, 4>(op, a);
    case 442: return launch_cg_step_t<4, 4, 2>(op, a);
    case 842: return launch_cg_step_t<8, 4, 2>(op, a);
    case 841: return launch_cg_step_t<8, 4, 1>(op, a);
    case 243: return launch_cg_step_t<2, 4, 3>(op, a);
    case 246: return launch_cg_step_t<2, 4, 6>(op, a);
    default: return launch_cg_step_t<4, 4, 3>(op, a);
  }
}

int launch_cg_step_halo_init(glb_operator* op, const void* r, const void* q, const void* p, const CgStepArgs& a,
                             unsigned int* ticket) {
  glb_context* ctx = op->ctx;
  const int grid = 16;
  cg_step_halo_init_kernel<<<grid, 256, 0, ctx->stream>>>((const cplx*)r, (const cplx*)q, (const cplx*)p, op->X, op->Yloc,
                                                          a.push_down, a.push_up, a.flag_down, a.flag_up, a.push_seq,
                                                          ticket);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

}  // namespace glb

// Select the CG loop of glb_cg_solve on the one-pass staggered D^dag D operator: on != 0 the single-kernel
// iteration (default), 0 the two-kernel loop; variant > 0 picks an instantiated kernel shape (100*consumer warps +
// 10*stages + blocks per SM), 0 keeps the current one.  Returns the previous on/off setting.
extern "C" int glb_cg_step_mode(int on, int variant) {
  if (glb::g_cgstep_enabled < 0) {
    const char* e = getenv("GLB_CGSTEP");
    glb::g_cgstep_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int prev = glb::g_cgstep_enabled;
  glb::g_cgstep_enabled = on ? 1 : 0;
  if (variant > 0) glb::g_cgstep_variant = variant;
  return prev;
}
