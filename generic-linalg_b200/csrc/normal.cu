// normal.cu -- D^dagger D of the 2-D U(1) staggered operator in ONE pass over HBM.
//
// The reference forms the normal operator through a temporary (operators.cpp:444-453:
// tmp = D psi ; out = D^dag tmp), i.e. two full stencil passes: 128 B/site.  Here the temporary
// never leaves the SM: a warp streams over the rows of a 64-site window keeping psi(y), psi(y+1)
// and t(y-1), t(y) (t = D psi) in registers; each step loads psi(y+2) and the links of row y+1,
// forms t(y+1) and then out(y) = D^dag t.  Traffic: 16 psi + 32 links + 16 out = 64 B/site
// (96 B/site with the fused CG direction update p = r + beta p_old, which also writes p).
//
//   * warp-tiled with a 2-site overlap on each side: a warp loads 64 consecutive sites (one
//     32-byte LDG.256 per lane per array) but produces the 60 inner ones; every x neighbour of
//     both stencil applications is then a warp shuffle -- no edge loads, no shared memory, no
//     block-level synchronisation.  The overlapped sites hit L1/L2, DRAM traffic is unchanged.
//   * persistent grid, (strip,row) units split evenly over all warps; next row prefetched.
//   * t is rounded to double exactly where the reference stores tmp, and both stencils evaluate
//     the reference's expression order without FMA contraction: results are bit-identical to
//     square_staggered_normal_u1.
//   * fused epilogue reductions <w,out>, |out|^2 and the device-resident CG hooks as in stencil.cu.
#include <cstdlib>
#include <type_traits>

#include "normal_args.cuh"

namespace glb {

constexpr int NORM_THREADS = 128;
constexpr int NORM_WARPS = NORM_THREADS / 32;
constexpr int NORM_OUT_PER_WARP = 60;  // 64 loaded - 2 halo sites on each side


template <bool FUSE>
struct NormLoad {  // everything fetched one row ahead: psi(y+2) (raw), U(y+1)
  cplx a[2];
  cplx b[2];
  cplx ux[2];
  cplx uy[2];
};

// STAGES == 0: the next row is prefetched into registers.  STAGES >= 2: every warp owns a ring of
// STAGES slots in shared memory filled by cp.async (LDGSTS), STAGES-1 rows in flight per warp and no
// staging registers -- bytes in flight no longer depend on the register-limited occupancy.
template <bool FUSE_XPAY, int NDOT, int STAGES, bool UNROLL3, int LAYOUT>
__global__ void __launch_bounds__(NORM_THREADS, 3) normal_kernel(const NormArgs a) {
  extern __shared__ __align__(32) unsigned char ring_raw[];
  double beta = 0.0;
  if (a.cg != nullptr) {
    if (a.cg->done) return;
    if (FUSE_XPAY) beta = xdiv(a.cg->rsq_new, a.cg->rsq_old);  // generic_cg.cpp:344
  }
  halo_wait_block(a.wait);  // slabs over peer memory: the neighbours' rows must have landed
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;

  const int lane = threadIdx.x & 31;
  const int X = a.X, Y = a.Y;
  const int nstrips = (X + NORM_OUT_PER_WARP - 1) / NORM_OUT_PER_WARP;
  // Work items are laid out so that the warps of a block (and neighbouring blocks) sweep ADJACENT
  // strips over the SAME rows at the same time: the 2-site overlaps then hit L1/L2 instead of DRAM
  // and the chip streams whole contiguous rows (DRAM page locality), like a 2-D tiled sweep.
  const long long nitems = (long long)nstrips * a.nrb;
  const long long nwarps = (long long)gridDim.x * NORM_WARPS;
  const long long wid = (long long)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);

  auto wrap_row = [&](int y) -> size_t { return (size_t)(((y % Y) + Y) % Y) * X; };
  const bool slab = (a.g_lo != nullptr);
  // links carry periodic ghost rows: U(.,y) is addressable for y in [-2, Y+2)
  auto link_row = [&](int y) -> ptrdiff_t { return (ptrdiff_t)y * X; };

  for (long long item = wid; item < nitems; item += nwarps) {
    const int strip = (int)(item % nstrips);
    const int rb = (int)(item / nstrips);
    const int ya = (int)((long long)Y * rb / a.nrb);
    const int yb = (int)((long long)Y * (rb + 1) / a.nrb);
    if (ya >= yb) continue;

    // this lane's pair of sites (periodic in x; pairs never straddle the seam because X is even)
    const int xs = strip * NORM_OUT_PER_WARP - 2 + 2 * lane;
    const int x0 = ((xs % X) + X) % X;
    const bool active = (lane >= 1) && (lane <= 30) && (xs < X);

    auto load_psi = [&](int y, cplx(&v)[2]) {  // the (possibly fused) input row y at this lane's pair
      if (slab && (y < 0 || y >= Y)) {  // ghost rows hold final values
        ldv<2>((y < 0 ? a.g_lo + (size_t)(y + 2) * X : a.g_hi + (size_t)(y - Y) * X) + x0, v);
        return;
      }
      const size_t o = wrap_row(y) + x0;
      if (FUSE_XPAY) {
        cplx rr[2], pp[2];
        ldv<2>(a.r + o, rr);
        ldv<2>(a.pold + o, pp);
        v[0] = fadd(rr[0], fscale(beta, pp[0]));
        v[1] = fadd(rr[1], fscale(beta, pp[1]));
      } else {
        ldv<2>(a.in + o, v);
      }
    };
    auto fetch = [&](int y, NormLoad<FUSE_XPAY>& L) {  // for output row y: psi(y+2), U(y+1)
      const ptrdiff_t o1 = link_row(y + 1) + x0;
      if (slab && y + 2 >= Y) {  // ghost row above the slab: final values; b = 0 makes the fused form a no-op
        ldv<2>(a.g_hi + (size_t)(y + 2 - Y) * X + x0, L.a);
        L.b[0] = mk(0.0, 0.0);
        L.b[1] = mk(0.0, 0.0);
      } else {
        const size_t o2 = wrap_row(y + 2) + x0;
        if (FUSE_XPAY) {
          ldv<2>(a.r + o2, L.a);
          ldv<2>(a.pold + o2, L.b);
        } else {
          ldv<2>(a.in + o2, L.a);
        }
      }
      ldv_nc<2>(a.Ux + o1, L.ux);
      ldv_nc<2>(a.Uy + o1, L.uy);
    };

    // ---- register windows: three-slot rings indexed by (row - ya) % 3, so that unrolling the row
    // loop by three renames registers instead of moving them.  At row y (k = (y - ya) % 3):
    //   psi(y), psi(y+1), psi(y+2) = p[k], p[k+1], p[k+2];  t(y-1), t(y), t(y+1) = t[k], t[k+1], t[k+2]
    //   U_y(y-1), U_y(y), U_y(y+1) = uy[k], uy[k+1], uy[k+2];  U_x(y), U_x(y+1) = ux[k+1], ux[k+2]
    // (slot indices mod 3; uxl = U_x of the site to the left of the pair)
    cplx p[3][2], t[3][2], ux[3][2], uy[3][2];
    cplx uxl[3];
    {  // prologue: t(ya-1), t(ya) from psi(ya-2 .. ya+1)
      cplx p_mm[2], p_m[2], ux_m[2], uy_mm[2];
      load_psi(ya - 2, p_mm);
      load_psi(ya - 1, p_m);
      load_psi(ya, p[0]);
      load_psi(ya + 1, p[1]);
      ldv_nc<2>(a.Uy + link_row(ya - 2) + x0, uy_mm);
      ldv_nc<2>(a.Ux + link_row(ya - 1) + x0, ux_m);
      ldv_nc<2>(a.Uy + link_row(ya - 1) + x0, uy[0]);
      ldv_nc<2>(a.Ux + link_row(ya) + x0, ux[1]);
      ldv_nc<2>(a.Uy + link_row(ya) + x0, uy[1]);
      const cplx uxl_m = shfl_up_c(ux_m[1], 1);
      uxl[1] = shfl_up_c(ux[1][1], 1);
      stag_row<false>(t[0], p_mm, p_m, p[0], ux_m, uxl_m, uy[0], uy_mm, a.mass);     // t(ya-1)
      stag_row<false>(t[1], p_m, p[0], p[1], ux[1], uxl[1], uy[1], uy[0], a.mass);  // t(ya)
    }
    constexpr int NARR = FUSE_XPAY ? 4 : 3;
    constexpr int NST = STAGES > 0 ? STAGES : 1;
    // Ring slot of (stage, array).
    //  LAYOUT 0: every thread owns a private 32-byte slot and copies its own pair of sites (two 16-byte
    //            LDGSTS to adjacent addresses: each warp request touches all 32 sectors of the window).
    //  LAYOUT 1: the warp's 64-site window is stored line by line: lane l copies window sites l and l+32
    //            (consecutive lanes on consecutive 16-byte chunks: 4 whole lines per LDGSTS request, every
    //            sector requested once), each 128-byte line stays one shared-memory row, and odd rows swap
    //            neighbouring chunks (c ^ 1) so that the pair reads (lane l: sites 2l, 2l+1) are
    //            conflict-free LDS.128.
    auto phys = [](int j) -> int { return (j & ~7) | ((j & 7) ^ ((j >> 3) & 1)); };
    cplx* const ring_w = reinterpret_cast<cplx*>(ring_raw) +
                         (LAYOUT == 0 ? (size_t)threadIdx.x * 2 : (size_t)(threadIdx.x >> 5) * 64);
    auto slot = [&](int stage, int arr) -> cplx* { return ring_w + (size_t)(stage * NARR + arr) * (NORM_THREADS * 2); };
    const int win0 = strip * NORM_OUT_PER_WARP - 2;
    // copy c (0/1) of this lane: ring index and x coordinate
    const int ia0 = (LAYOUT == 0) ? 0 : phys(lane), ia1 = (LAYOUT == 0) ? 1 : phys(lane + 32);
    const int xa = (LAYOUT == 0) ? x0 : (((win0 + lane) % X) + X) % X;
    const int xb = (LAYOUT == 0) ? x0 + 1 : (((win0 + 32 + lane) % X) + X) % X;
    // reads: this lane's pair
    const int ir0 = (LAYOUT == 0) ? 0 : phys(2 * lane), ir1 = (LAYOUT == 0) ? 1 : phys(2 * lane + 1);
    // running state of the producer: output row of the next stage, its ring slot, and the (wrapped)
    // input row y+2 it reads -- incremented, never divided
    int is_y = ya, is_st = 0;
    int is_row2 = ya + 2;
    if (!slab && is_row2 >= Y) is_row2 -= Y;
    auto issue = [&]() {  // asynchronous copies for output row is_y: psi(is_y+2), U(is_y+1); always one commit group
      if (is_y < yb) {
        const ptrdiff_t o1 = link_row(is_y + 1);
        if (slab && is_row2 >= Y) {
          const cplx* g = a.g_hi + (size_t)(is_row2 - Y) * X;
          cp_async16(slot(is_st, 0) + ia0, g + xa);
          cp_async16(slot(is_st, 0) + ia1, g + xb);
          if (FUSE_XPAY) {
            slot(is_st, 3)[ia0] = mk(0.0, 0.0);
            slot(is_st, 3)[ia1] = mk(0.0, 0.0);
          }
        } else {
          const size_t o2 = (size_t)is_row2 * X;
          const cplx* pa = (FUSE_XPAY ? a.r : a.in) + o2;
          cp_async16(slot(is_st, 0) + ia0, pa + xa);
          cp_async16(slot(is_st, 0) + ia1, pa + xb);
          if (FUSE_XPAY) {
            cp_async16(slot(is_st, 3) + ia0, a.pold + o2 + xa);
            cp_async16(slot(is_st, 3) + ia1, a.pold + o2 + xb);
          }
        }
        cp_async16(slot(is_st, 1) + ia0, a.Ux + o1 + xa);
        cp_async16(slot(is_st, 1) + ia1, a.Ux + o1 + xb);
        cp_async16(slot(is_st, 2) + ia0, a.Uy + o1 + xa);
        cp_async16(slot(is_st, 2) + ia1, a.Uy + o1 + xb);
        is_y++;
        if (++is_row2 == Y && !slab) is_row2 = 0;
        if (++is_st == NST) is_st = 0;
      }
      cp_async_commit();
    };
    NormLoad<FUSE_XPAY> nxt;
    if (STAGES == 0) {
      fetch(ya, nxt);
    } else {
#pragma unroll
      for (int k = 0; k < (STAGES > 0 ? STAGES - 1 : 0); k++) issue();
    }

    int rd_st = 0;  // consumer's ring slot
    // one output row; K = (y - ya) % 3 selects the register slots at compile time
    auto row_step = [&](auto Kc, const int y) {
      constexpr int K0 = decltype(Kc)::value % 3, K1 = (K0 + 1) % 3, K2 = (K0 + 2) % 3;
      cplx la[2], lb[2];
      if (STAGES == 0) {
        la[0] = nxt.a[0];
        la[1] = nxt.a[1];
        lb[0] = nxt.b[0];
        lb[1] = nxt.b[1];
        ux[K2][0] = nxt.ux[0];
        ux[K2][1] = nxt.ux[1];
        uy[K2][0] = nxt.uy[0];
        uy[K2][1] = nxt.uy[1];
        if (y + 1 < yb) fetch(y + 1, nxt);  // prefetch while this row is computed
      } else {
        if (LAYOUT != 0) __syncwarp();                   // all lanes are done with the slot refilled next
        issue();                                         // keep STAGES-1 rows in flight
        cp_async_wait<(STAGES > 0 ? STAGES - 1 : 0)>();  // this lane's copies of row y have landed
        if (LAYOUT != 0) __syncwarp();                   // ... and so have the other lanes'
        la[0] = slot(rd_st, 0)[ir0];
        la[1] = slot(rd_st, 0)[ir1];
        ux[K2][0] = slot(rd_st, 1)[ir0];
        ux[K2][1] = slot(rd_st, 1)[ir1];
        uy[K2][0] = slot(rd_st, 2)[ir0];
        uy[K2][1] = slot(rd_st, 2)[ir1];
        if (FUSE_XPAY) {
          lb[0] = slot(rd_st, 3)[ir0];
          lb[1] = slot(rd_st, 3)[ir1];
        }
        if (++rd_st == NST) rd_st = 0;
      }
      if (FUSE_XPAY) {
        p[K2][0] = fadd(la[0], fscale(beta, lb[0]));
        p[K2][1] = fadd(la[1], fscale(beta, lb[1]));
      } else {
        p[K2][0] = la[0];
        p[K2][1] = la[1];
      }
      uxl[K2] = shfl_up_c(ux[K2][1], 1);
      cplx res[2];
      stag_row<false>(t[K2], p[K0], p[K1], p[K2], ux[K2], uxl[K2], uy[K2], uy[K1], a.mass);  // t(y+1) = D psi
      stag_row<true>(res, t[K0], t[K1], t[K2], ux[K1], uxl[K1], uy[K1], uy[K0], a.mass);     // out(y) = D^dag t
      if (active) {
        const size_t o = (size_t)y * X + x0;
        stv<2>(a.out + o, res);
        if (FUSE_XPAY) stv<2>(a.pnew + o, p[K0]);
        if (NDOT >= 1) {
          cplx wv[2];
          if (a.w == nullptr) {
            wv[0] = p[K0][0];
            wv[1] = p[K0][1];
          } else {
            ldv<2>(a.w + o, wv);
          }
          Field<cplx>::dot_acc(acc, wv[0], res[0]);
          Field<cplx>::dot_acc(acc, wv[1], res[1]);
        }
        if (NDOT >= 2) {
          acc[2] += fnorm(res[0]);
          acc[2] += fnorm(res[1]);
        }
      }
    };
    if (UNROLL3) {
      int y = ya;
#pragma unroll 1
      for (; y + 3 <= yb; y += 3) {
        row_step(std::integral_constant<int, 0>(), y);
        row_step(std::integral_constant<int, 1>(), y + 1);
        row_step(std::integral_constant<int, 2>(), y + 2);
      }
      if (y < yb) row_step(std::integral_constant<int, 0>(), y);
      if (y + 1 < yb) row_step(std::integral_constant<int, 1>(), y + 1);
    } else {
#pragma unroll 1
      for (int y = ya; y < yb; y++) {
        row_step(std::integral_constant<int, 0>(), y);
#pragma unroll
        for (int s2 = 0; s2 < 2; s2++) {  // roll the windows
          p[0][s2] = p[1][s2];
          p[1][s2] = p[2][s2];
          t[0][s2] = t[1][s2];
          t[1][s2] = t[2][s2];
          uy[0][s2] = uy[1][s2];
          uy[1][s2] = uy[2][s2];
          ux[1][s2] = ux[2][s2];
        }
        uxl[1] = uxl[2];
      }
    }
    if (STAGES > 0) {
      cp_async_wait<0>();
      __syncwarp();  // the next item's prologue refills slots other lanes may still be reading
    }
  }

  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total)) {  // the block that arrived last
      // slab run over peer memory: its first warp finishes the sum over ranks
      if (a.cg != nullptr && a.cg_role == 3) p2p_allreduce_block(a.pr, total, 2);
      if (threadIdx.x == 0 && a.cg != nullptr) {
        if (a.cg_role == 1 || a.cg_role == 3) {  // <p,Ap> ready: generic_cg.cpp:326 / :345
          a.cg->pAp_re = total[0];
          a.cg->pAp_im = total[1];
          a.cg->rsq_old = a.cg->rsq_new;
        } else if (a.cg_role == 2) {  // slab run over NCCL: rank-local part, summed on the stream next
          a.cg->partial[1] = total[0];
          a.cg->partial[2] = total[1];
        }
      }
    }
  }
}

template <bool FUSE, int NDOT, int STAGES, bool UNROLL3, int LAYOUT>
static int launch_normal_t(glb_operator* op, const NormArgs& a) {
  glb_context* ctx = op->ctx;
  auto kern = normal_kernel<FUSE, NDOT, STAGES, UNROLL3, LAYOUT>;
  const size_t smem = (size_t)STAGES * (FUSE ? 4 : 3) * NORM_THREADS * 2 * sizeof(cplx);
  static int per_sm = 0;
  if (per_sm == 0) {
    // static shared memory of the reduction epilogue rides on top of the dynamic ring
    if (smem + 2048 > 48 * 1024)
      GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NORM_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const long long nstrips = (a.X + NORM_OUT_PER_WARP - 1) / NORM_OUT_PER_WARP;
  const long long max_warps = (long long)ctx->sm_count * per_sm * NORM_WARPS;
  // row blocks: as many as there are warps to fill, but at least 8 rows each (4 halo rows per item)
  long long nrb = max_warps / nstrips;
  const long long nrb_cap = a.Y >= 16 ? a.Y / 8 : 1;
  if (nrb > nrb_cap) nrb = nrb_cap;
  if (nrb < 1) nrb = 1;
  NormArgs b = a;
  b.nrb = (int)nrb;
  long long blocks = (nstrips * nrb + NORM_WARPS - 1) / NORM_WARPS;
  if (blocks > (long long)ctx->sm_count * per_sm) blocks = (long long)ctx->sm_count * per_sm;
  if (blocks < 1) blocks = 1;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  ProfScope prof(ctx, FUSE ? PROF_NORMAL_FUSED : PROF_NORMAL);
  kern<<<(unsigned)blocks, NORM_THREADS, smem, ctx->stream>>>(b);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

template <int STAGES, bool UNROLL3, int LAYOUT>
static int launch_normal_s(glb_operator* op, const NormArgs& a, bool fuse, int ndot) {
  if (fuse) {
    if (ndot == 0) return launch_normal_t<true, 0, STAGES, UNROLL3, LAYOUT>(op, a);
    if (ndot == 1) return launch_normal_t<true, 1, STAGES, UNROLL3, LAYOUT>(op, a);
    return launch_normal_t<true, 2, STAGES, UNROLL3, LAYOUT>(op, a);
  }
  if (ndot == 0) return launch_normal_t<false, 0, STAGES, UNROLL3, LAYOUT>(op, a);
  if (ndot == 1) return launch_normal_t<false, 1, STAGES, UNROLL3, LAYOUT>(op, a);
  return launch_normal_t<false, 2, STAGES, UNROLL3, LAYOUT>(op, a);
}

// can the one-pass kernel serve this operator?  (gauged, even X; slabs at least two rows thick)
bool normal_fused_ok(const glb_operator* op) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GLB_NORMAL_FUSED");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return enabled && op->kind == OPK_STAGGERED && (op->flags & GLB_STAG_NORMAL) && op->has_links &&
         (op->X % 2 == 0) && op->X >= 2 && (op->ctx->nranks == 1 || op->Yloc >= 2);
}

int launch_normal(glb_operator* op, void* out, const void* in, const ApplyFusion& f) {
  glb_context* ctx = op->ctx;
  NormArgs a{};
  const bool fuse = (f.r != nullptr);
  if (fuse) {
    a.r = (const cplx*)f.r;
    a.pold = (const cplx*)f.p_old;
    a.pnew = (cplx*)f.p_new;
  } else {
    a.in = (const cplx*)in;
  }
  a.out = (cplx*)out;
  a.Ux = op->Ux;
  a.Uy = op->Uy;
  a.w = f.w_is_input ? nullptr : (const cplx*)f.w;
  a.X = op->X;
  a.Y = op->Yloc;
  if (ctx->nranks > 1) {
    a.g_lo = (const cplx*)op->ghost_lo;
    a.g_hi = (const cplx*)op->ghost_hi;
  }
  a.mass = op->mass;
  a.red = ctx->red;
  if (!f.to_host) a.red.result_host = nullptr;
  a.cg = (CgState*)f.cg_state;
  a.cg_role = f.cg_role;
  a.pr = f.pr;
  a.wait = f.wait;
  const int ndot = (f.w != nullptr || f.w_is_input) ? (f.want_norm ? 2 : 1) : 0;
  // ring depth (measured at 4096^2, profiles/): the fused-direction variant streams 4 arrays and gains
  // ~8 % from a 4-deep cp.async ring; the plain variant (3 arrays) is best with register prefetch.
  static int stages_fused = -1, stages_plain = -1;
  if (stages_fused < 0) {
    const char* e = getenv("GLB_NORMAL_STAGES");
    stages_fused = e ? atoi(e) : 4;
    const char* e2 = getenv("GLB_NORMAL_STAGES_PLAIN");
    stages_plain = e2 ? atoi(e2) : 0;
  }
  {
    // one site per thread (normal1.cu): GLB_NORMAL_SPT1 = 10*stages + min blocks per SM, 0 = off
    // Measured at 4096^2 (gpurun t08): the plain kernel runs at 0.183 ms = 5850 GB/s (89 % of the copy peak)
    // in this shape against 0.209 ms with two sites per thread; the variant with the fused CG direction update
    // is slower in it (0.330 vs 0.312 ms: twice the shuffles per site on top of a fourth input stream), so the
    // default is one site per thread for the plain kernel only.  GLB_NORMAL_SPT1: 0 = never, 34/44/... = variant
    // for both, unset = 34 for the plain kernel.
    static int spt1 = -2;
    if (spt1 == -2) {
      const char* e = getenv("GLB_NORMAL_SPT1");
      spt1 = e ? atoi(e) : -1;
    }
    if (spt1 > 0) return launch_normal_spt1(op, a, fuse, ndot, spt1);
    if (spt1 == -1 && !fuse) return launch_normal_spt1(op, a, fuse, ndot, 34);
  }
  const int stages = fuse ? stages_fused : stages_plain;
  // Measured at 4096^2 (gpurun t07): ring depth 3 vs 4, rolled vs unrolled row loop and private vs
  // line-contiguous slots all land within 0.5 % of each other (the kernel waits on DRAM, not on issue
  // slots or shared-memory wavefronts); the unrolled, line-contiguous form is kept because it issues
  // ~15 % fewer instructions and requests every sector from L2 once.
  if (stages == 4) return launch_normal_s<4, true, 1>(op, a, fuse, ndot);
  if (stages == 3) return launch_normal_s<3, true, 1>(op, a, fuse, ndot);
  return launch_normal_s<0, false, 0>(op, a, fuse, ndot);  // register prefetch (unrolling it spills)
}

}  // namespace glb
