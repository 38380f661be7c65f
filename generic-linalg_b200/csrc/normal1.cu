// normal1.cu -- the one-pass D^dagger D kernel of normal.cu with ONE site per thread.
//
// Why a second shape: with two sites per thread the kernel keeps 14 complex window values per site
// pair (165 registers, 12 warps per SM) and ncu shows it waiting on its own instruction latencies
// (FP64 dependency chains, shuffle / shared-memory round trips: 80 % of the stall samples), not on
// DRAM.  One site per thread halves the windows: ~100 registers, 16-20 warps per SM, and each warp
// row is half as long.  The price is a wider relative halo: a warp loads 32 consecutive sites and
// produces the 28 inner ones (12.5 % redundant L2 loads / flops instead of 6.25 %).
//
// Everything else is normal.cu: psi(y), psi(y+1), t(y-1), t(y) rolled in three-slot register rings
// (row loop unrolled by three), loads through a per-warp cp.async ring (one 16-byte LDGSTS per lane per
// array per row: a 512-byte contiguous run, read back conflict-free), x neighbours by warp shuffle, the
// reference's expression order without FMA contraction (bit-identical results), fused CG direction
// update and <p,Ap> epilogue, slabs with two ghost rows per side.
#include <cstdlib>
#include <type_traits>

#include "normal_args.cuh"

namespace glb {


constexpr int N1_THREADS = 128;
constexpr int N1_WARPS = N1_THREADS / 32;
// outputs per warp row = 32 - 2*HALO: HALO = 2 is the minimum (28 outputs); HALO = 4 makes every warp's stores whole
// 128-byte lines (24 outputs = 384 bytes, line-aligned when X is a multiple of 8) at the price of more redundant loads

// hopping term at this lane's site, reference order (operators.cpp:215-224); eta = -1 on odd x
__device__ __forceinline__ cplx hop1(bool eta_neg, cplx ux, cplx ux_m, cplx uy, cplx uy_m, cplx psi_xp, cplx psi_xm,
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
__device__ __forceinline__ cplx row1(bool eta_neg, cplx below, cplx centre, cplx above, cplx ux, cplx ux_left, cplx uy,
                                     cplx uy_below, double mass) {
  const cplx left = shfl_up_c(centre, 1);
  const cplx right = shfl_down_c(centre, 1);
  cplx h = hop1(eta_neg, ux, ux_left, uy, uy_below, right, left, above, below);
  if (DAGGER) h = fneg(h);
  return fadd(fscale(0.5, h), fscale(mass, centre));
}

template <bool FUSE_XPAY, int NDOT, int STAGES, int MINB, int HALO>
__global__ void __launch_bounds__(N1_THREADS, MINB) normal1_kernel(const NormArgs a) {
  constexpr int N1_OUT = 32 - 2 * HALO;
  extern __shared__ __align__(16) unsigned char ring_raw[];
  double beta = 0.0;
  if (a.cg != nullptr) {
    if (a.cg->done) return;
    if (FUSE_XPAY) beta = xdiv(a.cg->rsq_new, a.cg->rsq_old);  // generic_cg.cpp:344
  }
  halo_wait_block(a.wait);
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;

  const int lane = threadIdx.x & 31;
  const bool eta_neg = (lane & 1);  // windows start on even x (N1_OUT*strip - HALO, both even)
  const int X = a.X, Y = a.Y;
  const int nstrips = (X + N1_OUT - 1) / N1_OUT;
  const long long nitems = (long long)nstrips * a.nrb;
  const long long nwarps = (long long)gridDim.x * N1_WARPS;
  const long long wid = (long long)blockIdx.x * N1_WARPS + (threadIdx.x >> 5);
  const bool slab = (a.g_lo != nullptr);
  constexpr int NARR = FUSE_XPAY ? 4 : 3;
  cplx* const ring_w = reinterpret_cast<cplx*>(ring_raw) + (size_t)(threadIdx.x >> 5) * (STAGES * NARR * 32) + lane;
  auto slot = [&](int stage, int arr) -> cplx* { return ring_w + (stage * NARR + arr) * 32; };

  for (long long item = wid; item < nitems; item += nwarps) {
    const int strip = (int)(item % nstrips);
    const int rb = (int)(item / nstrips);
    const int ya = (int)((long long)Y * rb / a.nrb);
    const int yb = (int)((long long)Y * (rb + 1) / a.nrb);
    if (ya >= yb) continue;
    const int xs = strip * N1_OUT - HALO + lane;
    const int x0 = ((xs % X) + X) % X;
    const bool active = (lane >= HALO) && (lane < 32 - HALO) && (xs < X);

    auto load_psi = [&](int y) -> cplx {  // the (possibly fused) input at row y, this lane's site
      if (slab && (y < 0 || y >= Y)) return (y < 0 ? a.g_lo + (size_t)(y + 2) * X : a.g_hi + (size_t)(y - Y) * X)[x0];
      const size_t o = (size_t)(((y % Y) + Y) % Y) * X + x0;
      if (FUSE_XPAY) return fadd(a.r[o], fscale(beta, a.pold[o]));
      return a.in[o];
    };
    auto link = [&](const cplx* U, int y) -> cplx { return __ldg(U + (ptrdiff_t)y * X + x0); };

    // register windows, slot = (row - ya) % 3 (see normal.cu)
    cplx p[3], t[3], ux[3], uy[3], uxl[3];
    {
      const cplx p_mm = load_psi(ya - 2), p_m = load_psi(ya - 1);
      p[0] = load_psi(ya);
      p[1] = load_psi(ya + 1);
      const cplx uy_mm = link(a.Uy, ya - 2), ux_m = link(a.Ux, ya - 1);
      uy[0] = link(a.Uy, ya - 1);
      ux[1] = link(a.Ux, ya);
      uy[1] = link(a.Uy, ya);
      const cplx uxl_m = shfl_up_c(ux_m, 1);
      uxl[1] = shfl_up_c(ux[1], 1);
      t[0] = row1<false>(eta_neg, p_mm, p_m, p[0], ux_m, uxl_m, uy[0], uy_mm, a.mass);    // t(ya-1)
      t[1] = row1<false>(eta_neg, p_m, p[0], p[1], ux[1], uxl[1], uy[1], uy[0], a.mass);  // t(ya)
    }
    // producer: output row of the next stage, its slot, the (wrapped) input row y+2 it reads
    int is_y = ya, is_st = 0;
    int is_row2 = ya + 2;
    if (!slab && is_row2 >= Y) is_row2 -= Y;
    auto issue = [&]() {
      if (is_y < yb) {
        const ptrdiff_t o1 = (ptrdiff_t)(is_y + 1) * X + x0;
        if (slab && is_row2 >= Y) {
          cp_async16(slot(is_st, 0), a.g_hi + (size_t)(is_row2 - Y) * X + x0);
          if (FUSE_XPAY) *slot(is_st, 3) = mk(0.0, 0.0);
        } else {
          const size_t o2 = (size_t)is_row2 * X + x0;
          cp_async16(slot(is_st, 0), (FUSE_XPAY ? a.r : a.in) + o2);
          if (FUSE_XPAY) cp_async16(slot(is_st, 3), a.pold + o2);
        }
        cp_async16(slot(is_st, 1), a.Ux + o1);
        cp_async16(slot(is_st, 2), a.Uy + o1);
        is_y++;
        if (++is_row2 == Y && !slab) is_row2 = 0;
        if (++is_st == STAGES) is_st = 0;
      }
      cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < STAGES - 1; k++) issue();

    int rd_st = 0;
    auto row_step = [&](auto Kc, const int y) {
      constexpr int K0 = decltype(Kc)::value % 3, K1 = (K0 + 1) % 3, K2 = (K0 + 2) % 3;
      issue();                        // every lane reads back only what it copied itself: no warp sync needed
      cp_async_wait<STAGES - 1>();
      const cplx la = *slot(rd_st, 0);
      ux[K2] = *slot(rd_st, 1);
      uy[K2] = *slot(rd_st, 2);
      if (FUSE_XPAY) {
        const cplx lb = *slot(rd_st, 3);
        p[K2] = fadd(la, fscale(beta, lb));
      } else {
        p[K2] = la;
      }
      if (++rd_st == STAGES) rd_st = 0;
      uxl[K2] = shfl_up_c(ux[K2], 1);
      t[K2] = row1<false>(eta_neg, p[K0], p[K1], p[K2], ux[K2], uxl[K2], uy[K2], uy[K1], a.mass);  // t(y+1) = D psi
      const cplx res = row1<true>(eta_neg, t[K0], t[K1], t[K2], ux[K1], uxl[K1], uy[K1], uy[K0], a.mass);  // out(y)
      if (active) {
        const size_t o = (size_t)y * X + x0;
        a.out[o] = res;
        if (FUSE_XPAY) a.pnew[o] = p[K0];
        if (NDOT >= 1) {
          const cplx wv = (a.w == nullptr) ? p[K0] : a.w[o];
          Field<cplx>::dot_acc(acc, wv, res);
        }
        if (NDOT >= 2) acc[2] += fnorm(res);
      }
    };
    int y = ya;
#pragma unroll 1
    for (; y + 3 <= yb; y += 3) {
      row_step(std::integral_constant<int, 0>(), y);
      row_step(std::integral_constant<int, 1>(), y + 1);
      row_step(std::integral_constant<int, 2>(), y + 2);
    }
    if (y < yb) row_step(std::integral_constant<int, 0>(), y);
    if (y + 1 < yb) row_step(std::integral_constant<int, 1>(), y + 1);
    cp_async_wait<0>();
  }

  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total)) {
      if (a.cg != nullptr && a.cg_role == 3) p2p_allreduce_block(a.pr, total, 2);
      if (threadIdx.x == 0 && a.cg != nullptr) {
        if (a.cg_role == 1 || a.cg_role == 3) {
          a.cg->pAp_re = total[0];
          a.cg->pAp_im = total[1];
          a.cg->rsq_old = a.cg->rsq_new;
        } else if (a.cg_role == 2) {
          a.cg->partial[1] = total[0];
          a.cg->partial[2] = total[1];
        }
      }
    }
  }
}

template <bool FUSE, int NDOT, int STAGES, int MINB, int HALO>
static int launch_n1_t(glb_operator* op, const NormArgs& a) {
  glb_context* ctx = op->ctx;
  constexpr int N1_OUT = 32 - 2 * HALO;
  auto kern = normal1_kernel<FUSE, NDOT, STAGES, MINB, HALO>;
  const size_t smem = (size_t)N1_WARPS * STAGES * (FUSE ? 4 : 3) * 32 * sizeof(cplx);
  static int per_sm = 0;
  if (per_sm == 0) {
    if (smem + 2048 > 48 * 1024) GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, N1_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const long long nstrips = (a.X + N1_OUT - 1) / N1_OUT;
  const long long max_warps = (long long)ctx->sm_count * per_sm * N1_WARPS;
  long long nrb = max_warps / nstrips;
  const long long nrb_cap = a.Y >= 16 ? a.Y / 8 : 1;
  if (nrb > nrb_cap) nrb = nrb_cap;
  if (nrb < 1) nrb = 1;
  NormArgs b = a;
  b.nrb = (int)nrb;
  long long blocks = (nstrips * nrb + N1_WARPS - 1) / N1_WARPS;
  if (blocks > (long long)ctx->sm_count * per_sm) blocks = (long long)ctx->sm_count * per_sm;
  if (blocks < 1) blocks = 1;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  ProfScope prof(ctx, FUSE ? PROF_NORMAL_FUSED : PROF_NORMAL);
  kern<<<(unsigned)blocks, N1_THREADS, smem, ctx->stream>>>(b);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

template <int STAGES, int MINB, int HALO>
static int launch_n1_s(glb_operator* op, const NormArgs& a, bool fuse, int ndot) {
  if (fuse) {
    if (ndot == 0) return launch_n1_t<true, 0, STAGES, MINB, HALO>(op, a);
    if (ndot == 1) return launch_n1_t<true, 1, STAGES, MINB, HALO>(op, a);
    return launch_n1_t<true, 2, STAGES, MINB, HALO>(op, a);
  }
  if (ndot == 0) return launch_n1_t<false, 0, STAGES, MINB, HALO>(op, a);
  if (ndot == 1) return launch_n1_t<false, 1, STAGES, MINB, HALO>(op, a);
  return launch_n1_t<false, 2, STAGES, MINB, HALO>(op, a);
}

// variant (GLB_NORMAL_SPT1): 10*STAGES + min blocks per SM.  34 = 3 stages, 4 blocks per SM (<= 128 registers).
// Measured and not instantiated (gpurun t08, t09): five or six blocks per SM force spills (0.189 / 0.216 ms against
// 0.184 ms for the plain kernel at 4096^2); HALO = 4 (24 outputs per warp row, every store a whole 128-byte line)
// costs more in redundant loads than the aligned stores give back (0.219 ms).
int launch_normal_spt1(glb_operator* op, const NormArgs& a, bool fuse, int ndot, int variant) {
  if (variant == 44) return launch_n1_s<4, 4, 2>(op, a, fuse, ndot);
  return launch_n1_s<3, 4, 2>(op, a, fuse, ndot);
}

}  // namespace glb
