// coarse.cu -- data-driven nc x nc stencil: stencil_2d (stencil_2d/coarse_stencil.h:33) applied as
// apply_stencil_2d does on its DIR_ALL path (stencil_2d/coarse_stencil.cpp:29-172).
//
// dof i = site*nc + row, site = y*X + x.  Matrices keep the reference layout on the device:
//   clover[c + nc*i], hopping[c + nc*i + dir*nc*L] (dir = +x,+y,-x,-y), two_link[... + dir*nc*L]
//   (dir = +2x, +x+y, +2y, -x+y, -2x, -x-y, -2y, +x-y), L = V*nc  -- i.e. one contiguous row of nc
//   complex numbers per (dof, direction), rows of a site adjacent: a warp streams whole 128-byte
//   lines of matrix data, each element touched once (5*nc^2+2*nc complex per site, HBM-bound).
// One thread owns one output dof and accumulates clover, +x, +y, -x, -y, (two-link), shifts in
// exactly the reference's order without FMA contraction -> bit-identical results.
#include <cstdint>
#include <cstdlib>

#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

struct CoarseArgs {
  const cplx* in;
  const cplx* in_lo;  // row(s) below the slab: 1 row (2 if has_two), lowest y first
  const cplx* in_hi;  // row(s) above the slab
  cplx* out;
  const cplx* w;
  const cplx* clover;
  const cplx* hopping;
  const cplx* two_link;
  int X, Yloc, y0, nc, has_two;
  cplx shift, eo_shift, dof_shift;
  int use_shift, use_eo, use_dof;
  ReduceWs red;
  CgState* cg;
  int cg_role;
};

// pointer to the nc entries of site (x, y) where y may run from -2 to Yloc+1
__device__ __forceinline__ const cplx* site_ptr(const CoarseArgs& a, int x, int y) {
  const size_t rowlen = (size_t)a.X * a.nc;
  const int depth = a.has_two ? 2 : 1;
  if (y < 0) return a.in_lo + (size_t)(depth + y) * rowlen + (size_t)x * a.nc;
  if (y >= a.Yloc) return a.in_hi + (size_t)(y - a.Yloc) * rowlen + (size_t)x * a.nc;
  return a.in + (size_t)y * rowlen + (size_t)x * a.nc;
}

template <int NC>
__device__ __forceinline__ cplx row_times(cplx acc, const cplx* __restrict__ M, const cplx* __restrict__ v, int nc) {
  if (NC > 0) {
    cplx mm[NC > 0 ? NC : 1], vv[NC > 0 ? NC : 1];
    if (NC % 2 == 0) {
      // rows of an even nc start 32-byte aligned: LDG.256 halves the L1 wavefronts of this
      // row-per-thread pattern (lanes 16*nc bytes apart), which is what bounded the kernel at nc = 8
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        cplx pr[2];
        ldv_nc<2>(M + c, pr);
        mm[c] = pr[0];
        mm[c + (NC > 1 ? 1 : 0)] = pr[1];
      }
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        cplx pr[2];
        ldv<2>(v + c, pr);
        vv[c] = pr[0];
        vv[c + (NC > 1 ? 1 : 0)] = pr[1];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; c++) mm[c] = __ldg(M + c);
#pragma unroll
      for (int c = 0; c < NC; c++) vv[c] = v[c];
    }
#pragma unroll
    for (int c = 0; c < NC; c++) acc = fadd(acc, fmul(mm[c], vv[c]));
  } else {
    for (int c = 0; c < nc; c++) acc = fadd(acc, fmul(__ldg(M + c), v[c]));
  }
  return acc;
}

template <int NC, int NDOT>
__global__ void __launch_bounds__(256) coarse_kernel(const CoarseArgs a) {
  if (a.cg != nullptr && a.cg->done) return;
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;
  const int nc = (NC > 0) ? NC : a.nc;
  const int X = a.X;
  const size_t L = (size_t)X * a.Yloc * nc;  // local dofs
  const size_t plane = L * nc;               // one direction's matrices
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i % nc);
    const size_t site = i / nc;
    const int x = (int)(site % X), y = (int)(site / X);
    const int xp = (x + 1 == X) ? 0 : x + 1, xm = (x == 0) ? X - 1 : x - 1;
    cplx s = mk(0.0, 0.0);
    const cplx* Mrow = a.clover + i * nc;
    s = row_times<NC>(s, Mrow, site_ptr(a, x, y), nc);
    const cplx* H = a.hopping + i * nc;
    s = row_times<NC>(s, H, site_ptr(a, xp, y), nc);
    s = row_times<NC>(s, H + plane, site_ptr(a, x, y + 1), nc);
    s = row_times<NC>(s, H + 2 * plane, site_ptr(a, xm, y), nc);
    s = row_times<NC>(s, H + 3 * plane, site_ptr(a, x, y - 1), nc);
    if (a.has_two) {
      const int xpp = (x + 2) % X, xmm = (x - 2 + 2 * X) % X;
      const cplx* T = a.two_link + i * nc;
      s = row_times<NC>(s, T, site_ptr(a, xpp, y), nc);
      s = row_times<NC>(s, T + plane, site_ptr(a, xp, y + 1), nc);
      s = row_times<NC>(s, T + 2 * plane, site_ptr(a, x, y + 2), nc);
      s = row_times<NC>(s, T + 3 * plane, site_ptr(a, xm, y + 1), nc);
      s = row_times<NC>(s, T + 4 * plane, site_ptr(a, xmm, y), nc);
      s = row_times<NC>(s, T + 5 * plane, site_ptr(a, xm, y - 1), nc);
      s = row_times<NC>(s, T + 6 * plane, site_ptr(a, x, y - 2), nc);
      s = row_times<NC>(s, T + 7 * plane, site_ptr(a, xp, y - 1), nc);
    }
    const cplx self = a.in[i];
    if (a.use_shift) s = fadd(s, fmul(a.shift, self));  // coarse_stencil.cpp:153-156
    if (a.use_eo) {                                     // coarse_stencil.cpp:159-162
      const bool odd = ((x + y + a.y0) & 1);
      s = fadd(s, fmul(odd ? fneg(a.eo_shift) : a.eo_shift, self));
    }
    if (a.use_dof) {                                    // coarse_stencil.cpp:165-169
      s = fadd(s, fmul(row < nc / 2 ? a.dof_shift : fneg(a.dof_shift), self));
    }
    a.out[i] = s;
    if (NDOT >= 1) {
      const cplx wv = (a.w == nullptr) ? self : a.w[i];
      Field<cplx>::dot_acc(acc, wv, s);
    }
    if (NDOT >= 2) acc[2] += fnorm(s);
  }
  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total) && threadIdx.x == 0) {
      if (a.cg != nullptr && a.cg_role == 1) {
        a.cg->pAp_re = total[0];
        a.cg->pAp_im = total[1];
        a.cg->rsq_old = a.cg->rsq_new;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// nc = 1 (the fine-level stencil of get_square_staggered_u1_stencil, operators_stencil.cpp:14-63),
// five-point, even X: one thread per PAIR of adjacent sites.  The five matrix planes, the input pair
// and the rows above / below arrive as 32-byte LDG.256 (10 loads for two sites instead of 22 16-byte
// ones); the two outer x neighbours are 16-byte loads that hit L1.  112 B/site.
template <int NDOT>
__global__ void __launch_bounds__(256) stencil1_pair_kernel(const CoarseArgs a) {
  if (a.cg != nullptr && a.cg->done) return;
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;
  const int X = a.X, HX = X / 2;
  const size_t L = (size_t)X * a.Yloc;
  const size_t npairs = L / 2;
  for (size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pi < npairs; pi += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(pi / HX);
    const int x = 2 * (int)(pi % HX);
    const size_t i = (size_t)y * X + x;
    cplx cl[2], hxp[2], hyp[2], hxm[2], hym[2], c[2], up[2], dn[2];
    ldv_nc<2>(a.clover + i, cl);
    ldv_nc<2>(a.hopping + i, hxp);
    ldv_nc<2>(a.hopping + L + i, hyp);
    ldv_nc<2>(a.hopping + 2 * L + i, hxm);
    ldv_nc<2>(a.hopping + 3 * L + i, hym);
    ldv<2>(a.in + i, c);
    ldv<2>(site_ptr(a, x, y + 1), up);
    ldv<2>(site_ptr(a, x, y - 1), dn);
    const cplx left = a.in[(size_t)y * X + (x == 0 ? X - 1 : x - 1)];
    const cplx right = a.in[(size_t)y * X + (x + 2 == X ? 0 : x + 2)];
    cplx s[2];
    s[0] = fadd(mk(0.0, 0.0), fmul(cl[0], c[0]));
    s[1] = fadd(mk(0.0, 0.0), fmul(cl[1], c[1]));
    s[0] = fadd(s[0], fmul(hxp[0], c[1]));
    s[1] = fadd(s[1], fmul(hxp[1], right));
    s[0] = fadd(s[0], fmul(hyp[0], up[0]));
    s[1] = fadd(s[1], fmul(hyp[1], up[1]));
    s[0] = fadd(s[0], fmul(hxm[0], left));
    s[1] = fadd(s[1], fmul(hxm[1], c[0]));
    s[0] = fadd(s[0], fmul(hym[0], dn[0]));
    s[1] = fadd(s[1], fmul(hym[1], dn[1]));
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (a.use_shift) s[k] = fadd(s[k], fmul(a.shift, c[k]));  // coarse_stencil.cpp:153-156
      if (a.use_eo) {                                           // coarse_stencil.cpp:159-162
        const bool odd = ((x + k + y + a.y0) & 1);
        s[k] = fadd(s[k], fmul(odd ? fneg(a.eo_shift) : a.eo_shift, c[k]));
      }
      // dof_shift: row < nc/2 is never true at nc = 1 (coarse_stencil.cpp:165-169)
      if (a.use_dof) s[k] = fadd(s[k], fmul(fneg(a.dof_shift), c[k]));
    }
    stv<2>(a.out + i, s);
    if (NDOT >= 1) {
      cplx wv[2];
      if (a.w == nullptr) {
        wv[0] = c[0];
        wv[1] = c[1];
      } else {
        ldv<2>(a.w + i, wv);
      }
      Field<cplx>::dot_acc(acc, wv[0], s[0]);
      Field<cplx>::dot_acc(acc, wv[1], s[1]);
    }
    if (NDOT >= 2) {
      acc[2] += fnorm(s[0]);
      acc[2] += fnorm(s[1]);
    }
  }
  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total) && threadIdx.x == 0) {
      if (a.cg != nullptr && a.cg_role == 1) {
        a.cg->pAp_re = total[0];
        a.cg->pAp_im = total[1];
        a.cg->rsq_old = a.cg->rsq_new;
      }
    }
  }
}

// algorithmic bytes of one stencil apply over L dofs: (5 or 13) nc x nc matrices + in + out per site (SURVEY 8 d-bytes)
static double coarse_bytes(const CoarseArgs& a, size_t L) { return (double)L * ((a.has_two ? 13.0 : 5.0) * a.nc + 2.0) * 16.0; }

static int launch_stencil1_pair(glb_context* ctx, const CoarseArgs& a, int ndot, size_t L) {
  const int grid = blas_grid(ctx, L / 2, 256, 1);
  ProfScope prof(ctx, PROF_COARSE, coarse_bytes(a, L));
  if (ndot == 0)
    stencil1_pair_kernel<0><<<grid, 256, 0, ctx->stream>>>(a);
  else if (ndot == 1)
    stencil1_pair_kernel<1><<<grid, 256, 0, ctx->stream>>>(a);
  else
    stencil1_pair_kernel<2><<<grid, 256, 0, ctx->stream>>>(a);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

// ------------------------------------------------------------------------------------------------
// nc = 8 / 16: matrices streamed through a per-warp cp.async ring.
//
// With one output row per thread the direct kernel above makes every lane read its own 128-byte
// line (nc = 8): each LDG touches 32 lines and the L1 tag stage, not HBM, bounds it (63 % of peak,
// profiles/r01_configs.jsonl).  Here a warp owns a tile of 32 consecutive rows (= 32/nc sites); one
// pipeline stage is ONE direction's matrices of that tile, a contiguous 32*nc*16-byte run of the
// plane, copied by LDGSTS with consecutive lanes on consecutive 16-byte chunks (4 lines per request)
// together with the 512 bytes of neighbour-site input the stage multiplies.  Chunks are XOR-swizzled
// by row so that the row-per-thread LDS.128 reads are bank-conflict free.  RING_STAGES-1 stages per
// warp stay in flight; the warps never meet at a block barrier.  The accumulation runs clover, +x,
// +y, -x, -y, (two-link), shifts over c = 0..nc-1 inside one thread exactly as the direct kernel:
// bit-identical results.
constexpr int RING_THREADS = 128;
constexpr int RING_WARPS = RING_THREADS / 32;

__device__ __forceinline__ cplx lds_c(const cplx* p) {
  cplx r;
  const unsigned s = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(s));
  return r;
}

template <int NC, int NDOT, bool HAS_TWO, int STAGES>
__global__ void __launch_bounds__(RING_THREADS) coarse_ring_kernel(const CoarseArgs a, const long long ntiles) {
  extern __shared__ __align__(128) unsigned char ring_raw[];
  if (a.cg != nullptr && a.cg->done) return;
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  constexpr int NDIR = HAS_TWO ? 13 : 5;
  constexpr int SITES = 32 / NC;                    // sites per tile
  constexpr int STAGE_ELEMS = 32 * NC + 32;         // matrices + 32 chunks of input
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int X = a.X;
  const size_t L = (size_t)X * a.Yloc * NC;
  const size_t plane = L * NC;
  cplx* ring = reinterpret_cast<cplx*>(ring_raw) + (size_t)warp * STAGES * STAGE_ELEMS;
  const long long nwarps = (long long)gridDim.x * RING_WARPS;
  const long long first = (long long)blockIdx.x * RING_WARPS + warp;
  const long long mine = (first < ntiles) ? (ntiles - first + nwarps - 1) / nwarps : 0;

  // neighbour of direction d (order of apply_stencil_2d: clover, +x, +y, -x, -y, then the eight
  // two-link corners +2x, +x+y, +2y, -x+y, -2x, -x-y, -2y, +x-y)
  auto nbr = [&](int d, int x, int y, int& xn, int& yn) {
    int dx = 0, dy = 0;
    switch (d) {
      case 1: dx = 1; break;
      case 2: dy = 1; break;
      case 3: dx = -1; break;
      case 4: dy = -1; break;
      case 5: dx = 2; break;
      case 6: dx = 1; dy = 1; break;
      case 7: dy = 2; break;
      case 8: dx = -1; dy = 1; break;
      case 9: dx = -2; break;
      case 10: dx = -1; dy = -1; break;
      case 11: dy = -2; break;
      case 12: dx = 1; dy = -1; break;
      default: break;
    }
    xn = x + dx;
    if (xn >= X) xn -= X;
    if (xn < 0) xn += X;
    yn = y + dy;
  };

  // producer side: stage (tile, d) into ring slot `st`
  long long is_tile = first;  // tile / direction of the next stage to issue
  int is_d = 0;
  long long is_left = mine * NDIR;
  int is_slot = 0;
  auto issue = [&]() {
    if (is_left > 0) {
      cplx* dst = ring + (size_t)is_slot * STAGE_ELEMS;
      const cplx* M = (is_d == 0) ? a.clover : (is_d < 5 ? a.hopping + (size_t)(is_d - 1) * plane
                                                         : a.two_link + (size_t)(is_d - 5) * plane);
      M += (size_t)is_tile * 32 * NC;
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const int e = lane + 32 * k;
        const int row = e / NC, col = e % NC;
        cp_async16(dst + row * NC + (col ^ (row & 7)), M + e);
      }
      // input of the neighbour sites: lane -> (site of the tile, chunk)
      const size_t site = (size_t)is_tile * SITES + lane / NC;
      const int x = (int)(site % X), y = (int)(site / X);
      int xn, yn;
      nbr(is_d, x, y, xn, yn);
      cp_async16(dst + 32 * NC + lane, site_ptr(a, xn, yn) + (lane % NC));
      is_left--;
      if (++is_d == NDIR) {
        is_d = 0;
        is_tile += nwarps;
      }
      if (++is_slot == STAGES) is_slot = 0;
    }
    cp_async_commit();
  };

#pragma unroll
  for (int k = 0; k < STAGES - 1; k++) issue();

  int slot = 0;
  long long tile = first;
#pragma unroll 1
  for (long long t = 0; t < mine; t++, tile += nwarps) {
    cplx s = mk(0.0, 0.0);
    cplx self = mk(0.0, 0.0);
#pragma unroll 1
    for (int d = 0; d < NDIR; d++) {
      __syncwarp();  // every lane is done with the slot the next copy overwrites
      issue();
      cp_async_wait<STAGES - 1>();
      __syncwarp();  // ... and sees the chunks its neighbours copied
      const cplx* st = ring + (size_t)slot * STAGE_ELEMS;
      const cplx* mrow = st + lane * NC;
      const cplx* vrow = st + 32 * NC + (lane / NC) * NC;
      if (d == 0) self = lds_c(vrow + (lane % NC));
#pragma unroll
      for (int c = 0; c < NC; c++) s = fadd(s, fmul(lds_c(mrow + (c ^ (lane & 7))), lds_c(vrow + c)));
      if (++slot == STAGES) slot = 0;
    }
    const size_t i = (size_t)tile * 32 + lane;
    const int row = lane % NC;
    if (a.use_shift) s = fadd(s, fmul(a.shift, self));  // coarse_stencil.cpp:153-156
    if (a.use_eo) {                                     // coarse_stencil.cpp:159-162
      const size_t site = i / NC;
      const int x = (int)(site % X), y = (int)(site / X);
      const bool odd = ((x + y + a.y0) & 1);
      s = fadd(s, fmul(odd ? fneg(a.eo_shift) : a.eo_shift, self));
    }
    if (a.use_dof) s = fadd(s, fmul(row < NC / 2 ? a.dof_shift : fneg(a.dof_shift), self));  // :165-169
    a.out[i] = s;
    if (NDOT >= 1) {
      const cplx wv = (a.w == nullptr) ? self : a.w[i];
      Field<cplx>::dot_acc(acc, wv, s);
    }
    if (NDOT >= 2) acc[2] += fnorm(s);
  }
  cp_async_wait<0>();
  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total) && threadIdx.x == 0) {
      if (a.cg != nullptr && a.cg_role == 1) {
        a.cg->pAp_re = total[0];
        a.cg->pAp_im = total[1];
        a.cg->rsq_old = a.cg->rsq_new;
      }
    }
  }
}

template <int NC, int NDOT, bool HAS_TWO, int STAGES>
static int launch_ring_t(glb_context* ctx, const CoarseArgs& a, size_t L) {
  auto kern = coarse_ring_kernel<NC, NDOT, HAS_TWO, STAGES>;
  const size_t smem = (size_t)RING_WARPS * STAGES * (32 * NC + 32) * sizeof(cplx);
  static int per_sm = 0;
  if (per_sm == 0) {
    GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RING_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const long long ntiles = (long long)(L / 32);
  long long blocks = (ntiles + RING_WARPS - 1) / RING_WARPS;
  if (blocks > (long long)ctx->sm_count * per_sm) blocks = (long long)ctx->sm_count * per_sm;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  if (blocks < 1) blocks = 1;
  ProfScope prof(ctx, PROF_COARSE, coarse_bytes(a, L));
  kern<<<(unsigned)blocks, RING_THREADS, smem, ctx->stream>>>(a, ntiles);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

template <int NC, int STAGES>
static int launch_ring(glb_context* ctx, const CoarseArgs& a, int ndot, size_t L) {
  if (a.has_two) {
    if (ndot == 0) return launch_ring_t<NC, 0, true, STAGES>(ctx, a, L);
    if (ndot == 1) return launch_ring_t<NC, 1, true, STAGES>(ctx, a, L);
    return launch_ring_t<NC, 2, true, STAGES>(ctx, a, L);
  }
  if (ndot == 0) return launch_ring_t<NC, 0, false, STAGES>(ctx, a, L);
  if (ndot == 1) return launch_ring_t<NC, 1, false, STAGES>(ctx, a, L);
  return launch_ring_t<NC, 2, false, STAGES>(ctx, a, L);
}

static bool ring_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GLB_COARSE_RING");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return enabled != 0;
}

// ------------------------------------------------------------------------------------------------
// The partial applies of the preconditioned paths (coarse_stencil.cpp:395-1512, DIR_ALL):
//   EO / OE : hopping term only, output on even / odd sites, the other parity zeroed        (:395, :560)
//   TB / BT : clover + hopping (+ two-link) restricted to the top<-bottom / bottom<-top halves of the
//             colour index: rows < nc/2 summed over c >= nc/2, or the mirror; other rows zeroed (:725, :1120)
// No shifts.  One thread per output dof, reference accumulation order (bit-identical); these serve the e/o and
// t/b preconditioned solves, not the headline path.
//
// post selects what is stored (the even/odd and top/bottom preconditioned stencil paths, operators_stencil.cpp:179-236,
// mg_complex.cpp:1211-1330, fused into the pass that produces the hopping term s):
//   0: s | 0                              apply_stencil_2d_{eo,oe,tb,bt}
//   1: coef*aux - s | 0                   *prec_prepare (coef = shift), second pass of m^2 - D_ab D_ba (coef = shift^2)
//   2: coef.re*(aux - s) | in             *prec_reconstruct (coef.re = 1/Re shift; the other half is copied from in)
//   3: coef*aux - s on EVERY dof          first half of m^2 - D_ab D_ba - D_ba D_ab (s = 0 off the live half)
//   4: out - s on EVERY dof               second half of the same (out is read and rewritten)
//
// Thread mapping.  Half of the dofs are dead in every partial apply.  PAIRED = true gives each thread one LIVE dof and
// its dead partner (e/o: the x-neighbour inside the same pair of sites, needs an even X; t/b: the same row in the
// other half of the colour index, needs an even nc), so that no lane idles while the matrices stream in; the sum for
// the live dof runs in the reference's order either way.  PAIRED = false is the one-thread-per-dof form for odd X / nc.
// WIDE: the colour range is even and every array 32-byte aligned -> matrix rows and input sites are read two
// complex numbers at a time (LDG.256); the additions keep their order.
template <bool WIDE>
__device__ __forceinline__ cplx part_sum(const CoarseArgs& a, const int part, const size_t i, const int x, const int y,
                                         const size_t plane) {
  const int nc = a.nc, X = a.X;
  const size_t site = i / nc;
  cplx s = mk(0.0, 0.0);
  int c0 = 0, c1 = nc;
  const bool colour_split = (part == GLB_PART_TB || part == GLB_PART_BT);
  if (colour_split) {
    c0 = (part == GLB_PART_TB) ? nc / 2 : 0;
    c1 = (part == GLB_PART_TB) ? nc : nc / 2;
  }
  (void)site;
  const int xp = (x + 1 == X) ? 0 : x + 1, xm = (x == 0) ? X - 1 : x - 1;
  auto term = [&](const cplx* M, const cplx* v) {
    if (WIDE) {
      for (int c = c0; c < c1; c += 2) {
        cplx m2[2], v2[2];
        ldv_nc<2>(M + c, m2);
        ldv_nc<2>(v + c, v2);
        s = fadd(s, fmul(m2[0], v2[0]));
        s = fadd(s, fmul(m2[1], v2[1]));
      }
    } else {
      for (int c = c0; c < c1; c++) s = fadd(s, fmul(__ldg(M + c), v[c]));
    }
  };
  if (colour_split) term(a.clover + i * nc, site_ptr(a, x, y));
  const cplx* H = a.hopping + i * nc;
  term(H, site_ptr(a, xp, y));
  term(H + plane, site_ptr(a, x, y + 1));
  term(H + 2 * plane, site_ptr(a, xm, y));
  term(H + 3 * plane, site_ptr(a, x, y - 1));
  if (a.has_two && colour_split) {
    const int xpp = (x + 2) % X, xmm = (x - 2 + 2 * X) % X;
    const cplx* T = a.two_link + i * nc;
    term(T, site_ptr(a, xpp, y));
    term(T + plane, site_ptr(a, xp, y + 1));
    term(T + 2 * plane, site_ptr(a, x, y + 2));
    term(T + 3 * plane, site_ptr(a, xm, y + 1));
    term(T + 4 * plane, site_ptr(a, xmm, y));
    term(T + 5 * plane, site_ptr(a, xm, y - 1));
    term(T + 6 * plane, site_ptr(a, x, y - 2));
    term(T + 7 * plane, site_ptr(a, xp, y - 1));
  }
  return s;
}

__device__ __forceinline__ void part_store(const CoarseArgs& a, const int post, const bool live, const cplx s, const cplx coef,
                                           const cplx* __restrict__ aux, const size_t i) {
  cplx r;
  switch (post) {
    case 1: r = live ? fsub(fmul(coef, aux[i]), s) : mk(0.0, 0.0); break;
    case 2: r = live ? fscale(coef.x, fsub(aux[i], s)) : a.in[i]; break;
    case 3: r = fsub(fmul(coef, aux[i]), s); break;
    case 4:
      if (!live) return;  // out - 0 = out
      r = fsub(a.out[i], s);
      break;
    default: r = s; break;
  }
  a.out[i] = r;
}

template <bool PAIRED, bool WIDE>
__global__ void __launch_bounds__(256) coarse_part_kernel(const CoarseArgs a, const int part, const int post, const cplx coef,
                                                          const cplx* __restrict__ aux) {
  const int nc = a.nc;
  const int X = a.X;
  const size_t L = (size_t)X * a.Yloc * nc;
  const size_t plane = L * nc;
  const bool site_split = (part == GLB_PART_EO || part == GLB_PART_OE);
  if (PAIRED) {
    const int h = nc / 2, Xh = X / 2;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L / 2; j += (size_t)gridDim.x * blockDim.x) {
      size_t i_live, i_dead;
      int x, y;
      if (site_split) {
        const int row = (int)(j % nc);
        const size_t ps = j / nc;
        y = (int)(ps / Xh);
        const int k = (int)(ps % Xh);
        const int b = (y + a.y0) & 1;                               // parity of the row offset
        x = 2 * k + ((part == GLB_PART_EO) ? b : 1 - b);            // the site of the pair with the wanted parity
        i_live = ((size_t)y * X + x) * nc + row;
        i_dead = ((size_t)y * X + (x ^ 1)) * nc + row;
      } else {
        const int r = (int)(j % h);
        const size_t site = j / h;
        x = (int)(site % X);
        y = (int)(site / X);
        i_live = site * nc + ((part == GLB_PART_TB) ? r : r + h);
        i_dead = site * nc + ((part == GLB_PART_TB) ? r + h : r);
      }
      const cplx s = part_sum<WIDE>(a, part, i_live, x, y, plane);
      part_store(a, post, true, s, coef, aux, i_live);
      part_store(a, post, false, mk(0.0, 0.0), coef, aux, i_dead);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (size_t)gridDim.x * blockDim.x) {
      const int row = (int)(i % nc);
      const size_t site = i / nc;
      const int x = (int)(site % X), y = (int)(site / X);
      bool live;
      if (site_split) {
        const bool even = (((x + y + a.y0) & 1) == 0);
        live = (part == GLB_PART_EO) ? even : !even;
      } else {
        const bool top = row < nc / 2;
        live = (part == GLB_PART_TB) ? top : !top;
      }
      const cplx s = live ? part_sum<false>(a, part, i, x, y, plane) : mk(0.0, 0.0);
      part_store(a, post, live, s, coef, aux, i);
    }
  }
}

// lattice_epsilon / lattice_sigma3 (lattice/lattice_functions.h:11-53): out = +-in by site parity (mode 0) or by
// colour half (mode 1; a copy when nc is odd)
__global__ void __launch_bounds__(256) coarse_sign_kernel(cplx* __restrict__ out, const cplx* __restrict__ in, size_t n, int X,
                                                          int nc, int y0, int mode) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    bool flip;
    if (mode == 0) {
      const size_t site = i / nc;
      flip = (((site % X) + (site / X) + y0) & 1) != 0;
    } else {
      flip = (nc % 2 == 0) && ((int)(i % nc) >= nc / 2);
    }
    const cplx v = in[i];
    out[i] = flip ? fneg(v) : v;
  }
}

int launch_stencil2d_sign(glb_operator* op, void* out, const void* in, int mode) {
  glb_context* ctx = op->ctx;
  const size_t n = (size_t)op->X * op->Yloc * op->nc;
  const int grid = blas_grid(ctx, n, 256, 1);
  coarse_sign_kernel<<<grid, 256, 0, ctx->stream>>>((cplx*)out, (const cplx*)in, n, op->X, op->nc, op->y0, mode);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_stencil2d_part(glb_operator* op, void* out, const void* in, int part, int post, const double coef[2],
                          const void* aux) {
  glb_context* ctx = op->ctx;
  if (part < GLB_PART_EO || part > GLB_PART_BT) return fail(GLB_ERR_ARG, "stencil2d: unknown partial apply");
  if (post < 0 || post > 4 || ((post >= 1 && post <= 3) && !aux)) return fail(GLB_ERR_ARG, "stencil2d: bad post-operation");
  CoarseArgs a{};
  const size_t rowlen = (size_t)op->X * op->nc;
  const bool single = (ctx->nranks == 1);
  const int depth = op->has_two ? 2 : 1;
  if (op->Yloc < depth) return fail(GLB_ERR_ARG, "stencil2d: slab thinner than the stencil reach");
  a.in = (const cplx*)in;
  a.in_lo = single ? (const cplx*)in + (size_t)(op->Yloc - depth) * rowlen : (const cplx*)op->ghost_lo;
  a.in_hi = single ? (const cplx*)in : (const cplx*)op->ghost_hi;
  a.out = (cplx*)out;
  a.clover = op->clover;
  a.hopping = op->hopping;
  a.two_link = op->two_link;
  a.X = op->X;
  a.Yloc = op->Yloc;
  a.y0 = op->y0;
  a.nc = op->nc;
  a.has_two = op->has_two ? 1 : 0;
  const size_t L = rowlen * op->Yloc;
  const bool site_split = (part == GLB_PART_EO || part == GLB_PART_OE);
  const bool paired = site_split ? (op->X % 2 == 0) : (op->nc % 2 == 0);
  const cplx cf = coef ? make_double2(coef[0], coef[1]) : make_double2(0.0, 0.0);
  ProfScope prof(ctx, PROF_COARSE);
  // two complex numbers per load: the summed colour range [c0, c1) must start and end on an even colour
  // (e/o: nc even; t/b: nc a multiple of 4) and every array must be 32-byte aligned
  const bool even_range = site_split ? (op->nc % 2 == 0) : (op->nc % 4 == 0);
  const uintptr_t addrs = (uintptr_t)a.in | (uintptr_t)a.in_lo | (uintptr_t)a.in_hi | (uintptr_t)a.clover |
                          (uintptr_t)a.hopping | (uintptr_t)a.two_link;
  const bool wide = paired && even_range && (addrs & 31u) == 0 && ring_enabled();
  if (paired && wide)
    coarse_part_kernel<true, true><<<blas_grid(ctx, L / 2, 256, 1), 256, 0, ctx->stream>>>(a, part, post, cf, (const cplx*)aux);
  else if (paired)
    coarse_part_kernel<true, false><<<blas_grid(ctx, L / 2, 256, 1), 256, 0, ctx->stream>>>(a, part, post, cf, (const cplx*)aux);
  else
    coarse_part_kernel<false, false><<<blas_grid(ctx, L, 256, 1), 256, 0, ctx->stream>>>(a, part, post, cf, (const cplx*)aux);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

template <int NC>
static int launch_coarse_nc(glb_context* ctx, const CoarseArgs& a, int ndot, size_t L) {
  const int grid = blas_grid(ctx, L, 256, 1);
  ProfScope prof(ctx, PROF_COARSE, coarse_bytes(a, L));
  if (ndot == 0)
    coarse_kernel<NC, 0><<<grid, 256, 0, ctx->stream>>>(a);
  else if (ndot == 1)
    coarse_kernel<NC, 1><<<grid, 256, 0, ctx->stream>>>(a);
  else
    coarse_kernel<NC, 2><<<grid, 256, 0, ctx->stream>>>(a);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_stencil2d(glb_operator* op, void* out, const void* in, const ApplyFusion& f) {
  glb_context* ctx = op->ctx;
  CoarseArgs a{};
  const size_t rowlen = (size_t)op->X * op->nc;
  const bool single = (ctx->nranks == 1);
  const int depth = op->has_two ? 2 : 1;
  if (op->Yloc < depth) return fail(GLB_ERR_ARG, "stencil2d: slab thinner than the stencil reach");
  a.in = (const cplx*)in;
  a.in_lo = single ? (const cplx*)in + (size_t)(op->Yloc - depth) * rowlen : (const cplx*)op->ghost_lo;
  a.in_hi = single ? (const cplx*)in : (const cplx*)op->ghost_hi;
  a.out = (cplx*)out;
  a.w = f.w_is_input ? nullptr : (const cplx*)f.w;
  a.clover = op->clover;
  a.hopping = op->hopping;
  a.two_link = op->two_link;
  a.X = op->X;
  a.Yloc = op->Yloc;
  a.y0 = op->y0;
  a.nc = op->nc;
  a.has_two = op->has_two ? 1 : 0;
  a.shift = make_double2(op->shift[0], op->shift[1]);
  a.eo_shift = make_double2(op->eo_shift[0], op->eo_shift[1]);
  a.dof_shift = make_double2(op->dof_shift[0], op->dof_shift[1]);
  // the reference tests abs(shift) != 0.0 (coarse_stencil.cpp:153,159,165)
  a.use_shift = (op->shift[0] != 0.0 || op->shift[1] != 0.0);
  a.use_eo = (op->eo_shift[0] != 0.0 || op->eo_shift[1] != 0.0);
  a.use_dof = (op->dof_shift[0] != 0.0 || op->dof_shift[1] != 0.0);
  a.red = ctx->red;
  if (!f.to_host) a.red.result_host = nullptr;
  a.cg = (CgState*)f.cg_state;
  a.cg_role = f.cg_role;
  const int ndot = (f.w != nullptr || f.w_is_input) ? (f.want_norm ? 2 : 1) : 0;
  const size_t L = rowlen * op->Yloc;
  // pairs need 32-byte aligned rows in every array (even X) and the ghost rows to be rows of X sites
  if (op->nc == 1 && !op->has_two && op->X % 2 == 0 && op->X >= 4 && ring_enabled() &&
      ((((uintptr_t)a.in | (uintptr_t)a.in_lo | (uintptr_t)a.in_hi | (uintptr_t)a.out | (uintptr_t)a.w) & 31u) == 0))
    return launch_stencil1_pair(ctx, a, ndot, L);
  if (ring_enabled() && L % 32 == 0 && L >= 32) {
    static int stages8 = -1;
    if (stages8 < 0) {
      const char* e = getenv("GLB_COARSE_STAGES");
      stages8 = e ? atoi(e) : 3;  // measured (gpurun t06): 3 stages 96.9 %, 4: 95.0 %, 6: 90.6 % of HBM peak at 512^2
    }
    if (op->nc == 8) {
      if (stages8 == 3) return launch_ring<8, 3>(ctx, a, ndot, L);
      if (stages8 == 6) return launch_ring<8, 6>(ctx, a, ndot, L);
      return launch_ring<8, 4>(ctx, a, ndot, L);
    }
    if (op->nc == 16) return launch_ring<16, 3>(ctx, a, ndot, L);
  }
  switch (op->nc) {
    case 1: return launch_coarse_nc<1>(ctx, a, ndot, L);
    case 2: return launch_coarse_nc<2>(ctx, a, ndot, L);
    case 4: return launch_coarse_nc<4>(ctx, a, ndot, L);
    case 8: return launch_coarse_nc<8>(ctx, a, ndot, L);
    default: return launch_coarse_nc<0>(ctx, a, ndot, L);
  }
}

}  // namespace glb
