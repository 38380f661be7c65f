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

template <int NC>
static int launch_coarse_nc(glb_context* ctx, const CoarseArgs& a, int ndot, size_t L) {
  const int grid = blas_grid(ctx, L, 256, 1);
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
  switch (op->nc) {
    case 1: return launch_coarse_nc<1>(ctx, a, ndot, L);
    case 2: return launch_coarse_nc<2>(ctx, a, ndot, L);
    case 4: return launch_coarse_nc<4>(ctx, a, ndot, L);
    case 8: return launch_coarse_nc<8>(ctx, a, ndot, L);
    default: return launch_coarse_nc<0>(ctx, a, ndot, L);
  }
}

}  // namespace glb
