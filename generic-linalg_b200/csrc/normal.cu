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

#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

constexpr int NORM_THREADS = 128;
constexpr int NORM_WARPS = NORM_THREADS / 32;
constexpr int NORM_OUT_PER_WARP = 60;  // 64 loaded - 2 halo sites on each side

struct NormArgs {
  const cplx* in;    // plain input, or nullptr when fused
  const cplx* r;     // fused direction update: input := r + beta * pold
  const cplx* pold;
  cplx* pnew;
  cplx* out;
  const cplx* Ux;
  const cplx* Uy;
  const cplx* w;     // dot partner; nullptr = the input itself
  int X, Y;
  double mass;
  ReduceWs red;
  CgState* cg;
  int cg_role;
};

// hopping part of the staggered stencil at one site, reference order (operators.cpp:215-224):
//   h = -U_x(x) psi(x+1) + conj U_x(x-1) psi(x-1) - eta U_y(x,y) psi(y+1) + eta conj U_y(x,y-1) psi(y-1)
template <bool ETA_NEG>
__device__ __forceinline__ cplx stag_hop(cplx ux, cplx ux_m, cplx uy, cplx uy_m, cplx psi_xp, cplx psi_xm, cplx psi_yp,
                                         cplx psi_ym) {
  cplx h = mk(0.0, 0.0);
  h = fsub(h, fmul(ux, psi_xp));
  h = fadd(h, fcmul(ux_m, psi_xm));
  const cplx t3 = fmul(uy, psi_yp);
  h = ETA_NEG ? fadd(h, t3) : fsub(h, t3);
  const cplx t4 = fcmul(uy_m, psi_ym);
  h = ETA_NEG ? fsub(h, t4) : fadd(h, t4);
  return h;
}

// one full row of D (DAGGER=false) or D^dag (DAGGER=true) on this lane's pair of sites.
// below/centre/above: the three input rows; the x neighbours of `centre` come from the warp.
template <bool DAGGER>
__device__ __forceinline__ void stag_row(cplx (&res)[2], const cplx (&below)[2], const cplx (&centre)[2],
                                         const cplx (&above)[2], const cplx (&ux)[2], cplx ux_left,
                                         const cplx (&uy)[2], const cplx (&uy_below)[2], double mass) {
  const cplx left = shfl_up_c(centre[1], 1);     // psi(x0-1)
  const cplx right = shfl_down_c(centre[0], 1);  // psi(x0+2)
  // site 0 sits on an even x (eta = +1), site 1 on an odd x (eta = -1): pairs start on even sites
  cplx h0 = stag_hop<false>(ux[0], ux_left, uy[0], uy_below[0], centre[1], left, above[0], below[0]);
  cplx h1 = stag_hop<true>(ux[1], ux[0], uy[1], uy_below[1], right, centre[0], above[1], below[1]);
  if (DAGGER) {
    h0 = fneg(h0);
    h1 = fneg(h1);
  }
  res[0] = fadd(fscale(0.5, h0), fscale(mass, centre[0]));
  res[1] = fadd(fscale(0.5, h1), fscale(mass, centre[1]));
}

template <bool FUSE>
struct NormLoad {  // everything fetched one row ahead: psi(y+2) (raw), U(y+1)
  cplx a[2];
  cplx b[2];
  cplx ux[2];
  cplx uy[2];
};

template <bool FUSE_XPAY, int NDOT>
__global__ void __launch_bounds__(NORM_THREADS) normal_kernel(const NormArgs a) {
  double beta = 0.0;
  if (a.cg != nullptr) {
    if (a.cg->done) return;
    if (FUSE_XPAY) beta = xdiv(a.cg->rsq_new, a.cg->rsq_old);  // generic_cg.cpp:344
  }
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;

  const int lane = threadIdx.x & 31;
  const int X = a.X, Y = a.Y;
  const int nstrips = (X + NORM_OUT_PER_WARP - 1) / NORM_OUT_PER_WARP;
  const long long units = (long long)nstrips * Y;
  const long long nwarps = (long long)gridDim.x * NORM_WARPS;
  const long long wid = (long long)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
  const long long u_begin = units * wid / nwarps, u_end = units * (wid + 1) / nwarps;

  auto wrap_row = [&](int y) -> size_t { return (size_t)(((y % Y) + Y) % Y) * X; };

  long long u = u_begin;
  while (u < u_end) {
    const int strip = (int)(u / Y);
    const int ya = (int)(u - (long long)strip * Y);
    const int yb = (int)min((long long)Y, (long long)ya + (u_end - u));
    u += (yb - ya);

    // this lane's pair of sites (periodic in x; pairs never straddle the seam because X is even)
    const int xs = strip * NORM_OUT_PER_WARP - 2 + 2 * lane;
    const int x0 = ((xs % X) + X) % X;
    const bool active = (lane >= 1) && (lane <= 30) && (xs < X);

    auto load_psi = [&](int y, cplx(&v)[2]) {  // the (possibly fused) input row y at this lane's pair
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
      const size_t o2 = wrap_row(y + 2) + x0, o1 = wrap_row(y + 1) + x0;
      if (FUSE_XPAY) {
        ldv<2>(a.r + o2, L.a);
        ldv<2>(a.pold + o2, L.b);
      } else {
        ldv<2>(a.in + o2, L.a);
      }
      ldv_nc<2>(a.Ux + o1, L.ux);
      ldv_nc<2>(a.Uy + o1, L.uy);
    };

    // ---- prologue: t(ya-1), t(ya) from psi(ya-2 .. ya+1)
    cplx p_c[2], p_n[2], t_m[2], t_c[2], ux_c[2], uy_m[2], uy_c[2];
    cplx uxl_c;
    {
      cplx p_mm[2], p_m[2], ux_m[2], uy_mm[2];
      load_psi(ya - 2, p_mm);
      load_psi(ya - 1, p_m);
      load_psi(ya, p_c);
      load_psi(ya + 1, p_n);
      ldv_nc<2>(a.Uy + wrap_row(ya - 2) + x0, uy_mm);
      ldv_nc<2>(a.Ux + wrap_row(ya - 1) + x0, ux_m);
      ldv_nc<2>(a.Uy + wrap_row(ya - 1) + x0, uy_m);
      ldv_nc<2>(a.Ux + wrap_row(ya) + x0, ux_c);
      ldv_nc<2>(a.Uy + wrap_row(ya) + x0, uy_c);
      const cplx uxl_m = shfl_up_c(ux_m[1], 1);
      uxl_c = shfl_up_c(ux_c[1], 1);
      stag_row<false>(t_m, p_mm, p_m, p_c, ux_m, uxl_m, uy_m, uy_mm, a.mass);  // t(ya-1)
      stag_row<false>(t_c, p_m, p_c, p_n, ux_c, uxl_c, uy_c, uy_m, a.mass);    // t(ya)
    }
    NormLoad<FUSE_XPAY> nxt;
    fetch(ya, nxt);

#pragma unroll 1
    for (int y = ya; y < yb; y++) {
      const NormLoad<FUSE_XPAY> cur = nxt;
      if (y + 1 < yb) fetch(y + 1, nxt);  // prefetch while this row is computed
      cplx p_nn[2];
      if (FUSE_XPAY) {
        p_nn[0] = fadd(cur.a[0], fscale(beta, cur.b[0]));
        p_nn[1] = fadd(cur.a[1], fscale(beta, cur.b[1]));
      } else {
        p_nn[0] = cur.a[0];
        p_nn[1] = cur.a[1];
      }
      const cplx uxl_n = shfl_up_c(cur.ux[1], 1);
      cplx t_n[2], res[2];
      stag_row<false>(t_n, p_c, p_n, p_nn, cur.ux, uxl_n, cur.uy, uy_c, a.mass);  // t(y+1) = D psi
      stag_row<true>(res, t_m, t_c, t_n, ux_c, uxl_c, uy_c, uy_m, a.mass);        // out(y) = D^dag t
      if (active) {
        const size_t o = (size_t)y * X + x0;
        stv<2>(a.out + o, res);
        if (FUSE_XPAY) stv<2>(a.pnew + o, p_c);
        if (NDOT >= 1) {
          cplx wv[2];
          if (a.w == nullptr) {
            wv[0] = p_c[0];
            wv[1] = p_c[1];
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
#pragma unroll
      for (int s = 0; s < 2; s++) {  // roll the windows
        p_c[s] = p_n[s];
        p_n[s] = p_nn[s];
        t_m[s] = t_c[s];
        t_c[s] = t_n[s];
        ux_c[s] = cur.ux[s];
        uy_m[s] = uy_c[s];
        uy_c[s] = cur.uy[s];
      }
      uxl_c = uxl_n;
    }
  }

  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total) && threadIdx.x == 0) {
      if (a.cg != nullptr && a.cg_role == 1) {  // <p,Ap> ready: generic_cg.cpp:326 / :345
        a.cg->pAp_re = total[0];
        a.cg->pAp_im = total[1];
        a.cg->rsq_old = a.cg->rsq_new;
      }
    }
  }
}

template <bool FUSE, int NDOT>
static int launch_normal_t(glb_operator* op, const NormArgs& a) {
  glb_context* ctx = op->ctx;
  auto kern = normal_kernel<FUSE, NDOT>;
  static int per_sm = 0;
  if (per_sm == 0) {
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NORM_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
  }
  const long long nstrips = (a.X + NORM_OUT_PER_WARP - 1) / NORM_OUT_PER_WARP;
  const long long units = nstrips * a.Y;
  long long blocks = (long long)ctx->sm_count * per_sm;
  const long long max_useful = (units + 8 * NORM_WARPS - 1) / (8 * NORM_WARPS);  // >= 8 rows per warp: 4 halo rows each
  if (blocks > max_useful) blocks = max_useful;
  if (blocks < 1) blocks = 1;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  kern<<<(unsigned)blocks, NORM_THREADS, 0, ctx->stream>>>(a);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

// can the one-pass kernel serve this operator?  (single rank, gauged, even X)
bool normal_fused_ok(const glb_operator* op) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GLB_NORMAL_FUSED");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return enabled && op->ctx->nranks == 1 && op->kind == OPK_STAGGERED && (op->flags & GLB_STAG_NORMAL) &&
         op->has_links && (op->X % 2 == 0) && op->X >= 2;
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
  a.Y = op->Y;
  a.mass = op->mass;
  a.red = ctx->red;
  if (!f.to_host) a.red.result_host = nullptr;
  a.cg = (CgState*)f.cg_state;
  a.cg_role = f.cg_role;
  const int ndot = (f.w != nullptr || f.w_is_input) ? (f.want_norm ? 2 : 1) : 0;
  if (fuse) {
    if (ndot == 0) return launch_normal_t<true, 0>(op, a);
    if (ndot == 1) return launch_normal_t<true, 1>(op, a);
    return launch_normal_t<true, 2>(op, a);
  }
  if (ndot == 0) return launch_normal_t<false, 0>(op, a);
  if (ndot == 1) return launch_normal_t<false, 1>(op, a);
  return launch_normal_t<false, 2>(op, a);
}

}  // namespace glb
