// stencil.cu -- nearest-neighbour lattice operators on a y-slab, complex<double>, sm_100a.
//
// One kernel template serves the whole gauged family of the reference:
//   square_staggered_u1        operators.cpp:184   D
//   square_staggered_dagger_u1 operators.cpp:372   D^dagger      (hopping sign flipped)
//   square_staggered_gamma5_u1 operators.cpp:316   gamma5 D      (site parity sign)
//   square_staggered / _gamma5 operators.cpp:127,262  free field (no link loads)
//   square_laplace_u1          operators.cpp:73    gauged Laplacian
//
// Design (HBM-bound: 64 B/site compulsory = 16 psi + 32 links + 16 out):
//   * persistent grid: #blocks = #SMs x resident blocks; the (strip, row) work units are split
//     evenly and statically, so there is no tail wave;
//   * a block owns a strip of blockDim*SPT consecutive x sites and marches over its rows keeping
//     psi(y-1), psi(y), psi(y+1) and U_y(y-1) in registers: every psi / link element is loaded
//     ONCE per strip (plus 2 halo rows per segment), 16-byte loads, 512 B or 1 KB per warp row;
//   * the row after next is prefetched into registers while the current row is computed;
//   * x neighbours come from warp shuffles; only the two edge lanes of a warp touch memory
//     (L1 hits: the neighbouring warp of the same block loaded that line one step earlier);
//   * links are stored as two site-major planes (Ux, Uy) so every access is unit-stride;
//   * optional fusions: input formed on the fly as r + beta*p_old and written out as p_new
//     (CG direction update, generic_cg.cpp:348-351), and dot products <w,out>, |out|^2 accumulated
//     in the epilogue with warp shuffles (generic_cg.cpp:326) -- no extra pass over HBM.
//   * arithmetic follows the reference's expression order without FMA contraction: the result is
//     bit-identical to the CPU code.
#include <cstdlib>

#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

constexpr int STAG_THREADS = 128;

struct StagArgs {
  // plain input (rows 0..Yloc-1), and the rows just below / above the slab
  const cplx* in;
  const cplx* in_lo;
  const cplx* in_hi;
  // fused direction update: input := r + beta * pold (beta from the device CG state)
  const cplx* r;
  const cplx* r_lo;
  const cplx* r_hi;
  const cplx* pold;
  const cplx* pold_lo;
  const cplx* pold_hi;
  cplx* pnew;
  cplx* out;
  const cplx* Ux;
  const cplx* Uy;
  const cplx* Uy_lo;  // U_y of the row below the slab
  const cplx* w;      // dot partner; nullptr = use the input itself
  int X, Yloc, y0;
  double mass;   // staggered: m ; laplace_u1: 4+m
  int dagger;    // flip the hopping sign
  int gamma5;    // multiply the result by (-1)^(x+y)
  ReduceWs red;
  CgState* cg;
  int cg_role;   // 1: epilogue publishes <p,Ap> into the CG state
  int nrb;       // number of row blocks per strip
  HaloWait wait; // slabs over peer memory: the ghost rows' flags; only row blocks that touch a ghost row wait for them
  // FAM_STAG_EO: the even/odd pieces of operators.cpp:456-616
  int eo_parity;    // 0: update even sites (D_eo), 1: update odd sites (D_oe); the other parity gets the `else` value
  int eo_post;      // 0: out = h/2 | 0          (square_staggered_deo_u1 / _doe_u1)
                    // 1: out = coef*aux - h/2 | 0          (m2mdeodoe: coef = m^2; eoprec_prepare: coef = m)
                    // 2: out = coef*(aux - h/2) | in       (eoprec_reconstruct: coef = 1/m)
  double eo_coef;
  const cplx* aux;  // second input of eo_post 1 / 2 (same layout as out)
};

enum { FAM_STAGGERED = 0, FAM_LAPLACE_U1 = 1, FAM_STAG_EO = 2 };

template <int SPT>
struct RowLoad {  // everything fetched one row ahead
  cplx a[SPT];    // psi(y+1)  (or r(y+1) when fused)
  cplx b[SPT];    // pold(y+1) when fused
  cplx ux[SPT];   // U_x(y)
  cplx uy[SPT];   // U_y(y)
};

template <int SPT, bool HAS_U, bool FUSE_XPAY, int NDOT, int FAM, int PF>
__global__ void __launch_bounds__(STAG_THREADS) stag_kernel(const StagArgs a) {
  double beta = 0.0;
  if (a.cg != nullptr) {
    if (a.cg->done) return;
    if (FUSE_XPAY) beta = xdiv(a.cg->rsq_new, a.cg->rsq_old);  // generic_cg.cpp:344
  }
  const int lane = threadIdx.x & 31;
  const int strip_w = STAG_THREADS * SPT;
  const int nstrips = (a.X + strip_w - 1) / strip_w;
  // work item i -> strip i % nstrips, row block i / nstrips: neighbouring blocks sweep neighbouring
  // strips over the same rows at the same time (whole lattice rows stream through DRAM and the
  // edge-lane loads of a block hit lines its neighbours fetch)
  const long long nitems = (long long)nstrips * a.nrb;

  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;

  const int X = a.X, Yloc = a.Yloc;
  // row pointer selection: y=-1 -> *_lo, y=Yloc -> *_hi
  auto rowp = [&](const cplx* base, const cplx* lo, const cplx* hi, int y) -> const cplx* {
    return (y < 0) ? lo : ((y >= Yloc) ? hi : base + (size_t)y * X);
  };
  auto psi_at = [&](int y, int x) -> cplx {  // the (possibly fused) input at one site
    if (FUSE_XPAY) {
      const cplx rr = rowp(a.r, a.r_lo, a.r_hi, y)[x];
      const cplx pp = rowp(a.pold, a.pold_lo, a.pold_hi, y)[x];
      return fadd(rr, fscale(beta, pp));
    } else {
      return rowp(a.in, a.in_lo, a.in_hi, y)[x];
    }
  };

  // a.nrb == 0 selects the alternative partition: one contiguous run of (strip,row) units per block
  const bool lockstep = a.nrb > 0;
  const long long units = (long long)nstrips * Yloc;
  const long long u_end = lockstep ? nitems : units * (blockIdx.x + 1) / gridDim.x;
  long long it = lockstep ? (long long)blockIdx.x : units * blockIdx.x / gridDim.x;
  while (it < u_end) {
    int strip, ya, yb;
    if (lockstep) {
      strip = (int)(it % nstrips);
      const int rb = (int)(it / nstrips);
      ya = (int)((long long)Yloc * rb / a.nrb);
      yb = (int)((long long)Yloc * (rb + 1) / a.nrb);
      it += gridDim.x;
      if (ya >= yb) continue;
      if (a.wait.seq != 0 && (ya == 0 || yb == Yloc)) {
        // this row block reads a ghost row: the neighbour's push of this exchange must have landed (interior row
        // blocks never wait -- the exchange overlaps with their work)
        if (threadIdx.x == 0) {
          if (ya == 0) spin_until(a.wait.flag_lo, a.wait.seq, a.wait.budget);
          if (yb == Yloc) spin_until(a.wait.flag_hi, a.wait.seq, a.wait.budget);
        }
        __syncthreads();
      }
    } else {
      if (a.wait.seq != 0) {  // contiguous partition: any block may touch a ghost row
        if (threadIdx.x == 0) {
          spin_until(a.wait.flag_lo, a.wait.seq, a.wait.budget);
          spin_until(a.wait.flag_hi, a.wait.seq, a.wait.budget);
        }
        __syncthreads();
      }
      strip = (int)(it / Yloc);
      ya = (int)(it - (long long)strip * Yloc);
      yb = (int)min((long long)Yloc, (long long)ya + (u_end - it));
      it += (yb - ya);
    }

    int x0 = strip * strip_w + threadIdx.x * SPT;
    const bool active = x0 < X;
    if (!active) x0 = X - SPT;  // keep addresses legal; results are discarded
    const int xl = (x0 == 0) ? X - 1 : x0 - 1;
    const int xr = (x0 + SPT >= X) ? 0 : x0 + SPT;
    const bool edge_l = (lane == 0);
    const bool edge_r = (lane == 31) || (x0 + SPT >= X);

    // ---- prologue: rows ya-1 and ya of the input, U_y of row ya-1
    cplx m[SPT], c[SPT], uym[SPT];
    auto psi_row = [&](int y, cplx(&v)[SPT]) {  // own sites of row y, one 16/32-byte access per array
      if (FUSE_XPAY) {
        cplx rr[SPT], pp[SPT];
        ldv<SPT>(rowp(a.r, a.r_lo, a.r_hi, y) + x0, rr);
        ldv<SPT>(rowp(a.pold, a.pold_lo, a.pold_hi, y) + x0, pp);
#pragma unroll
        for (int s = 0; s < SPT; s++) v[s] = fadd(rr[s], fscale(beta, pp[s]));
      } else {
        ldv<SPT>(rowp(a.in, a.in_lo, a.in_hi, y) + x0, v);
      }
    };
    psi_row(ya - 1, m);
    psi_row(ya, c);
    if (HAS_U) ldv_nc<SPT>((ya == 0 ? a.Uy_lo : a.Uy + (size_t)(ya - 1) * X) + x0, uym);
    auto fetch = [&](int y, RowLoad<SPT>& L) {  // loads for centre row y: psi(y+1), U(y)
      if (FUSE_XPAY) {
        ldv<SPT>(rowp(a.r, a.r_lo, a.r_hi, y + 1) + x0, L.a);
        ldv<SPT>(rowp(a.pold, a.pold_lo, a.pold_hi, y + 1) + x0, L.b);
      } else {
        ldv<SPT>(rowp(a.in, a.in_lo, a.in_hi, y + 1) + x0, L.a);
      }
      if (HAS_U) {
        ldv_nc<SPT>(a.Ux + (size_t)y * X + x0, L.ux);
        ldv_nc<SPT>(a.Uy + (size_t)y * X + x0, L.uy);
      }
    };
    // PF rows of loads are kept in flight per thread (PF = 1: next row only)
    RowLoad<SPT> st[PF];
#pragma unroll
    for (int k = 0; k < PF; k++)
      if (ya + k < yb) fetch(ya + k, st[k]);

    auto row_body = [&](const int y, const RowLoad<SPT>& cur) {
      // edge lanes: x neighbours that live in another warp / across the periodic seam
      cplx cl, cr, uxl;
      if (edge_l) {
        cl = psi_at(y, xl);
        if (HAS_U) uxl = __ldg(&a.Ux[(size_t)y * X + xl]);
      }
      if (edge_r) cr = psi_at(y, xr);
      {
        const cplx t_l = shfl_up_c(c[SPT - 1], 1);
        const cplx t_r = shfl_down_c(c[0], 1);
        if (!edge_l) cl = t_l;
        if (!edge_r) cr = t_r;
        if (HAS_U) {
          const cplx t_u = shfl_up_c(cur.ux[SPT - 1], 1);
          if (!edge_l) uxl = t_u;
        }
      }
      cplx p[SPT];
#pragma unroll
      for (int s = 0; s < SPT; s++) p[s] = FUSE_XPAY ? fadd(cur.a[s], fscale(beta, cur.b[s])) : cur.a[s];

      const int yg = a.y0 + y;
      cplx resv[SPT];
#pragma unroll
      for (int s = 0; s < SPT; s++) {
        const cplx psi_xp = (s + 1 < SPT) ? c[s + 1] : cr;
        const cplx psi_xm = (s > 0) ? c[s - 1] : cl;
        const int x = x0 + s;
        cplx h = mk(0.0, 0.0);
        cplx res;
        if (FAM == FAM_STAGGERED || FAM == FAM_STAG_EO) {
          const bool eta_neg = (x & 1);  // eta1 = 1 - 2*(x%2), operators.cpp:212
          if (HAS_U) {
            const cplx ux_m = (s > 0) ? cur.ux[s - 1] : uxl;
            h = fsub(h, fmul(cur.ux[s], psi_xp));   // -   U_x(x,y)     psi(x+1,y)
            h = fadd(h, fcmul(ux_m, psi_xm));       // + conj U_x(x-1,y) psi(x-1,y)
            const cplx t3 = fmul(cur.uy[s], p[s]);  // -eta U_y(x,y)     psi(x,y+1)
            h = eta_neg ? fadd(h, t3) : fsub(h, t3);
            const cplx t4 = fcmul(uym[s], m[s]);    // +eta conj U_y(x,y-1) psi(x,y-1)
            h = eta_neg ? fsub(h, t4) : fadd(h, t4);
          } else {
            h = fsub(h, psi_xp);
            h = fadd(h, psi_xm);
            h = eta_neg ? fadd(h, p[s]) : fsub(h, p[s]);
            h = eta_neg ? fsub(h, m[s]) : fadd(h, m[s]);
          }
          if (FAM == FAM_STAG_EO) {
            // operators.cpp:456-616: hopping term only, on one parity; the other parity is zeroed (deo/doe,
            // m2mdeodoe, prepare) or copied from the input (reconstruct)
            h = fscale(0.5, h);                                       // :486 / :521
            const bool upd = (((x + yg) & 1) == a.eo_parity);
            if (a.eo_post == 0) {
              res = upd ? h : mk(0.0, 0.0);
            } else {
              const cplx ax = a.aux[(size_t)y * X + x];
              if (a.eo_post == 1)
                res = upd ? fsub(fscale(a.eo_coef, ax), h) : mk(0.0, 0.0);   // :541, :564
              else
                res = upd ? fscale(a.eo_coef, fsub(ax, h)) : c[s];          // :589-595
            }
          } else {
          if (a.dagger) h = fneg(h);               // operators.cpp:405-414: every hop changes sign
          h = fscale(0.5, h);                      // operators.cpp:227
          res = fadd(h, fscale(a.mass, c[s]));     // operators.cpp:231
          if (a.gamma5 && ((x + yg) & 1)) res = fneg(res);  // operators.cpp:345 eo_sign on every term
          }
        } else {  // gauged Laplacian, operators.cpp:103-116
          const cplx ux_m = (s > 0) ? cur.ux[s - 1] : uxl;
          h = fsub(h, fmul(cur.ux[s], psi_xp));
          h = fsub(h, fcmul(ux_m, psi_xm));
          h = fsub(h, fmul(cur.uy[s], p[s]));
          h = fsub(h, fcmul(uym[s], m[s]));
          res = fadd(h, fscale(a.mass, c[s]));     // a.mass carries (4+mass)
        }
        resv[s] = res;
        if (active) {
          if (NDOT >= 1) {
            const cplx wv = (a.w == nullptr) ? c[s] : a.w[(size_t)y * X + x];
            Field<cplx>::dot_acc(acc, wv, res);
          }
          if (NDOT >= 2) acc[2] += fnorm(res);
        }
      }
      if (active) {
        stv<SPT>(a.out + (size_t)y * X + x0, resv);
        if (FUSE_XPAY) stv<SPT>(a.pnew + (size_t)y * X + x0, c);
      }
      // roll the window
#pragma unroll
      for (int s = 0; s < SPT; s++) {
        m[s] = c[s];
        c[s] = p[s];
        if (HAS_U) uym[s] = cur.uy[s];
      }
    };

#pragma unroll 1
    for (int y = ya; y < yb; y += PF) {
#pragma unroll
      for (int k = 0; k < PF; k++) {
        const int yy = y + k;
        if (yy < yb) {
          const RowLoad<SPT> cur = st[k];
          if (yy + PF < yb) fetch(yy + PF, st[k]);  // refill this stage while the row is computed
          row_body(yy, cur);
        }
      }
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

// ------------------------------------------------------------------------------------------
// gamma_5 alone (operators.cpp:242-259) and the free Laplacian with Nc colours
// (square_laplace.cpp:182, imag_laplace.cpp:126, operators.cpp:28, multishift.cpp:634).
// Simple one-thread-per-element kernels: these are not on the headline path; neighbours are
// served by L1/L2.
// ------------------------------------------------------------------------------------------
__global__ void gamma5_kernel(cplx* out, const cplx* in, int X, int Yloc, int y0) {
  const size_t n = (size_t)X * Yloc;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % X), y = (int)(i / X) + y0;
    const cplx v = in[i];
    out[i] = ((x + y) & 1) ? fneg(v) : v;
  }
}

template <typename T>
struct LapArgs {
  const T* in;
  const T* in_lo;
  const T* in_hi;
  T* out;
  const T* w;
  int X, Yloc, Nc;
  T diag;
  int stag_free;  // 1: free staggered stencil (tests/multishift/multishift.cpp:677), diag carries the mass
  ReduceWs red;
  CgState* cg;
  int cg_role;
};

template <typename T, int NDOT>
__global__ void __launch_bounds__(256) laplace_kernel(const LapArgs<T> a) {
  if (a.cg != nullptr && a.cg->done) return;
  constexpr int NC = Field<T>::NCOMP;
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? NC : NC + 1);
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;
  const int RX = a.X * a.Nc;  // row length in elements
  const size_t n = (size_t)RX * a.Yloc;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / RX);
    const int e = (int)(i - (size_t)y * RX);  // x*Nc + c
    const int ep = (e + a.Nc >= RX) ? e + a.Nc - RX : e + a.Nc;
    const int em = (e - a.Nc < 0) ? e - a.Nc + RX : e - a.Nc;
    const T* row = a.in + (size_t)y * RX;
    const T* rowp = (y + 1 >= a.Yloc) ? a.in_hi : row + RX;
    const T* rowm = (y == 0) ? a.in_lo : row - RX;
    T h = Field<T>::zero();
    const T self = row[e];
    T res;
    if (a.stag_free) {  // -(x+1) + (x-1) - eta (y+1) + eta (y-1), halved, + m*self ; eta = (-1)^x
      const bool eta_neg = ((e / a.Nc) & 1);
      h = fsub(h, row[ep]);
      h = fadd(h, row[em]);
      h = eta_neg ? fadd(h, rowp[e]) : fsub(h, rowp[e]);
      h = eta_neg ? fsub(h, rowm[e]) : fadd(h, rowm[e]);
      res = fadd(fscale(0.5, h), fmul(a.diag, self));
    } else {
      h = fsub(h, row[ep]);   // + e1
      h = fsub(h, row[em]);   // - e1
      h = fsub(h, rowp[e]);   // + e2
      h = fsub(h, rowm[e]);   // - e2
      res = fadd(h, fmul(a.diag, self));
    }
    a.out[i] = res;
    if (NDOT >= 1) {
      const T wv = (a.w == nullptr) ? self : a.w[i];
      Field<T>::dot_acc(acc, wv, res);
    }
    if (NDOT >= 2) acc[NC] += fnorm(res);
  }
  if (NDOT > 0) {
    double total[NRED];
    if (grid_sum<NRED>(acc, a.red, total) && threadIdx.x == 0) {
      if (a.cg != nullptr && a.cg_role == 1) {
        a.cg->pAp_re = total[0];
        a.cg->pAp_im = (NC == 2) ? total[NC - 1] : 0.0;
        a.cg->rsq_old = a.cg->rsq_new;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// tunables (environment, read once): GLB_STAG_PF = rows of loads in flight per thread (1 or 2),
// GLB_STAG_SPT = sites per thread (1 or 2).  Defaults are the measured-best values.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int stag_pf() {
  static int v = env_int("GLB_STAG_PF", 1);
  return v == 2 ? 2 : 1;
}
static int stag_spt() {
  static int v = env_int("GLB_STAG_SPT", 2);
  return v == 1 ? 1 : 2;
}

template <int SPT, bool HAS_U, bool FUSE, int NDOT, int FAM, int PF>
static int launch_stag_t(glb_operator* op, const StagArgs& a) {
  glb_context* ctx = op->ctx;
  auto kern = stag_kernel<SPT, HAS_U, FUSE, NDOT, FAM, PF>;
  static int per_sm = 0;  // one value per instantiation
  if (per_sm == 0) {
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STAG_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
  }
  const int strip_w = STAG_THREADS * SPT;
  const long long nstrips = (a.X + strip_w - 1) / strip_w;
  const long long cap = (long long)ctx->sm_count * per_sm;
  // row blocks: enough to fill the resident grid, at least 4 rows each (2 halo rows per item)
  long long nrb = cap / nstrips;
  const long long nrb_cap = a.Yloc >= 8 ? a.Yloc / 4 : 1;
  if (nrb > nrb_cap) nrb = nrb_cap;
  if (nrb < 1) nrb = 1;
  StagArgs b = a;
  static int lockstep = env_int("GLB_STAG_LOCKSTEP", 1);
  b.nrb = lockstep ? (int)nrb : 0;
  long long blocks = lockstep ? nstrips * nrb : (nstrips * a.Yloc + 3) / 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (blocks > MAX_PARTIAL_BLOCKS) blocks = MAX_PARTIAL_BLOCKS;
  // in + out (+ the link pair) per site; the fused direction update adds r, p_old in and p_new out
  ProfScope prof(ctx, PROF_STAG, (double)a.X * a.Yloc * ((HAS_U ? 64.0 : 32.0) + (FUSE ? 32.0 : 0.0)));
  kern<<<(unsigned)blocks, STAG_THREADS, 0, ctx->stream>>>(b);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

template <int SPT, bool HAS_U, int FAM, int PF>
static int launch_stag_p(glb_operator* op, const StagArgs& a, bool fuse, int ndot) {
  if (fuse) {
    if (ndot == 0) return launch_stag_t<SPT, HAS_U, true, 0, FAM, PF>(op, a);
    if (ndot == 1) return launch_stag_t<SPT, HAS_U, true, 1, FAM, PF>(op, a);
    return launch_stag_t<SPT, HAS_U, true, 2, FAM, PF>(op, a);
  }
  if (ndot == 0) return launch_stag_t<SPT, HAS_U, false, 0, FAM, PF>(op, a);
  if (ndot == 1) return launch_stag_t<SPT, HAS_U, false, 1, FAM, PF>(op, a);
  return launch_stag_t<SPT, HAS_U, false, 2, FAM, PF>(op, a);
}

template <int SPT, bool HAS_U, int FAM>
static int launch_stag_f(glb_operator* op, const StagArgs& a, bool fuse, int ndot) {
  if (stag_pf() == 2) return launch_stag_p<SPT, HAS_U, FAM, 2>(op, a, fuse, ndot);
  return launch_stag_p<SPT, HAS_U, FAM, 1>(op, a, fuse, ndot);
}

int launch_staggered(glb_operator* op, void* out, const void* in, bool dagger, const ApplyFusion& f) {
  glb_context* ctx = op->ctx;
  if (op->X < 2 || op->Yloc < 1) return fail(GLB_ERR_ARG, "staggered/gauged stencil needs X >= 2");
  StagArgs a{};
  const size_t X = op->X;
  const bool single = (ctx->nranks == 1);
  auto lo_of = [&](const void* base, const void* ghost) -> const cplx* {  // row -1: the LAST ghost row below
    return single ? (const cplx*)base + (size_t)(op->Yloc - 1) * X
                  : (const cplx*)ghost + (size_t)(op->ghost_depth - 1) * X;
  };
  auto hi_of = [&](const void* base, const void* ghost) -> const cplx* {
    return single ? (const cplx*)base : (const cplx*)ghost;
  };
  const bool fuse = (f.r != nullptr);
  if (fuse) {
    if (!single) return fail(GLB_ERR_STATE, "fused direction update is single-rank only (slab path exchanges p first)");
    a.r = (const cplx*)f.r;
    a.r_lo = lo_of(f.r, nullptr);
    a.r_hi = hi_of(f.r, nullptr);
    a.pold = (const cplx*)f.p_old;
    a.pold_lo = lo_of(f.p_old, nullptr);
    a.pold_hi = hi_of(f.p_old, nullptr);
    a.pnew = (cplx*)f.p_new;
  } else {
    a.in = (const cplx*)in;
    a.in_lo = lo_of(in, op->ghost_lo);
    a.in_hi = hi_of(in, op->ghost_hi);
  }
  a.out = (cplx*)out;
  a.Ux = op->Ux;
  a.Uy = op->Uy;
  a.Uy_lo = op->Uy_lo;
  a.w = f.w_is_input ? nullptr : (const cplx*)f.w;
  a.X = op->X;
  a.Yloc = op->Yloc;
  a.y0 = op->y0;
  a.dagger = dagger ? 1 : 0;
  a.gamma5 = (op->flags & GLB_STAG_GAMMA5) ? 1 : 0;
  a.red = ctx->red;
  if (!f.to_host) a.red.result_host = nullptr;
  a.wait = f.wait;
  a.cg = (CgState*)f.cg_state;
  a.cg_role = f.cg_role;
  const int ndot = (f.w != nullptr || f.w_is_input) ? (f.want_norm ? 2 : 1) : 0;
  if (ndot == 0 && f.want_norm) return fail(GLB_ERR_ARG, "want_norm requires a dot partner");
  const bool spt2 = (op->X % 2 == 0) && stag_spt() == 2;
  if (op->kind == OPK_LAPLACE_U1) {
    a.mass = 4 + op->mass;  // operators.cpp:116 : (4+mass), int + double
    if (spt2) return launch_stag_f<2, true, FAM_LAPLACE_U1>(op, a, fuse, ndot);
    return launch_stag_f<1, true, FAM_LAPLACE_U1>(op, a, fuse, ndot);
  }
  a.mass = op->mass;
  if (op->has_links) {
    if (spt2) return launch_stag_f<2, true, FAM_STAGGERED>(op, a, fuse, ndot);
    return launch_stag_f<1, true, FAM_STAGGERED>(op, a, fuse, ndot);
  }
  if (spt2) return launch_stag_f<2, false, FAM_STAGGERED>(op, a, fuse, ndot);
  return launch_stag_f<1, false, FAM_STAGGERED>(op, a, fuse, ndot);
}

// even/odd pieces (operators.cpp:456-616): hopping term on one parity with an optional second input
int launch_staggered_eo(glb_operator* op, void* out, const void* in, int parity, int post, double coef, const void* aux,
                        const ApplyFusion& f) {
  glb_context* ctx = op->ctx;
  if (op->X < 2 || op->Yloc < 1 || !op->has_links) return fail(GLB_ERR_ARG, "even/odd staggered pieces need links and X >= 2");
  if (f.r != nullptr) return fail(GLB_ERR_ARG, "even/odd staggered pieces have no fused direction update");
  if (post != 0 && aux == nullptr) return fail(GLB_ERR_ARG, "even/odd post-operation needs its second input");
  StagArgs a{};
  const size_t X = op->X;
  const bool single = (ctx->nranks == 1);
  a.in = (const cplx*)in;
  a.in_lo = single ? (const cplx*)in + (size_t)(op->Yloc - 1) * X : (const cplx*)op->ghost_lo + (size_t)(op->ghost_depth - 1) * X;
  a.in_hi = single ? (const cplx*)in : (const cplx*)op->ghost_hi;
  a.out = (cplx*)out;
  a.Ux = op->Ux;
  a.Uy = op->Uy;
  a.Uy_lo = op->Uy_lo;
  a.w = f.w_is_input ? nullptr : (const cplx*)f.w;
  a.X = op->X;
  a.Yloc = op->Yloc;
  a.y0 = op->y0;
  a.mass = op->mass;
  a.red = ctx->red;
  if (!f.to_host) a.red.result_host = nullptr;
  a.wait = f.wait;
  a.cg = (CgState*)f.cg_state;
  a.cg_role = f.cg_role;
  a.eo_parity = parity;
  a.eo_post = post;
  a.eo_coef = coef;
  a.aux = (const cplx*)aux;
  const int ndot = (f.w != nullptr || f.w_is_input) ? (f.want_norm ? 2 : 1) : 0;
  const bool spt2 = (op->X % 2 == 0) && stag_spt() == 2;
  if (spt2) {
    if (ndot == 0) return launch_stag_t<2, true, false, 0, FAM_STAG_EO, 1>(op, a);
    if (ndot == 1) return launch_stag_t<2, true, false, 1, FAM_STAG_EO, 1>(op, a);
    return launch_stag_t<2, true, false, 2, FAM_STAG_EO, 1>(op, a);
  }
  if (ndot == 0) return launch_stag_t<1, true, false, 0, FAM_STAG_EO, 1>(op, a);
  if (ndot == 1) return launch_stag_t<1, true, false, 1, FAM_STAG_EO, 1>(op, a);
  return launch_stag_t<1, true, false, 2, FAM_STAG_EO, 1>(op, a);
}

int launch_gamma5(glb_operator* op, void* out, const void* in) {
  glb_context* ctx = op->ctx;
  const size_t n = (size_t)op->X * op->Yloc;
  const int grid = blas_grid(ctx, n, 256, 2);
  gamma5_kernel<<<grid, 256, 0, ctx->stream>>>((cplx*)out, (const cplx*)in, op->X, op->Yloc, op->y0);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

template <typename T>
static int launch_laplace_t(glb_operator* op, void* out, const void* in, const ApplyFusion& f, T diag) {
  glb_context* ctx = op->ctx;
  LapArgs<T> a{};
  const size_t RX = (size_t)op->X * op->nc;
  const bool single = (ctx->nranks == 1);
  a.in = (const T*)in;
  a.in_lo = single ? (const T*)in + (size_t)(op->Yloc - 1) * RX : (const T*)op->ghost_lo;
  a.in_hi = single ? (const T*)in : (const T*)op->ghost_hi;
  a.out = (T*)out;
  a.w = f.w_is_input ? nullptr : (const T*)f.w;
  a.X = op->X;
  a.Yloc = op->Yloc;
  a.Nc = op->nc;
  a.diag = diag;
  a.stag_free = (op->flags & 0x100u) ? 1 : 0;
  a.red = ctx->red;
  if (!f.to_host) a.red.result_host = nullptr;
  a.cg = (CgState*)f.cg_state;
  a.cg_role = f.cg_role;
  const int ndot = (f.w != nullptr || f.w_is_input) ? (f.want_norm ? 2 : 1) : 0;
  const int grid = blas_grid(ctx, RX * op->Yloc, 256, 1);
  ProfScope prof(ctx, PROF_LAPLACE, (double)RX * op->Yloc * sizeof(T) * 2);
  if (ndot == 0)
    laplace_kernel<T, 0><<<grid, 256, 0, ctx->stream>>>(a);
  else if (ndot == 1)
    laplace_kernel<T, 1><<<grid, 256, 0, ctx->stream>>>(a);
  else
    laplace_kernel<T, 2><<<grid, 256, 0, ctx->stream>>>(a);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_laplace(glb_operator* op, void* out, const void* in, const ApplyFusion& f) {
  if (op->dtype == GLB_COMPLEX)
    return launch_laplace_t<cplx>(op, out, in, f, make_double2(op->diag_re, op->diag_im));
  return launch_laplace_t<double>(op, out, in, f, op->diag_re);
}

}  // namespace glb
