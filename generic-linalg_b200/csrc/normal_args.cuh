// normal_args.cuh -- kernel arguments shared by the two shapes of the one-pass D^dagger D kernel
// (normal.cu: two sites per thread; normal1.cu: one site per thread).
#pragma once
#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

struct NormArgs {
  const cplx* in;    // plain input, or nullptr when fused
  const cplx* r;     // fused direction update: input := r + beta * pold
  const cplx* pold;
  cplx* pnew;
  cplx* out;
  const cplx* Ux;
  const cplx* Uy;
  const cplx* w;     // dot partner; nullptr = the input itself
  // slabs: the two input rows below / above the slab (already final values, never fused);
  // nullptr on a single rank, where rows wrap periodically inside the slab
  const cplx* g_lo;
  const cplx* g_hi;
  int X, Y;          // Y = rows of this slab
  double mass;
  int nrb;      // row blocks: work item i -> strip i % nstrips, rows [Y*rb/nrb, Y*(rb+1)/nrb), rb = i / nstrips
  P2PRed pr;    // cg_role 3: the last block finishes the sum over ranks itself (peer memory)
  HaloWait wait;  // peer-memory slabs: ghost-row flags to wait for before touching g_lo / g_hi
  ReduceWs red;
  CgState* cg;
  int cg_role;
};

#ifdef __CUDACC__
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

#endif

// normal1.cu
int launch_normal_spt1(glb_operator* op, const NormArgs& a, bool fuse, int ndot, int variant);

}  // namespace glb
