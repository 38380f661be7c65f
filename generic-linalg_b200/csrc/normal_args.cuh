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

// normal1.cu
int launch_normal_spt1(glb_operator* op, const NormArgs& a, bool fuse, int ndot, int variant);

}  // namespace glb
