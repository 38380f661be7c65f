// cgstep.cuh -- arguments of the single-kernel CG iteration (cgstep.cu) and its launchers.
#pragma once
#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

struct CgStepArgs {
  // ping-pong vectors of the recurrence (q = D^dag D p) and the solution, updated in place
  const cplx* r_in;
  const cplx* q_in;
  const cplx* p_in;
  cplx* x;
  cplx* r_out;
  cplx* q_out;
  cplx* p_out;
  const cplx* Ux;
  const cplx* Uy;
  // slabs: the neighbours' boundary rows of (r, q, p) written by the PREVIOUS step, layout [vector][row 0..1][X];
  // g_lo = rows -2, -1, g_hi = rows Y, Y+1; nullptr on a single rank (rows wrap inside the slab)
  const cplx* g_lo;
  const cplx* g_hi;
  // where THIS step's boundary rows go (the neighbours' ghost buffers of the other parity, remote pointers)
  cplx* push_down;  // rows 0, 1      -> g_hi of the rank below
  cplx* push_up;    // rows Y-2, Y-1  -> g_lo of the rank above
  unsigned long long* flag_down;  // raised (to push_seq) once all boundary rows of that side are stored
  unsigned long long* flag_up;
  unsigned int* push_count;  // two local counters, self-resetting
  unsigned long long push_seq;
  HaloWait wait;  // local flags the boundary row blocks wait for before reading g_lo / g_hi
  int X, Y;       // Y = rows of this slab
  double mass;
  int nstrips, nrb;  // filled in by the launcher
  CgState* st;
  double* hist;
  ReduceWs red;
  P2PRed pr;  // seq == 0 on a single rank
};

bool cg_step_ok(const glb_operator* op);
int launch_cg_step(glb_operator* op, const CgStepArgs& a);
int launch_cg_step_halo_init(glb_operator* op, const void* r, const void* q, const void* p, const CgStepArgs& a,
                             unsigned int* ticket);

}  // namespace glb
