// cgstep.cuh -- arguments of the single-kernel CG iteration (cgstep.cu) and its launchers.
#pragma once
#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

struct CgStepArgs {
  // ping-pong vectors of the recurrence (q = D^dag D p): step s reads set s&1 and writes set (s&1)^1;
  // the solution x is updated in place
  cplx* r[2];
  cplx* q[2];
  cplx* p[2];
  cplx* x;
  const cplx* Ux;
  const cplx* Uy;
  // slabs over peer memory (nullptr on a single rank, where rows wrap inside the slab).  Every rank's arena holds,
  // at the same offset, [parity 0: lo | hi][parity 1: lo | hi] with lo / hi = [vector r, q, p][row 0..1][X]:
  // lo = this rank's rows -2, -1 (written by the rank below), hi = rows Y, Y+1 (written by the rank above).
  // Step s has exchange number seq_base + s: it reads parity (seq-1)&1 here, writes parity seq&1 there.
  const cplx* ghost;  // this rank's buffers
  cplx* peer_down;    // the same buffers of the rank below / above (remote pointers)
  cplx* peer_up;
  unsigned long long* flag_down;  // raised to seq once all boundary rows of that side are stored
  unsigned long long* flag_up;
  unsigned int* push_count;  // two local counters, self-resetting
  unsigned long long seq_base;
  HaloWait wait;  // local flags the boundary row blocks wait for (its seq is unused); budget also bounds the step spin
  int X, Y;       // Y = rows of this slab
  double mass;
  int nstrips, nrb;  // filled in by the launcher
  int nsteps;        // CG steps this launch may run (1: one kernel per iteration; > 1: persistent, cooperative launch)
  CgState* st;
  double* hist;
  ReduceWs red;
  P2PRed pr;                  // seq = number of step 0's rank-wide reduction (step s uses seq + s); 0 on a single rank
  const int* rb_rows;         // guided schedule: first row of every row block (nrb + 1 entries), else nullptr
  unsigned int* counters;     // [1] blocks that posted the step's sums, [2] blocks whose stores are in; zero at launch
  unsigned int* queue;        // dynamic schedule: item counter (nullptr = static partition), reset by the last block
  unsigned long long* step_ns;  // optional: %globaltimer at the end of every step (measurement), step_ns_cap entries
  int step_ns_cap;
  int trace_step;             // measurement aid (GLB_CGSTEP_TRACE): per-CTA record of this step
  unsigned long long* trace;
};

bool cg_step_ok(const glb_operator* op);
bool cg_step_persistent();
int launch_cg_step(glb_operator* op, const CgStepArgs& a);
int launch_cg_step_halo_init(glb_operator* op, const void* r, const void* q, const void* p, const CgStepArgs& a,
                             unsigned long long seq, unsigned int* ticket);

}  // namespace glb
