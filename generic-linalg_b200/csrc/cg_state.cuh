// cg_state.cuh -- device-resident scalar state of the CG recurrence (generic_cg.cpp:324-354).
//
// alpha, beta and the stopping test live on the device: kernels read what their predecessors in
// the stream published here, so a CG iteration needs no host round trip.  Once `done` is set
// every later kernel of the (speculatively enqueued) batch returns immediately.
#pragma once

namespace glb {

struct CgState {
  double rsq_old;   // |r_k|^2 that the current search direction was built with ("rsq")
  double rsq_new;   // |r_{k+1}|^2 ("rsqNew")
  double pAp_re;    // <p,Ap>
  double pAp_im;
  double bnorm;     // sqrt(|b|^2)
  double eps;
  int iter;         // completed x/r updates (k+1)
  int max_iter;
  int done;         // stopping test fired
  int hit_max;      // ... because k == max_iter-1
  int hist_cap;
  int pad;
  // slab runs: this rank's partial sums, summed over ranks in place (ncclAllReduce on the stream)
  // before a one-thread kernel folds them into the recurrence: [0] |r|^2, [1..2] <p,Ap>
  double partial[4];
  // single-kernel iteration (cgstep.cu): the scalars the NEXT step starts from
  double alpha_re, alpha_im;  // alpha_i = |r_i|^2 / <p_i, q_i>
  double beta;                // beta_{i+1} = rsq_pred / |r_i|^2
  double rsq_pred;            // predicted |r_{i+1}|^2 = |r_i|^2 - 2 Re(alpha <r_i,q_i>) + |alpha|^2 |q_i|^2
  double pred_err;            // largest relative deviation of a prediction from the exact sum one step later
  int step;                   // steps completed (step 0 is the set-up pass)
  int pad2;
};

// Device-resident BiCGStab / CR (krylov.cu).  `cg` first: the operator kernels take this object as their CgState
// (early exit on cg.done).  The staggered kernel's fused BiCGStab inputs (stencil.cu, FUSE 2 / 3) read rho, beta and
// omega here and leave alpha.
struct KrylovState {
  CgState cg;       // done, iter, max_iter, eps, bnorm, hit_max, hist_cap, rsq_new
  double rho[2];    // BiCGStab: <r0,r> the current direction was built with ; CR: |Ap|^2 in rho[0]
  double alpha[2];  // BiCGStab: alpha of the current iteration (formed where s is formed, used by the x/r update);
                    // CR: <Ap,r> of the coming iteration
  double omega[2];
  double beta[2];
};

}  // namespace glb
