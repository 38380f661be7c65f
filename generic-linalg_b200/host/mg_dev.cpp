// mg_dev.cpp -- mg_preconditioner on device vectors (multigrid/aa_mg/mg_complex.cpp:514-822).
//
// The control flow, the parameters handed to the smoother and the inner solvers, the printed lines and
// the dslash bookkeeping follow the reference statement by statement; vectors are device arrays and
// every operation is a C-ABI call (operator applies, BLAS-1, glb_mg_prolong / glb_mg_restrict).
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "dev_internal.hpp"
#include "mg_complex.h"

using namespace glbx;

dslash_tracker::dslash_tracker(int refine) : n_refine(refine) {
  krylov = new int[refine + 1];
  presmooth = new int[refine + 1];
  postsmooth = new int[refine + 1];
  residual = new int[refine + 1];
  nullvectors = new int[refine + 1];
  for (int i = 0; i < refine + 1; i++) krylov[i] = presmooth[i] = postsmooth[i] = residual[i] = nullvectors[i] = 0;
}
dslash_tracker::~dslash_tracker() {
  delete[] krylov;
  delete[] presmooth;
  delete[] postsmooth;
  delete[] residual;
  delete[] nullvectors;
}

// mg_complex.cpp:469-510 : only the level index is state here (sizes come from the operators)
void level_down(mg_operator_struct_complex_dev* mg) {
  if (mg->curr_level < mg->n_refine - 1) mg->curr_level++;
}
void level_up(mg_operator_struct_complex_dev* mg) {
  if (mg->curr_level > 0) mg->curr_level--;
}

namespace {

// D^dag D of one level as a device callback (fine_ / coarse_square_staggered_normal, mg_complex.cpp:147-172: the
// operator into a temporary, its dagger out of it)
struct NormalOp {
  glb_operator* d;
  glb_operator* ddag;
  zcplx* tmp;
};
void normal_apply_dev(zcplx* lhs, zcplx* rhs, void* e) {
  NormalOp* n = (NormalOp*)e;
  GLBX(glb_op_apply(n->d, n->tmp, rhs));
  GLBX(glb_op_apply(n->ddag, lhs, n->tmp));
}
glb_operator* dagger_of_level(mg_operator_struct_complex_dev* mg, int level) {
  glb_operator* d = mg->dagger_stencils ? mg->dagger_stencils[level] : 0;
  if (!d) throw Error("mg_preconditioner_dev: the normal-equation variants need mgstruct->dagger_stencils on every level they touch");
  return d;
}

void cycle(zcplx* lhs, zcplx* rhs, int size, mg_precond_struct_complex_dev* pc, inversion_verbose_struct* verb) {
  mg_operator_struct_complex_dev* mg = pc->mgstruct;
  const int lvl = mg->curr_level;
  const bool say = !pc->quiet;
  const bool normal_mg = pc->normal_eqn_mg, normal_any = pc->normal_eqn_smooth || pc->normal_eqn_mg;
  if (say) std::cout << "[MG]: Entered mg_preconditioner.\n";
  glb_operator* fine = mg->stencils[lvl];
  glb_operator* coarse = mg->stencils[lvl + 1];
  glb_mg_transfer* tr = mg->transfers[lvl];
  const int fine_size = (int)glb_op_local_size(fine);
  const int coarse_length = (int)glb_op_local_size(coarse);
  if (fine_size != size) throw Error("mg_preconditioner_dev: vector size does not match the current level");
  if ((size_t)fine_size != glb_mg_fine_size(tr) || (size_t)coarse_length != glb_mg_coarse_size(tr))
    throw Error("mg_preconditioner_dev: transfer and operators disagree on the level sizes");
  void (*apply)(zcplx*, zcplx*, void*) = &glb200_apply_dev;
  glb_context* ctx = glb_op_context(fine);
  Blas<zcplx> B = {ctx, (size_t)fine_size};
  Blas<zcplx> Bc = {ctx, (size_t)coarse_length};
  Work<zcplx> W(B);
  Work<zcplx> Wc(Bc);
  inversion_info invif;
  // The operators of this level (mg_complex.cpp:531-545 and the driver's wiring, aa_mg_square_staggered_u1.cpp:552-577):
  //   A        fine_matrix_vector    D of the level, D^dag D with normal_eqn_mg
  //   Acoarse  coarse_matrix_vector  the same one level down
  //   S        smoothing operator    A, or D^dag D with normal_eqn_smooth (then on the right-hand side D^dag r)
  void (*A)(zcplx*, zcplx*, void*) = apply;
  void* A_extra = (void*)fine;
  void (*Acoarse)(zcplx*, zcplx*, void*) = apply;
  void* Acoarse_extra = (void*)coarse;
  void (*S)(zcplx*, zcplx*, void*) = apply;
  void* S_extra = (void*)fine;
  NormalOp Nf = {fine, 0, 0}, Nc = {coarse, 0, 0};
  glb_operator* fine_dagger = 0;
  zcplx* rhs_smooth = 0;  // D^dag r of the CGNR smoother
  if (normal_any) {
    fine_dagger = dagger_of_level(mg, lvl);
    Nf.ddag = fine_dagger;
    Nf.tmp = W.get();
    S = &normal_apply_dev;
    S_extra = (void*)&Nf;
    if (normal_mg) {
      A = S;
      A_extra = S_extra;
      Nc.ddag = dagger_of_level(mg, lvl + 1);
      Nc.tmp = Wc.get();
      Acoarse = &normal_apply_dev;
      Acoarse_extra = (void*)&Nc;
    } else {
      rhs_smooth = W.get();
    }
  }
  const int smooth_ops = normal_any ? 2 : 1, coarse_ops = normal_mg ? 2 : 1;  // dslash counting, :580, :646, :803

  // 1. pre-smooth: z1 ~ A^-1 rhs from a zero guess, r1 = rhs - A z1   (mg_complex.cpp:548-595)
  zcplx* z1 = W.get();
  zcplx* r1 = W.get();
  B.zero(z1);
  if (pc->n_pre_smooth[lvl] > 0 && pc->in_smooth_type != MINV_INVALID) {
    minv_inverter_params pre_solve;
    pre_solve.tol = 1e-20;
    pre_solve.max_iters = pc->n_pre_smooth[lvl];
    pre_solve.restart = false;
    pre_solve.restart_freq = -1;
    pre_solve.sor_omega = 1.0;
    pre_solve.minres_omega = 1.0;
    pre_solve.bicgstabl_l = pc->n_pre_smooth[lvl];
    zcplx* b_smooth = rhs;
    if (rhs_smooth) {  // :556-560
      GLBX(glb_op_apply(fine_dagger, rhs_smooth, rhs));
      b_smooth = rhs_smooth;
    }
    invif = minv_unpreconditioned_dev(z1, b_smooth, fine_size, pc->in_smooth_type, pre_solve, S, S_extra);
    if (say) {
      printf("[L%d Presmooth]: Iterations %d Res %.8e Err N Algorithm %s\n", lvl + 1, invif.iter, sqrt(invif.resSq),
             invif.name.c_str());
      fflush(stdout);
    }
    mg->dslash_count->presmooth[lvl] += smooth_ops * invif.ops_count;
    A(r1, z1, A_extra);
    mg->dslash_count->residual[lvl]++;
    B.sub(rhs, r1, r1);
  } else {
    B.copy(r1, rhs);
  }

  // 2. coarse correction z2 = P (P^dag A P)^-1 P^dag r1, lhs = z1 + z2   (mg_complex.cpp:598-758)
  const bool coarsest = (lvl + 1 == mg->n_refine);
  if (pc->in_solve_type != NONE || !coarsest) {
    zcplx* z2 = W.get();
    zcplx* rhs_coarse = Wc.get();
    zcplx* lhs_coarse = Wc.get();
    GLBX(glb_mg_restrict(tr, rhs_coarse, r1));
    Bc.zero(lhs_coarse);
    if (pc->in_solve_type != NONE && coarsest) {
      const double tol = pc->rel_res[lvl];
      switch (pc->in_solve_type) {
        case CG:
          invif = minv_vector_cg_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol, Acoarse, Acoarse_extra, verb);
          break;
        case GCR:
          invif = minv_vector_gcr_restart_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol, pc->n_restart, Acoarse,
                                              Acoarse_extra, verb);
          break;
        case BICGSTAB:
        case BICGSTAB_L:
          invif = minv_vector_bicgstab_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol, Acoarse, Acoarse_extra,
                                           verb);
          break;
        case CR:
          invif = minv_vector_cr_restart_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol, pc->n_restart, Acoarse,
                                             Acoarse_extra, verb);
          break;
        case MINRES:  // mg_complex.cpp:622-624
          invif = minv_vector_minres_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol, Acoarse, Acoarse_extra, verb);
          break;
        default:
          throw Error("mg_preconditioner_dev: unknown coarse solver");
      }
      if (say)
        printf("[L%d]: Iterations %d RelRes %.8e Err N Algorithm %s\n", lvl + 2, invif.iter,
               sqrt(invif.resSq) / sqrt(Bc.norm2sq(rhs_coarse)), invif.name.c_str());
      mg->dslash_count->krylov[lvl + 1] += coarse_ops * invif.ops_count;
    } else {
      // not on the coarsest level: the level below is preconditioned by its own cycle
      if (say) {
        printf(pc->in_solve_type != NONE ? "About to enter coarser solve.\n" : "[L%d]: About to enter coarser solve.\n",
               lvl + 1);
        fflush(stdout);
      }
      // the level counter is restored whatever happens below (a throw in the recursive solve would otherwise leave
      // every later preconditioner call on the wrong level)
      struct LevelGuard {
        mg_operator_struct_complex_dev* m;
        bool armed;
        explicit LevelGuard(mg_operator_struct_complex_dev* mm) : m(mm), armed(true) { level_down(m); }
        void release() {
          if (armed) level_up(m);
          armed = false;
        }
        ~LevelGuard() { release(); }
      } guard(mg);
      if (pc->in_solve_type == NONE || pc->mlevel_type == MLEVEL_SMOOTH || pc->in_solve_type == MINRES) {
        cycle(lhs_coarse, rhs_coarse, coarse_length, pc, verb);
      } else {
        // mg_complex.cpp:672-695: the level below solved by a flexible Krylov method preconditioned by its own cycle --
        // flexible CG, VPGCR (also standing in for CR) or preconditioned BiCGStab (also for BiCGStab-l)
        void (*self)(zcplx*, zcplx*, int, void*, inversion_verbose_struct*) = &mg_preconditioner_dev;
        const double tol_rec = pc->rel_res[mg->curr_level - 1];
        if (pc->in_solve_type == CG)
          invif = minv_vector_cg_flex_precond_restart_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol_rec,
                                                          pc->n_restart, Acoarse, Acoarse_extra, self, (void*)pc, verb);
        else if (pc->in_solve_type == GCR || pc->in_solve_type == CR)
          invif = minv_vector_gcr_var_precond_restart_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol_rec,
                                                          pc->n_restart, Acoarse, Acoarse_extra, self, (void*)pc, verb);
        else if (pc->in_solve_type == BICGSTAB || pc->in_solve_type == BICGSTAB_L)
          invif = minv_vector_bicgstab_precond_dev(lhs_coarse, rhs_coarse, coarse_length, pc->n_max, tol_rec, Acoarse,
                                                   Acoarse_extra, self, (void*)pc, verb);
        else
          throw Error("mg_preconditioner_dev: unknown inner solver of the recursive cycle");
        if (say)
          printf("[L%d]: Iterations %d RelRes %.8e Err N Algorithm %s\n", mg->curr_level + 1, invif.iter,
                 sqrt(invif.resSq) / sqrt(Bc.norm2sq(rhs_coarse)), invif.name.c_str());
        mg->dslash_count->krylov[mg->curr_level] += invif.ops_count;
      }
      guard.release();
      if (say) {
        printf(pc->in_solve_type != NONE ? "Exited coarser solve.\n" : "[L%d]: Exited coarse solve.\n", lvl + 1);
        fflush(stdout);
      }
    }
    GLBX(glb_mg_prolong(tr, z2, lhs_coarse));
    B.add(z1, z2, lhs);
  } else {
    B.copy(lhs, z1);  // no inner solver on the coarsest level: lhs = z1
  }

  // 3. post-smooth on the residual equation: lhs += z3, z3 ~ A^-1 (rhs - A lhs)   (mg_complex.cpp:760-806)
  if (pc->n_post_smooth[lvl] > 0 && pc->in_smooth_type != MINV_INVALID) {
    zcplx* r2 = W.get();
    zcplx* z3 = W.get();
    A(r2, lhs, A_extra);
    mg->dslash_count->residual[lvl]++;
    B.sub(rhs, r2, r2);
    zcplx* b_smooth = r2;
    if (rhs_smooth) {  // :779-783
      GLBX(glb_op_apply(fine_dagger, rhs_smooth, r2));
      b_smooth = rhs_smooth;
    }
    minv_inverter_params post_solve;
    post_solve.tol = 1e-20;
    post_solve.max_iters = pc->n_post_smooth[lvl];
    post_solve.restart = false;
    post_solve.restart_freq = -1;
    post_solve.sor_omega = 1.0;
    post_solve.minres_omega = 1.0;
    post_solve.bicgstabl_l = pc->n_post_smooth[lvl];
    B.zero(z3);
    invif = minv_unpreconditioned_dev(z3, b_smooth, fine_size, pc->in_smooth_type, post_solve, S, S_extra);
    if (say)
      printf("[L%d Postsmooth]: Iterations %d Res %.8e Err N Algorithm %s\n", lvl + 1, invif.iter, sqrt(invif.resSq),
             invif.name.c_str());
    mg->dslash_count->postsmooth[lvl] += smooth_ops * invif.ops_count;
    B.add(lhs, z3, lhs);
  }
  if (say) std::cout << "[MG]: Exited mg_preconditioner.\n";
}

}  // namespace

void mg_preconditioner_dev(zcplx* d_lhs, zcplx* d_rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  mg_precond_struct_complex_dev* pc = (mg_precond_struct_complex_dev*)extra_data;
  try {
    cycle(d_lhs, d_rhs, size, pc, verb);
  } catch (const std::exception& e) {
    // The preconditioner contract has no error channel (generic_gcr_var_precond.h:16), and lhs may be half written
    // or not written at all: an outer flexible solver that kept iterating on it would silently return garbage.  Same
    // policy as the operator callback (dropin.cpp direct_apply): report and stop.
    std::cerr << "[glb200] mg_preconditioner failed: " << e.what() << std::endl;
    std::abort();
  }
}
