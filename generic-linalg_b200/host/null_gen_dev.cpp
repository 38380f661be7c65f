// null_gen_dev.cpp -- multigrid set-up on device vectors: null_partition_*, null_generate_random_smooth
// (multigrid/aa_mg/null_gen.cpp:13-400), block_orthonormalize (mg_complex.cpp:259-370) and
// generate_coarse_from_fine_stencil (mg_complex.cpp:827-1026).
//
// The statements follow the reference one by one; vectors are device arrays and every operation is a C-ABI call:
// operator applies, BLAS-1 (glb_dot / glb_norm2sq / glb_axpy / glb_rscale / glb_conj), glb_mg_partition,
// glb_mg_block_orthonormalize, glb_mg_transfer_create_dev and glb_mg_galerkin.  The one thing that stays on the
// host is the random source: it is drawn from the caller's std::mt19937 exactly as generic_vector.h:48-60 draws it
// and uploaded, so the reference and this code start from the same numbers.
#include <cmath>
#include <iostream>

#include "dev_internal.hpp"
#include "null_gen.h"

using namespace glbx;

void mg_level_dims(const mg_operator_struct_complex_dev* mg, int level, int* X, int* Y, int* dof) {
  int x = mg->x_fine, y = mg->y_fine, d = 1;
  for (int l = 0; l < level; l++) {
    x /= mg->blocksize_x[l];
    y /= mg->blocksize_y[l];
    d = mg->n_vectors[l];
  }
  *X = x;
  *Y = y;
  *dof = d;
}

namespace {

struct Level {
  glb_context* ctx;
  int X, Y, dof;   // the GLOBAL lattice of the level
  int y0, Yloc;    // this rank's rows (y-slabs; the whole lattice on one rank)
  int size;        // local vector length X*Yloc*dof
  int global_size;
};

Level level_of(const mg_operator_struct_complex_dev* mg) {
  if (!mg || !mg->stencils || !mg->blocksize_x || !mg->blocksize_y || !mg->n_vectors || !mg->null_vectors)
    throw Error("multigrid set-up: the set-up fields of mg_operator_struct_complex_dev are not filled in");
  glb_operator* op = mg->stencils[mg->curr_level];
  if (!op) throw Error("multigrid set-up: no operator on the current level");
  Level L;
  L.ctx = glb_op_context(op);
  mg_level_dims(mg, mg->curr_level, &L.X, &L.Y, &L.dof);
  GLBX(glb_slab_bounds(L.ctx, L.Y, &L.y0, &L.Yloc));
  L.size = L.X * L.Yloc * L.dof;
  L.global_size = L.X * L.Y * L.dof;
  if ((size_t)L.size != glb_op_local_size(op)) throw Error("multigrid set-up: operator and level sizes disagree");
  return L;
}

// generic_vector.h:173-203 normalize: res = 1/sqrt(|v|^2); if (res > 0) v *= res
void normalize_dev(const Blas<zcplx>& B, zcplx* v) {
  const double res = 1.0 / sqrt(B.norm2sq(v));
  if (res > 0.0) GLBX(glb_rscale(B.ctx, GLB_COMPLEX, B.n, v, res, v));
}

// generic_vector.h:236-249 orthogonal: v1 += (-<v2,v1>/|v2|^2) v2
void orthogonal_dev(const Blas<zcplx>& B, zcplx* v1, const zcplx* v2) {
  const zcplx alpha = -B.dot(v2, v1) / B.norm2sq(v2);
  B.axpy(alpha, v2, v1);
}

void conj_dev(const Blas<zcplx>& B, zcplx* v) { GLBX(glb_conj(B.ctx, GLB_COMPLEX, B.n, v, v)); }

// generic_vector.h:48-60: g++ evaluates the two constructor arguments right to left, so the reference draws the
// imaginary part first (pinned against the reference build in tests/test_oracle_cpu.py)
void gaussian_host(std::vector<zcplx>& v, std::mt19937& generator) {
  std::normal_distribution<> dist(0.0, 1.0);
  for (size_t i = 0; i < v.size(); i++) {
    const double im = dist(generator);
    const double re = dist(generator);
    v[i] = zcplx(re, im);
  }
}

// a composite view of a stencil2d operator for the duration of one solve
struct StencilView {
  glb_operator* view;
  StencilView(glb_operator* base, int kind) : view(0) { GLBX(glb_op_create_stencil_view(base, kind, 0, &view)); }
  ~StencilView() {
    if (view) glb_op_destroy(view);
  }
};

void partition(mg_operator_struct_complex_dev* mg, int num_null_vec, blocking_strategy bstrat, bool by_colour) {
  const Level L = level_of(mg);
  const int lvl = mg->curr_level;
  // null_gen.cpp:114: below the top level the colour index is taken modulo n_vectors[curr_level] (the number of
  // vectors being built on this level, not the dofs per site of this level -- kept as the reference has it)
  const int period = by_colour ? mg->n_vectors[lvl] : 0;
  zcplx** null = mg->null_vectors[lvl];
  switch (bstrat) {
    case BLOCK_NONE:
      return;
    case BLOCK_EO:
      GLBX(glb_mg_partition(L.ctx, L.X, L.Y, L.dof, period, null[num_null_vec], null[num_null_vec + mg->n_vectors[lvl] / 2]));
      return;
    case BLOCK_CORNER:  // null_gen.cpp:74-88, :132-152: class k goes to vector num_null_vec + k*n_vectors/4
      for (int k = 1; k < 4; k++)
        GLBX(glb_mg_partition_corner(L.ctx, L.X, L.Y, L.dof, period, k, null[num_null_vec],
                                     null[num_null_vec + k * mg->n_vectors[lvl] / 4]));
      return;
    case BLOCK_TOPO: {
      // null_gen.cpp:36-71 (top level; below it null_partition_coarse treats BLOCK_TOPO like BLOCK_EO): with
      // Gamma_5 = (i/2)(S_x S_y - S_y S_x) of the symmetric shifts, v -> (1 +- Gamma_5)/2 v
      if (by_colour) throw Error("null_partition: BLOCK_TOPO below the top level is the BLOCK_EO split");
      if (!mg->symmshift_x || !mg->symmshift_y)
        throw Error("null_partition: BLOCK_TOPO needs mgstruct->symmshift_x / symmshift_y (the symmetric shifts of the gauge field)");
      zcplx* v = null[num_null_vec];
      zcplx* g5v = null[num_null_vec + mg->n_vectors[lvl] / 2];
      Blas<zcplx> B = {L.ctx, (size_t)L.size};
      Work<zcplx> W(B);
      zcplx *tmp = W.get(), *tmp2 = W.get();
      GLBX(glb_op_apply(mg->symmshift_y, tmp, v));
      GLBX(glb_op_apply(mg->symmshift_x, tmp2, tmp));
      B.axpy(zcplx(0.0, 0.5), tmp2, g5v);    // += (i/2) S_x S_y v
      GLBX(glb_op_apply(mg->symmshift_x, tmp, v));
      GLBX(glb_op_apply(mg->symmshift_y, tmp2, tmp));
      B.axpy(-zcplx(0.0, 0.5), tmp2, g5v);   // -= (i/2) S_y S_x v
      B.add(v, g5v, v);                      // v <- 0.5 (v + Gamma_5 v)
      GLBX(glb_rscale(L.ctx, GLB_COMPLEX, B.n, v, 0.5, v));
      B.sub(v, g5v, g5v);                    // Gamma_5 slot <- v_new - Gamma_5 v
      return;
    }
    default:
      throw Error("null_partition: unknown blocking strategy");
  }
}

}  // namespace

void null_partition_staggered_dev(mg_operator_struct_complex_dev* mg, int num_null_vec, blocking_strategy bstrat) {
  partition(mg, num_null_vec, bstrat, false);
}

void null_partition_coarse_dev(mg_operator_struct_complex_dev* mg, int num_null_vec, blocking_strategy bstrat) {
  // null_gen.cpp:112: BLOCK_TOPO partitions the coarse levels like BLOCK_EO
  partition(mg, num_null_vec, bstrat == BLOCK_TOPO ? BLOCK_EO : bstrat, true);
}

// null_gen.cpp:162-191: one constant vector (times a gauge transformation on the top level), partitioned and
// normalised part by part.  gauge_trans is a HOST array of the top-level lattice size.
void null_generate_free_dev(mg_operator_struct_complex_dev* mg, null_vector_params* nv, bool do_gauge_transform,
                            std::complex<double>* gauge_trans) {
  const Level L = level_of(mg);
  const int lvl = mg->curr_level;
  if (nv->null_partitions < 1 || (int)nv->n_null_vectors.size() <= lvl)
    throw Error("null_generate_free_dev: null_partitions / n_null_vectors are not filled in");
  std::vector<zcplx> host(L.size);
  const size_t slab_off = (size_t)L.y0 * L.X * L.dof;  // gauge_trans is a global array
  for (int i = 0; i < L.size; i++) {
    host[i] = 1;
    if (do_gauge_transform && lvl == 0) host[i] *= gauge_trans[slab_off + i];
  }
  zcplx** null = mg->null_vectors[lvl];
  Blas<zcplx> B = {L.ctx, (size_t)L.size};
  GLBX(glb_vec_upload(L.ctx, GLB_COMPLEX, B.n, null[0], host.data()));
  if (lvl == 0)
    null_partition_staggered_dev(mg, 0, nv->bstrat);
  else
    null_partition_coarse_dev(mg, 0, nv->bstrat);
  for (int k = 0; k < nv->null_partitions; k++) normalize_dev(B, null[k * nv->n_null_vectors[lvl]]);
}

void null_generate_random_smooth_dev(mg_operator_struct_complex_dev* mg, null_vector_params* nv,
                                     inversion_verbose_struct* verb, std::mt19937* generator) {
  const Level L = level_of(mg);
  const int lvl = mg->curr_level;
  if (nv->null_partitions < 1 || mg->n_vectors[lvl] % nv->null_partitions != 0)
    throw Error("null_generate_random_smooth_dev: n_vectors must be a multiple of null_partitions");
  if ((int)nv->n_null_vectors.size() <= lvl || (int)nv->null_precisions.size() <= lvl || (int)nv->null_max_iters.size() <= lvl)
    throw Error("null_generate_random_smooth_dev: n_null_vectors / null_precisions / null_max_iters need one entry per level");
  if (nv->null_prec != NULL_PRECOND_NONE && nv->null_prec != NULL_PRECOND_EO && nv->null_prec != NULL_PRECOND_NORMAL)
    throw Error("null_generate_random_smooth_dev: unknown null_precond_strategy");
  glb_operator* op = mg->stencils[lvl];
  zcplx** null = mg->null_vectors[lvl];
  const int n_gen = mg->n_vectors[lvl] / nv->null_partitions;
  const int stride = nv->n_null_vectors[lvl];
  const int n_split = nv->do_ortho_eo ? nv->null_partitions : 1;
  void (*apply)(zcplx*, zcplx*, void*) = &glb200_apply_dev;

  // null_gen.cpp:199-206
  minv_inverter_params solve;
  solve.tol = nv->null_precisions[lvl];
  solve.max_iters = nv->null_max_iters[lvl];
  solve.restart = nv->null_restart;
  solve.restart_freq = nv->null_restart_freq;
  solve.minres_omega = nv->null_relaxation;
  solve.sor_omega = nv->null_relaxation;
  solve.bicgstabl_l = nv->null_bicgstab_l;

  Blas<zcplx> B = {L.ctx, (size_t)L.size};
  Work<zcplx> W(B);
  zcplx* rand_guess = W.get();
  zcplx* Arand_guess = W.get();
  zcplx* prep = (nv->null_prec != NULL_PRECOND_NONE) ? W.get() : 0;      // Arand_guess_prep (:213)
  zcplx* prec_soln = (nv->null_prec == NULL_PRECOND_EO) ? W.get() : 0;   // Arand_guess_prec_soln (:214)
  if (prec_soln) B.zero(prec_soln);  // the reference hands this new[]-ed array to the solver as its initial guess
  // the reference draws one global vector per source; every rank draws all of it and keeps its rows, so that the
  // random numbers do not depend on the number of ranks
  std::vector<zcplx> host(L.global_size);
  const size_t slab_off = (size_t)L.y0 * L.X * L.dof;

  for (int i = 0; i < n_gen; i++) {
    // a gaussian source (:219), orthogonal to the vectors found so far (:222-238)
    gaussian_host(host, *generator);
    GLBX(glb_vec_upload(L.ctx, GLB_COMPLEX, B.n, rand_guess, host.data() + slab_off));
    for (int j = 0; j < i; j++) {
      for (int k = 0; k < n_split; k++) {
        zcplx* prev = null[j + k * stride];
        orthogonal_dev(B, rand_guess, prev);
        if (nv->do_global_ortho_conj) {
          conj_dev(B, prev);
          orthogonal_dev(B, rand_guess, prev);
          conj_dev(B, prev);
        }
      }
    }

    // the residual equation A x = -A x0 (:241-250)
    B.zero(Arand_guess);
    GLBX(glb_op_apply(op, Arand_guess, rand_guess));
    mg->dslash_count->nullvectors[lvl]++;
    GLBX(glb_rscale(L.ctx, GLB_COMPLEX, B.n, Arand_guess, -1.0, Arand_guess));

    if (nv->null_prec == NULL_PRECOND_NONE) {
      // :252-258 (the initial guess is whatever null[i] holds: zero after allocation)
      inversion_info invif = minv_unpreconditioned_dev(null[i], Arand_guess, L.size, nv->null_gen, solve, apply, (void*)op, verb);
      mg->dslash_count->nullvectors[lvl] += invif.ops_count;
    } else if (nv->null_prec == NULL_PRECOND_EO) {
      // :259-289: even/odd preconditioned on the top level, top/bottom (colour halves) below it.  The prepare is
      // half an apply and is counted with the reconstruct.
      const int tb = (lvl == 0) ? 0 : 1;
      StencilView M(op, tb ? GLB_SV_M2MDTBDBT : GLB_SV_M2MDEODOE);
      GLBX(glb_stencil_prec_prepare(op, tb, prep, Arand_guess));
      inversion_info invif = minv_unpreconditioned_dev(prec_soln, prep, L.size, nv->null_gen, solve, apply, (void*)M.view, verb);
      mg->dslash_count->nullvectors[lvl] += invif.ops_count;
      GLBX(glb_stencil_prec_reconstruct(op, tb, null[i], prec_soln, Arand_guess));
      mg->dslash_count->nullvectors[lvl]++;
    } else {
      // :290-313 NULL_PRECOND_NORMAL: D^dag rhs through epsilon D epsilon (sigma_3 D sigma_3 below the top level),
      // then the non-Galerkin normal operator m^2 - D_eo D_oe - D_oe D_eo; every solver step counts two applies
      const int tb = (lvl == 0) ? 0 : 1;
      StencilView Dd(op, tb ? GLB_SV_DAGGER_TB : GLB_SV_DAGGER_EO);
      StencilView N(op, tb ? GLB_SV_NORMAL_TB : GLB_SV_NORMAL_EO);
      GLBX(glb_op_apply(Dd.view, prep, Arand_guess));
      mg->dslash_count->nullvectors[lvl]++;
      inversion_info invif = minv_unpreconditioned_dev(null[i], prep, L.size, nv->null_gen, solve, apply, (void*)N.view, verb);
      mg->dslash_count->nullvectors[lvl] += 2 * invif.ops_count;
    }

    // undo the residual equation (:316-319)
    B.add(null[i], rand_guess, null[i]);

    // split now if asked to (:322-335)
    if (nv->do_ortho_eo) {
      if (lvl == 0)
        null_partition_staggered_dev(mg, i, nv->bstrat);
      else
        null_partition_coarse_dev(mg, i, nv->bstrat);
    }
    for (int k = 0; k < n_split; k++) normalize_dev(B, null[i + k * stride]);  // :338-341

    // orthogonalise against the previous vectors (:344-366)
    for (int j = 0; j < i; j++) {
      if (!nv->quiet) std::cout << "[L" << lvl + 1 << "_NULLVEC]: Pre-orthog cosines of " << j << "," << i << " are: ";
      for (int k = 0; k < n_split; k++) {
        zcplx* vi = null[i + k * stride];
        zcplx* vj = null[j + k * stride];
        if (!nv->quiet) std::cout << std::abs(B.dot(vi, vj) / sqrt(B.norm2sq(vi) * B.norm2sq(vj))) << " ";
        orthogonal_dev(B, vi, vj);
        if (nv->do_global_ortho_conj) {
          conj_dev(B, vj);
          orthogonal_dev(B, vi, vj);
          conj_dev(B, vj);
        }
      }
      if (!nv->quiet) std::cout << "\n";
    }
    for (int k = 0; k < n_split; k++) normalize_dev(B, null[i + k * stride]);  // :369-372
  }

  // split afterwards otherwise (:376-397)
  if (!nv->do_ortho_eo) {
    for (int i = 0; i < n_gen; i++) {
      if (lvl == 0)
        null_partition_staggered_dev(mg, i, nv->bstrat);
      else
        null_partition_coarse_dev(mg, i, nv->bstrat);
      for (int k = 0; k < nv->null_partitions; k++) normalize_dev(B, null[i + k * stride]);
    }
  }
}

void block_orthonormalize_dev(mg_operator_struct_complex_dev* mg) {
  const Level L = level_of(mg);
  const int lvl = mg->curr_level;
  if (L.Yloc % mg->blocksize_y[lvl] != 0 || L.y0 % mg->blocksize_y[lvl] != 0)
    throw Error("block_orthonormalize_dev: slab boundaries must coincide with block boundaries");
  GLBX(glb_mg_block_orthonormalize(L.ctx, L.X, L.Yloc, L.dof, mg->blocksize_x[lvl], mg->blocksize_y[lvl], mg->n_vectors[lvl],
                                   (void* const*)mg->null_vectors[lvl]));
}

void generate_coarse_from_fine_stencil_dev(mg_operator_struct_complex_dev* mg, bool ignore_shifts) {
  const Level L = level_of(mg);
  const int lvl = mg->curr_level;
  if (!mg->transfers) throw Error("generate_coarse_from_fine_stencil_dev: no transfer table");
  if (mg->transfers[lvl]) {
    glb_mg_transfer_destroy(mg->transfers[lvl]);
    mg->transfers[lvl] = 0;
  }
  GLBX(glb_mg_transfer_create_dev(L.ctx, L.X, L.Yloc, L.dof, mg->blocksize_x[lvl], mg->blocksize_y[lvl], mg->n_vectors[lvl],
                                  (const void* const*)mg->null_vectors[lvl], &mg->transfers[lvl]));
  if (mg->stencils[lvl + 1]) {
    glb_op_destroy(mg->stencils[lvl + 1]);
    mg->stencils[lvl + 1] = 0;
  }
  GLBX(glb_mg_galerkin(mg->transfers[lvl], mg->stencils[lvl], ignore_shifts ? 1 : 0, &mg->stencils[lvl + 1]));
}
