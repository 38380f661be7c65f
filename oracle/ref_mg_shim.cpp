// ref_mg_shim.cpp -- C interface to the reference's OWN adaptive-multigrid code (TEST INFRASTRUCTURE ONLY).
//
// Compiled together with the unmodified reference sources (multigrid/aa_mg/mg_complex.cpp,
// generic_gcr_var_precond.cpp, ... -- see oracle/Makefile) into oracle/_ref/libref_oracle.so.  It builds
// the two-level (or deeper) structure the reference driver builds in
// multigrid/aa_mg/aa_mg_square_staggered_u1.cpp:351-497,607-651,897 from caller-supplied null vectors:
//   Lattice per level, fine stencil = get_square_staggered_u1_stencil (operators_stencil.cpp:14),
//   block_orthonormalize (mg_complex.cpp:259), coarse stencil = generate_coarse_from_fine_stencil
//   (mg_complex.cpp:827), and then exposes the reference's prolong / restrict / mg_preconditioner and
//   minv_vector_gcr_var_precond_restart so that the GPU path can be checked against them on the same
//   inputs.  Nothing here is used by the product.
#include <complex>
#include <cstring>
#include <iostream>
#include <vector>
using namespace std;

#include "generic_vector.h"
#include "verbosity.h"
#include "generic_inverters.h"
#include "generic_inverters_precond.h"
#include "mg.h"
#include "mg_complex.h"
#include "lattice.h"
#include "operators.h"
#include "operators_stencil.h"
#include "coarse_stencil.h"
#include "null_gen.h"
#include <random>
#include <malloc.h>

namespace {

typedef complex<double> zc;

struct RefMg {
  staggered_u1_op stagif;
  vector<zc> links;
  mg_operator_struct_complex mg;
  mg_precond_struct_complex pre;
  vector<int> n_pre, n_post;
  vector<double> rel_res;
  int n_refine;
};

}  // namespace

extern "C" {

// links: reference layout, 2*X*Y complex.  block[l], nvec[l] for l < n_refine.  null[l][v]: host arrays of the
// level-l lattice size (level 0: X*Y; level l: Vol_l * nvec[l-1]).  Levels below the first are set up after the
// coarse stencils above them exist, exactly in the driver's order.
static RefMg* make_struct(int X, int Y, const double* links, double mass, int n_refine, const int* block, const int* nvec,
                          const double* const* const* null_in) {
  RefMg* h = new RefMg();
  h->n_refine = n_refine;
  h->links.assign((const zc*)links, (const zc*)links + 2 * (size_t)X * Y);
  h->stagif.lattice = h->links.data();
  h->stagif.mass = mass;
  h->stagif.x_fine = X;
  h->stagif.y_fine = Y;
  h->stagif.Nc = 1;
  h->stagif.wilson_coeff = 0.0;

  mg_operator_struct_complex& mg = h->mg;
  mg.x_fine = X;
  mg.y_fine = Y;
  mg.Nc = 1;
  mg.n_refine = n_refine;
  mg.dslash_count = new dslash_tracker(n_refine);
  mg.blocksize_x = new int[n_refine];
  mg.blocksize_y = new int[n_refine];
  mg.n_vectors = new int[n_refine];
  for (int i = 0; i < n_refine; i++) {
    mg.blocksize_x[i] = block[i];
    mg.blocksize_y[i] = block[i];
    mg.n_vectors[i] = nvec[i];
  }
  mg.matrix_vector = square_staggered_u1;
  mg.matrix_vector_dagger = square_staggered_dagger_u1;
  mg.matrix_extra_data = (void*)&h->stagif;
  mg.latt = new Lattice*[n_refine + 1];
  int dims[2] = {X, Y};
  mg.latt[0] = new Lattice(2, dims, 1);
  for (int i = 1; i <= n_refine; i++) {
    dims[0] = mg.latt[i - 1]->get_lattice_dimension(0) / mg.blocksize_x[i - 1];
    dims[1] = mg.latt[i - 1]->get_lattice_dimension(1) / mg.blocksize_y[i - 1];
    mg.latt[i] = new Lattice(2, dims, mg.n_vectors[i - 1]);
  }
  mg.curr_level = 0;
  mg.curr_dof_fine = mg.latt[0]->get_nc();
  mg.curr_x_fine = mg.latt[0]->get_lattice_dimension(0);
  mg.curr_y_fine = mg.latt[0]->get_lattice_dimension(1);
  mg.curr_fine_size = mg.latt[0]->get_lattice_size();
  mg.curr_dof_coarse = mg.latt[1]->get_nc();
  mg.curr_x_coarse = mg.latt[1]->get_lattice_dimension(0);
  mg.curr_y_coarse = mg.latt[1]->get_lattice_dimension(1);
  mg.curr_coarse_size = mg.latt[1]->get_lattice_size();

  mg.null_vectors = new zc**[n_refine];
  for (int i = 0; i < n_refine; i++) {
    mg.null_vectors[i] = new zc*[mg.n_vectors[i]];
    const int sz = mg.latt[i]->get_lattice_size();
    for (int j = 0; j < mg.n_vectors[i]; j++) {
      mg.null_vectors[i][j] = new zc[sz];
      if (null_in)
        memcpy(mg.null_vectors[i][j], null_in[i][j], sizeof(zc) * sz);
      else
        zero<double>(mg.null_vectors[i][j], sz);  // aa_mg_square_staggered_u1.cpp:613-614
    }
  }
  mg.stencils = new stencil_2d*[n_refine + 1];
  for (int i = 0; i <= n_refine; i++) mg.stencils[i] = new stencil_2d(mg.latt[i], 1);
  mg.have_dagger_stencil = false;
  mg.dagger_stencils = 0;

  mg_precond_struct_complex& p = h->pre;
  p.in_smooth_type = MINV_GCR;
  p.omega_smooth = 0.67;
  p.mlevel_type = MLEVEL_SMOOTH;
  p.in_solve_type = GCR;
  p.n_max = 1024;
  p.n_restart = 64;
  p.mgstruct = &mg;
  p.matrix_extra_data = (void*)&mg;
  h->n_pre.assign(n_refine, 6);
  h->n_post.assign(n_refine, 6);
  h->rel_res.assign(n_refine, 1e-2);
  p.n_pre_smooth = h->n_pre.data();
  p.n_post_smooth = h->n_post.data();
  p.rel_res = h->rel_res.data();
  p.normal_eqn_smooth = false;
  p.normal_eqn_mg = false;
  p.coarse_matrix_vector = coarse_square_staggered;
  p.fine_matrix_vector = fine_square_staggered;
  p.coarse_matrix_vector_dagger = coarse_square_staggered_dagger;
  p.fine_matrix_vector_dagger = fine_square_staggered_dagger;
  p.coarse_matrix_vector_normal = coarse_square_staggered_normal;
  p.fine_matrix_vector_normal = fine_square_staggered_normal;
  return h;
}

// links: reference layout, 2*X*Y complex.  block[l], nvec[l] for l < n_refine.  null[l][v]: host arrays of the
// level-l lattice size (level 0: X*Y; level l: Vol_l * nvec[l-1]).  ignore_shifts: the flag handed to
// generate_coarse_from_fine_stencil; when set the coarse shift is copied from the fine one as the driver does
// (aa_mg_square_staggered_u1.cpp:1080-1086).
void* refmg_create2(int X, int Y, const double* links, double mass, int n_refine, const int* block, const int* nvec,
                    const double* const* const* null_in, int ignore_shifts) {
  RefMg* h = make_struct(X, Y, links, mass, n_refine, block, nvec, null_in);
  mg_operator_struct_complex& mg = h->mg;
  if (ignore_shifts) {  // the driver's way: mass in the shift, not in the stencil (:986-996)
    h->stagif.mass = 0.0;
    get_square_staggered_u1_stencil(mg.stencils[0], &h->stagif);
    h->stagif.mass = mass;
    mg.stencils[0]->shift = mass;
  } else {
    get_square_staggered_u1_stencil(mg.stencils[0], &h->stagif);
  }
  // level by level: orthonormalise this level's null vectors, build the next stencil, step down
  for (int n = 0; n < n_refine; n++) {
    block_orthonormalize(&mg);
    generate_coarse_from_fine_stencil(mg.stencils[n + 1], mg.stencils[n], &mg, ignore_shifts != 0);
    if (ignore_shifts) mg.stencils[n + 1]->shift = mg.stencils[n]->shift;
    if (n != n_refine - 1) level_down(&mg);
  }
  for (int n = n_refine - 1; n > 0; n--) level_up(&mg);
  return h;
}
void* refmg_create(int X, int Y, const double* links, double mass, int n_refine, const int* block, const int* nvec,
                   const double* const* const* null_in) {
  return refmg_create2(X, Y, links, mass, n_refine, block, nvec, null_in, 0);
}

// The complete set-up of the reference's driver for --operator staggered --null-operator staggered
// (aa_mg_square_staggered_u1.cpp:716-1143): null vectors from null_generate_random_smooth (null_gen.cpp:193) on the
// null-generation stencils (mass = null_mass in the shift), block_orthonormalize, coarse null-generation stencil with
// ignore_shifts = true and the shift copied down, level_down; afterwards the final stencils with the true mass.
//   nvec[l]       total null vectors of refinement l (= n_null[l] * partitions)
//   bstrat        blocking_strategy (null_gen.h:15-21): 0 none (1 partition), 1 even/odd (2 partitions)
//   null_gen      minv_inverter of the smoothing solve; tol[l], max_iter[l] per refinement
//   restart_freq  > 0: restarted solver;  bicgstab_l: l of BiCGStab-l
//   seed          std::mt19937 seed of the gaussian sources
//   do_free       free-field null vectors (null_generate_free, null_gen.cpp:162) instead of smoothed random ones
//   null_prec     null_precond_strategy (null_gen.h:24-29): 0 none, 1 even/odd (top/bottom below the top level), 2 normal
void* refmg_setup(int X, int Y, const double* links, double mass, int n_refine, const int* block, const int* nvec,
                  int bstrat, double null_mass, int null_gen, const double* tol, const int* max_iter, int restart_freq,
                  int bicgstab_l, int do_ortho_eo, int do_global_ortho_conj, unsigned seed, int verbosity, int null_prec,
                  int do_free) {
  RefMg* h = make_struct(X, Y, links, mass, n_refine, block, nvec, 0);
  mg_operator_struct_complex& mg = h->mg;
  null_vector_params nv;
  nv.opt_null = STAGGERED;
  nv.null_gen = (minv_inverter)null_gen;
  nv.null_prec = (null_precond_strategy)null_prec;
  nv.null_restart = restart_freq > 0;
  nv.null_restart_freq = restart_freq;
  nv.null_bicgstab_l = bicgstab_l;
  nv.null_mass = null_mass;
  nv.bstrat = (blocking_strategy)bstrat;
  nv.null_partitions = (bstrat == BLOCK_EO || bstrat == BLOCK_TOPO) ? 2 : (bstrat == BLOCK_CORNER) ? 4 : 1;  // aa_mg_square_staggered_u1.cpp:412-427
  nv.do_global_ortho_conj = do_global_ortho_conj != 0;
  nv.do_ortho_eo = do_ortho_eo != 0;
  for (int i = 0; i < n_refine; i++) {
    nv.n_null_vectors.push_back(nvec[i] / nv.null_partitions);
    nv.null_precisions.push_back(tol[i]);
    nv.null_max_iters.push_back(max_iter[i]);
  }
  std::mt19937 generator(seed);
  inversion_verbose_struct verb;
  verb.verbosity = (inversion_verbose_level)verbosity;
  verb.verb_prefix = "";
  verb.precond_verbosity = VERB_NONE;
  verb.precond_verb_prefix = "";
  // null-generation stencils (:724-757): mass encoded in the shift
  h->stagif.mass = 0.0;
  get_square_staggered_u1_stencil(mg.stencils[0], &h->stagif);
  h->stagif.mass = mass;
  mg.stencils[0]->shift = null_mass;
  // null_gen.cpp:214,266: with NULL_PRECOND_EO the reference hands a new[]-ed, never initialised array to the solver
  // as its initial guess.  Make that read defined for the comparison: glibc fills fresh allocations with the
  // complement of M_PERTURB's low byte, so 0xFF gives the zero-filled memory a fresh mmap would have given.
  if (null_prec == NULL_PRECOND_EO) mallopt(M_PERTURB, 0xFF);
  for (int n = 0; n < n_refine; n++) {
    verb.verb_prefix = "[L" + to_string(mg.curr_level + 1) + "_NULLVEC]: ";
    if (do_free)  // aa_mg_square_staggered_u1.cpp:792-795
      null_generate_free(&mg, &nv, false, 0);
    else
      null_generate_random_smooth(&mg, &nv, &verb, &generator);
    block_orthonormalize(&mg);
    if (n != n_refine - 1) {
      generate_coarse_from_fine_stencil(mg.stencils[mg.curr_level + 1], mg.stencils[mg.curr_level], &mg, true);
      mg.stencils[mg.curr_level + 1]->shift = mg.stencils[mg.curr_level]->shift;
      level_down(&mg);
    }
  }
  if (null_prec == NULL_PRECOND_EO) mallopt(M_PERTURB, 0);
  for (int n = 1; n < n_refine; n++) level_up(&mg);
  // final stencils with the true mass (:982-1143)
  for (int n = 0; n <= n_refine; n++) mg.stencils[n]->clear_stencils();
  h->stagif.mass = 0.0;
  get_square_staggered_u1_stencil(mg.stencils[0], &h->stagif);
  h->stagif.mass = mass;
  mg.stencils[0]->shift = mass;
  for (int n = 0; n < n_refine; n++) {
    generate_coarse_from_fine_stencil(mg.stencils[n + 1], mg.stencils[n], &mg, true);
    mg.stencils[n + 1]->shift = mg.stencils[n]->shift;
    if (n != n_refine - 1) level_down(&mg);
  }
  for (int n = 1; n < n_refine; n++) level_up(&mg);
  return h;
}

// dslash_tracker::nullvectors per level (mg_complex.h:104-136)
void refmg_null_counts(void* hv, int* out) {
  RefMg* h = (RefMg*)hv;
  for (int i = 0; i <= h->n_refine; i++) out[i] = h->mg.dslash_count->nullvectors[i];
}

void refmg_free(void* hv) {
  // test infrastructure: the reference's structs own raw arrays without destructors; leak them
  (void)hv;
}

// lattice of level l: X, Y, dofs per site
void refmg_level_dims(void* hv, int level, int* X, int* Y, int* nc) {
  RefMg* h = (RefMg*)hv;
  *X = h->mg.latt[level]->get_lattice_dimension(0);
  *Y = h->mg.latt[level]->get_lattice_dimension(1);
  *nc = h->mg.latt[level]->get_nc();
}

// block-orthonormalised null vector v of level l
void refmg_get_null(void* hv, int level, int v, double* out) {
  RefMg* h = (RefMg*)hv;
  memcpy(out, h->mg.null_vectors[level][v], sizeof(zc) * h->mg.latt[level]->get_lattice_size());
}

// stencil of level l: clover nc*nc*V, hopping 4*nc*nc*V, shifts = {shift, eo_shift, dof_shift} (3 complex)
void refmg_get_stencil(void* hv, int level, double* clover, double* hopping, double* shifts) {
  RefMg* h = (RefMg*)hv;
  stencil_2d* s = h->mg.stencils[level];
  const size_t m = (size_t)s->lat->get_volume() * s->lat->get_nc() * s->lat->get_nc();
  memcpy(clover, s->clover, sizeof(zc) * m);
  memcpy(hopping, s->hopping, sizeof(zc) * 4 * m);
  const zc sh[3] = {s->shift, s->eo_shift, s->dof_shift};
  memcpy(shifts, sh, sizeof(sh));
}

static void goto_level(RefMg* h, int level) {
  while (h->mg.curr_level < level) level_down(&h->mg);
  while (h->mg.curr_level > level) level_up(&h->mg);
}

// transfers between level l (fine) and l+1 (coarse)
void refmg_prolong(void* hv, int level, double* fine, const double* coarse) {
  RefMg* h = (RefMg*)hv;
  goto_level(h, level);
  prolong((zc*)fine, (zc*)coarse, &h->mg);
  goto_level(h, 0);
}
void refmg_restrict(void* hv, int level, double* coarse, const double* fine) {
  RefMg* h = (RefMg*)hv;
  goto_level(h, level);
  restrict((zc*)coarse, (zc*)fine, &h->mg);
  goto_level(h, 0);
}

// the level-l operator as the preconditioner sees it (fine_square_staggered at curr_level = l)
void refmg_apply_level(void* hv, int level, double* lhs, const double* rhs) {
  RefMg* h = (RefMg*)hv;
  goto_level(h, level < h->n_refine ? level : h->n_refine - 1);
  if (level < h->n_refine)
    fine_square_staggered((zc*)lhs, (zc*)rhs, (void*)&h->mg);
  else
    coarse_square_staggered((zc*)lhs, (zc*)rhs, (void*)&h->mg);
  goto_level(h, 0);
}

// the same for the daggered (which = 1) and the normal (which = 2) operator of level l (mg_complex.cpp:93-172)
void refmg_apply_level_variant(void* hv, int level, int which, double* lhs, const double* rhs) {
  RefMg* h = (RefMg*)hv;
  goto_level(h, level < h->n_refine ? level : h->n_refine - 1);
  void* e = (void*)&h->mg;
  if (level < h->n_refine) {
    if (which == 1) fine_square_staggered_dagger((zc*)lhs, (zc*)rhs, e);
    else if (which == 2) fine_square_staggered_normal((zc*)lhs, (zc*)rhs, e);
    else fine_square_staggered((zc*)lhs, (zc*)rhs, e);
  } else {
    if (which == 1) coarse_square_staggered_dagger((zc*)lhs, (zc*)rhs, e);
    else if (which == 2) coarse_square_staggered_normal((zc*)lhs, (zc*)rhs, e);
    else coarse_square_staggered((zc*)lhs, (zc*)rhs, e);
  }
  goto_level(h, 0);
}

// mg_precond_struct_complex settings (mg_complex.h:185-236); smoother / inner solver enums are the reference's
void refmg_set_precond(void* hv, int in_smooth_type, int n_pre, int n_post, int in_solve_type, int n_max, int n_restart,
                       double rel_res, int mlevel_type) {
  RefMg* h = (RefMg*)hv;
  h->pre.in_smooth_type = (minv_inverter)in_smooth_type;
  h->pre.in_solve_type = (inner_solver)in_solve_type;
  h->pre.n_max = n_max;
  h->pre.n_restart = n_restart;
  h->pre.mlevel_type = (mg_multilevel_type)mlevel_type;
  for (int i = 0; i < h->n_refine; i++) {
    h->n_pre[i] = n_pre;
    h->n_post[i] = n_post;
    h->rel_res[i] = rel_res;
  }
}

// The normal-equation variants of the cycle, wired as the reference's driver does (aa_mg_square_staggered_u1.cpp:550-577,
// :656-681, :990-1116): dagger stencils on every level -- the daggered staggered stencil on top, its Galerkin products
// below, with the same treatment of the mass as the stencils of this handle (ignore_shifts of refmg_create2) -- and
//   normal_smooth  the smoother runs on D^dag D z = D^dag r (CGNR)
//   normal_mg      fine, coarse and smoothing operator are D^dag D of their level
//   ignore_shifts < 0: no dagger stencils at all (the levels below the top dagger by prolong / restrict, mg_complex.cpp:101-111)
void refmg_set_normal(void* hv, int normal_smooth, int normal_mg, int ignore_shifts) {
  RefMg* h = (RefMg*)hv;
  mg_operator_struct_complex& mg = h->mg;
  goto_level(h, 0);
  if (!mg.have_dagger_stencil && ignore_shifts >= 0) {
    mg.have_dagger_stencil = true;
    mg.dagger_stencils = new stencil_2d*[h->n_refine + 1];
    for (int i = 0; i <= h->n_refine; i++) mg.dagger_stencils[i] = new stencil_2d(mg.latt[i], 1);
    const double mass = h->stagif.mass;
    if (ignore_shifts) h->stagif.mass = 0.0;
    get_square_staggered_dagger_u1_stencil(mg.dagger_stencils[0], &h->stagif);
    h->stagif.mass = mass;
    if (ignore_shifts) mg.dagger_stencils[0]->shift = mass;
    for (int n = 0; n < h->n_refine; n++) {
      generate_coarse_from_fine_stencil(mg.dagger_stencils[n + 1], mg.dagger_stencils[n], &mg, ignore_shifts != 0);
      if (ignore_shifts) mg.dagger_stencils[n + 1]->shift = mg.dagger_stencils[n]->shift;
      if (n != h->n_refine - 1) level_down(&mg);
    }
    goto_level(h, 0);
  }
  mg_precond_struct_complex& p = h->pre;
  p.normal_eqn_smooth = normal_smooth != 0;
  p.normal_eqn_mg = normal_mg != 0;
  if (normal_mg) {
    p.coarse_matrix_vector = coarse_square_staggered_normal;
    p.fine_matrix_vector = fine_square_staggered_normal;
    p.coarse_matrix_vector_dagger = coarse_square_staggered_normal;
    p.fine_matrix_vector_dagger = fine_square_staggered_normal;
  } else {
    p.coarse_matrix_vector = coarse_square_staggered;
    p.fine_matrix_vector = fine_square_staggered;
    p.coarse_matrix_vector_dagger = coarse_square_staggered_dagger;
    p.fine_matrix_vector_dagger = fine_square_staggered_dagger;
  }
  p.coarse_matrix_vector_normal = coarse_square_staggered_normal;
  p.fine_matrix_vector_normal = fine_square_staggered_normal;
}
// dslash_tracker of the handle (mg_complex.h:104-136): out[4][n_refine+1] = krylov, presmooth, postsmooth, residual
void refmg_counts(void* hv, int* out) {
  RefMg* h = (RefMg*)hv;
  const int n = h->n_refine + 1;
  for (int i = 0; i < n; i++) {
    out[i] = h->mg.dslash_count->krylov[i];
    out[n + i] = h->mg.dslash_count->presmooth[i];
    out[2 * n + i] = h->mg.dslash_count->postsmooth[i];
    out[3 * n + i] = h->mg.dslash_count->residual[i];
  }
}

// one application of mg_preconditioner (mg_complex.cpp:514) on the top level: lhs = M^-1 rhs
void refmg_vcycle(void* hv, double* lhs, const double* rhs) {
  RefMg* h = (RefMg*)hv;
  goto_level(h, 0);
  mg_preconditioner((zc*)lhs, (zc*)rhs, h->mg.curr_fine_size, (void*)&h->pre, 0);
}

// minv_vector_gcr_var_precond_restart (generic_gcr_var_precond.cpp:368) on the level-0 stencil operator with
// mg_preconditioner; restart_freq <= 0 selects the unrestarted minv_vector_gcr_var_precond.
// out: {resSq, iter, success, ops_count}
void refmg_vpgcr(void* hv, double* phi, const double* phi0, int max_iter, double res, int restart_freq, int verbosity,
                 double* out) {
  RefMg* h = (RefMg*)hv;
  goto_level(h, 0);
  inversion_verbose_struct verb;
  verb.verbosity = (inversion_verbose_level)verbosity;
  verb.verb_prefix = "[L1]: ";
  verb.precond_verbosity = VERB_NONE;
  verb.precond_verb_prefix = "";
  const int n = h->mg.curr_fine_size;
  inversion_info inf;
  // aa_mg_square_staggered_u1.cpp:1696-1705: with normal_eqn_mg the outer operator is D^dag D as well
  void (*fine_op)(zc*, zc*, void*) = h->pre.normal_eqn_mg ? fine_square_staggered_normal : fine_square_staggered;
  if (restart_freq > 0)
    inf = minv_vector_gcr_var_precond_restart((zc*)phi, (zc*)phi0, n, max_iter, res, restart_freq, fine_op,
                                              (void*)&h->mg, mg_preconditioner, (void*)&h->pre, &verb);
  else
    inf = minv_vector_gcr_var_precond((zc*)phi, (zc*)phi0, n, max_iter, res, fine_op, (void*)&h->mg,
                                      mg_preconditioner, (void*)&h->pre, &verb);
  out[0] = inf.resSq;
  out[1] = inf.iter;
  out[2] = inf.success ? 1.0 : 0.0;
  out[3] = inf.ops_count;
}

}  // extern "C"
