// mg_host.cpp -- the reference's HOST-pointer multigrid interface (multigrid/aa_mg/mg_complex.h, null_gen.h) on top of the
// device routines: every function here moves its operands to the device, calls the _dev routine / C-ABI kernel that does
// the work, and moves the result back.  No arithmetic on vectors happens on the host.  Inside a drop-in solve the
// hierarchy is uploaded once (dropin.cpp recognises fine_square_staggered / mg_preconditioner as callbacks).
#include <iostream>
#include <vector>

#include "dev_internal.hpp"
#include "mg_complex.h"
#include "null_gen.h"
#include "operators.h"
#include "operators_stencil.h"

using namespace glbx;

// ---------------------------------------------------------------------------------------------- level bookkeeping
// mg_complex.cpp:469-510
static void refresh_level(mg_operator_struct_complex* mg) {
  Lattice* f = mg->latt[mg->curr_level];
  Lattice* c = mg->latt[mg->curr_level + 1];
  mg->curr_dof_fine = f->get_nc();
  mg->curr_x_fine = f->get_lattice_dimension(0);
  mg->curr_y_fine = f->get_lattice_dimension(1);
  mg->curr_fine_size = f->get_lattice_size();
  mg->curr_dof_coarse = c->get_nc();
  mg->curr_x_coarse = c->get_lattice_dimension(0);
  mg->curr_y_coarse = c->get_lattice_dimension(1);
  mg->curr_coarse_size = c->get_lattice_size();
}
void level_down(mg_operator_struct_complex* mg) {
  if (mg->curr_level < mg->n_refine - 1) {
    mg->curr_level++;
    refresh_level(mg);
  }
}
void level_up(mg_operator_struct_complex* mg) {
  if (mg->curr_level > 0) {
    mg->curr_level--;
    refresh_level(mg);
  }
}

namespace glb200_mg_host {

glb_operator* upload_stencil(glb_context* ctx, stencil_2d* st) {
  if (!st || !st->generated) throw Error("multigrid (host interface): every level needs a generated stencil");
  if (st->sdir != DIR_ALL) throw Error("multigrid (host interface): only sdir == DIR_ALL stencils");
  const double sh[2] = {st->shift.real(), st->shift.imag()}, eo[2] = {st->eo_shift.real(), st->eo_shift.imag()},
               df[2] = {st->dof_shift.real(), st->dof_shift.imag()};
  glb_operator* op = 0;
  GLBX(glb_op_create_stencil2d(ctx, st->clover, st->hopping, st->has_two ? st->two_link : 0, st->lat->get_lattice_dimension(0),
                               st->lat->get_lattice_dimension(1), st->lat->get_nc(), sh, eo, df, &op));
  return op;
}

// The adjoint of a generated stencil as a device operator: entry ((x,r),(x+d,c)) of D^dag is the conjugate of entry
// ((x+d,c),(x,r)) of D, i.e. of the opposite direction's matrix at the neighbour, transposed; the three shifts are
// diagonal and conjugate.  Stands in for a level's dagger stencil where the host struct has none: in exact arithmetic
// it is what the reference's fall-back computes (prolong -> dagger above -> restrict = P^dag D^dag P, mg_complex.cpp:101-111).
glb_operator* upload_adjoint_stencil(glb_context* ctx, stencil_2d* st) {
  if (!st || !st->generated) throw Error("multigrid (host interface): every level needs a generated stencil");
  const int X = st->lat->get_lattice_dimension(0), Y = st->lat->get_lattice_dimension(1), nc = st->lat->get_nc();
  const size_t V = (size_t)X * Y, plane = V * nc * nc;
  std::vector<zcplx> cl(plane), hp(4 * plane), tw(st->has_two ? 8 * plane : 0);
  static const int hop[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
  static const int two[8][2] = {{2, 0}, {1, 1}, {0, 2}, {-1, 1}, {-2, 0}, {-1, -1}, {0, -2}, {1, -1}};
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < X; x++) {
      const size_t site = (size_t)y * X + x;
      for (int r = 0; r < nc; r++)
        for (int c = 0; c < nc; c++) {
          cl[(site * nc + r) * nc + c] = conj(st->clover[(site * nc + c) * nc + r]);
          for (int d = 0; d < 4; d++) {
            const size_t nb = (size_t)((y + hop[d][1] + Y) % Y) * X + (size_t)((x + hop[d][0] + X) % X);
            hp[d * plane + (site * nc + r) * nc + c] = conj(st->hopping[((d + 2) % 4) * plane + (nb * nc + c) * nc + r]);
          }
          if (st->has_two)
            for (int d = 0; d < 8; d++) {
              const size_t nb = (size_t)((y + two[d][1] + 2 * Y) % Y) * X + (size_t)((x + two[d][0] + 2 * X) % X);
              tw[d * plane + (site * nc + r) * nc + c] = conj(st->two_link[((d + 4) % 8) * plane + (nb * nc + c) * nc + r]);
            }
        }
    }
  const double sh[2] = {st->shift.real(), -st->shift.imag()}, eo[2] = {st->eo_shift.real(), -st->eo_shift.imag()},
               df[2] = {st->dof_shift.real(), -st->dof_shift.imag()};
  glb_operator* op = 0;
  GLBX(glb_op_create_stencil2d(ctx, cl.data(), hp.data(), st->has_two ? tw.data() : 0, X, Y, nc, sh, eo, df, &op));
  return op;
}

glb_mg_transfer* upload_transfer(glb_context* ctx, mg_operator_struct_complex* mg, int level) {
  Lattice* f = mg->latt[level];
  glb_mg_transfer* t = 0;
  GLBX(glb_mg_transfer_create(ctx, f->get_lattice_dimension(0), f->get_lattice_dimension(1), f->get_nc(), mg->blocksize_x[level],
                              mg->blocksize_y[level], mg->n_vectors[level], (const void* const*)mg->null_vectors[level], &t));
  return t;
}

// the device image of a host hierarchy: operators of every level, transfers of every refinement
Hierarchy::Hierarchy(mg_operator_struct_complex* host, bool with_dagger) : ctx(glb200_default_context()) {
  const int n = host->n_refine;
  ops.assign(n + 1, (glb_operator*)0);
  trs.assign(n, (glb_mg_transfer*)0);
  try {
    for (int l = 0; l <= n; l++) ops[l] = upload_stencil(ctx, host->stencils[l]);
    for (int l = 0; l < n; l++) trs[l] = upload_transfer(ctx, host, l);
    if (with_dagger) {
      // D^dag per level for the normal-equation variants: on the top level the reference applies the FUNCTION
      // matrix_vector_dagger (mg_complex.cpp:119-123), below it the dagger stencil (:97-100) and without one
      // P^dag D^dag P by prolong / restrict (:101-111) -- here the adjoint of the level's stencil.
      const bool have = host->have_dagger_stencil && host->dagger_stencils;
      dag.assign(n + 1, (glb_operator*)0);
      if (host->matrix_vector_dagger) dag[0] = glb200_operator_from_callback(host->matrix_vector_dagger, host->matrix_extra_data);
      for (int l = dag[0] ? 1 : 0; l <= n; l++)
        dag[l] = (have && host->dagger_stencils[l] && host->dagger_stencils[l]->generated) ? upload_stencil(ctx, host->dagger_stencils[l])
                                                                                          : upload_adjoint_stencil(ctx, host->stencils[l]);
    }
  } catch (...) {
    release();
    throw;
  }
  mg = mg_operator_struct_complex_dev();
  mg.n_refine = n;
  mg.stencils = ops.data();
  mg.transfers = trs.data();
  mg.dagger_stencils = dag.empty() ? 0 : dag.data();
  mg.curr_level = host->curr_level;
  mg.dslash_count = host->dslash_count;  // the caller's counters keep counting
  pc = mg_precond_struct_complex_dev();
  pc.mgstruct = &mg;
}
void Hierarchy::set_precond(const mg_precond_struct_complex* p) {
  pc.in_smooth_type = p->in_smooth_type;
  pc.omega_smooth = p->omega_smooth;
  pc.n_pre_smooth = p->n_pre_smooth;
  pc.n_post_smooth = p->n_post_smooth;
  pc.normal_eqn_mg = p->normal_eqn_mg;
  pc.normal_eqn_smooth = p->normal_eqn_smooth;
  pc.mlevel_type = p->mlevel_type;
  pc.in_solve_type = p->in_solve_type;
  pc.n_max = p->n_max;
  pc.n_restart = p->n_restart;
  pc.rel_res = p->rel_res;
  pc.quiet = false;  // the reference prints its "[MG]: ..." / "[Lk ...]: ..." lines unconditionally
  mg.curr_level = p->mgstruct->curr_level;
}
void Hierarchy::release() {
  for (size_t i = 0; i < ops.size(); i++)
    if (ops[i]) glb_op_destroy(ops[i]);
  for (size_t i = 0; i < dag.size(); i++)
    if (dag[i]) glb_op_destroy(dag[i]);
  dag.clear();
  for (size_t i = 0; i < trs.size(); i++)
    if (trs[i]) glb_mg_transfer_destroy(trs[i]);
  ops.clear();
  trs.clear();
}
Hierarchy::~Hierarchy() { release(); }

}  // namespace glb200_mg_host

using namespace glb200_mg_host;

namespace {

// run `body(ctx, device vectors...)` with host vectors copied in and selected ones copied back
struct DevVec {
  glb_context* ctx;
  zcplx* d;
  size_t n;
  DevVec(glb_context* c, size_t len, const zcplx* init) : ctx(c), d(0), n(len) {
    void* p = 0;
    GLBX(glb_vec_alloc(ctx, GLB_COMPLEX, n, &p));
    d = (zcplx*)p;
    if (init)
      GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d, init));
    else
      GLBX(glb_vec_zero(ctx, GLB_COMPLEX, n, d));
  }
  void to_host(zcplx* h) { GLBX(glb_vec_download(ctx, GLB_COMPLEX, n, h, d)); }
  ~DevVec() {
    if (d) glb_vec_free(ctx, d);
  }

 private:
  DevVec(const DevVec&);
  DevVec& operator=(const DevVec&);
};

void fail_hard(const char* what, const std::exception& e) {
  std::cerr << "[glb200] " << what << " failed: " << e.what() << std::endl;
  std::abort();  // these interfaces have no error channel
}

void apply_level_stencil(zcplx* lhs, zcplx* rhs, mg_operator_struct_complex* mg, int level, bool dagger) {
  stencil_2d* st = 0;
  if (dagger && level == 0 && mg->matrix_vector_dagger) {  // mg_complex.cpp:119-123: the top level daggers by FUNCTION
    mg->matrix_vector_dagger(lhs, rhs, mg->matrix_extra_data);
    return;
  }
  if (dagger && mg->have_dagger_stencil && mg->dagger_stencils && mg->dagger_stencils[level] && mg->dagger_stencils[level]->generated)
    st = mg->dagger_stencils[level];
  if (st) {
    apply_stencil_2d(lhs, rhs, (void*)st);
    return;
  }
  st = mg->stencils[level];
  if (!st || !st->generated) {
    if (level == 0 && !dagger && mg->matrix_vector) {  // mg_complex.cpp:80-84
      mg->matrix_vector(lhs, rhs, mg->matrix_extra_data);
      return;
    }
    if (level == 0 && dagger && mg->matrix_vector_dagger) {
      mg->matrix_vector_dagger(lhs, rhs, mg->matrix_extra_data);
      return;
    }
    std::cerr << "[glb200] multigrid (host interface): level " << level << " has no generated stencil" << std::endl;
    std::abort();
  }
  if (!dagger)
    apply_stencil_2d(lhs, rhs, (void*)st);
  else if (level == 0)
    apply_square_staggered_dagger_eo_stencil(lhs, rhs, (void*)st);  // epsilon D epsilon
  else
    apply_square_staggered_dagger_tb_stencil(lhs, rhs, (void*)st);  // sigma_3 D sigma_3
}

}  // namespace

// ---------------------------------------------------------------------------------------------- operators of a level
// mg_complex.cpp:28-185
void fine_square_staggered(zcplx* lhs, zcplx* rhs, void* e) {
  mg_operator_struct_complex* mg = (mg_operator_struct_complex*)e;
  apply_level_stencil(lhs, rhs, mg, mg->curr_level, false);
}
void coarse_square_staggered(zcplx* lhs, zcplx* rhs, void* e) {
  mg_operator_struct_complex* mg = (mg_operator_struct_complex*)e;
  apply_level_stencil(lhs, rhs, mg, mg->curr_level + 1, false);
}
// mg_complex.cpp:93-133.  The top level daggers by function (or, without one, through its dagger stencil / epsilon D epsilon);
// a level below it through its dagger stencil, and without one as the reference does: prolong, dagger one level up,
// restrict -- P^dag D^dag P, the adjoint of the Galerkin operator whatever the null vectors look like.
void fine_square_staggered_dagger(zcplx* lhs, zcplx* rhs, void* e) {
  mg_operator_struct_complex* mg = (mg_operator_struct_complex*)e;
  if (mg->curr_level == 0) {
    apply_level_stencil(lhs, rhs, mg, 0, true);
  } else {
    level_up(mg);
    coarse_square_staggered_dagger(lhs, rhs, e);
    level_down(mg);
  }
}
void coarse_square_staggered_dagger(zcplx* lhs, zcplx* rhs, void* e) {
  mg_operator_struct_complex* mg = (mg_operator_struct_complex*)e;
  const int level = mg->curr_level + 1;
  if (mg->have_dagger_stencil && mg->dagger_stencils && mg->dagger_stencils[level] && mg->dagger_stencils[level]->generated) {
    apply_stencil_2d(lhs, rhs, (void*)mg->dagger_stencils[level]);
    return;
  }
  std::vector<zcplx> fine_in(mg->curr_fine_size), fine_out(mg->curr_fine_size);
  prolong(fine_in.data(), rhs, mg);
  fine_square_staggered_dagger(fine_out.data(), fine_in.data(), e);
  restrict(lhs, fine_out.data(), mg);
}
void fine_square_staggered_normal(zcplx* lhs, zcplx* rhs, void* e) {
  mg_operator_struct_complex* mg = (mg_operator_struct_complex*)e;
  std::vector<zcplx> tmp(mg->curr_fine_size);
  fine_square_staggered(tmp.data(), rhs, e);
  fine_square_staggered_dagger(lhs, tmp.data(), e);
}
void coarse_square_staggered_normal(zcplx* lhs, zcplx* rhs, void* e) {
  mg_operator_struct_complex* mg = (mg_operator_struct_complex*)e;
  std::vector<zcplx> tmp(mg->curr_coarse_size);
  coarse_square_staggered(tmp.data(), rhs, e);
  coarse_square_staggered_dagger(lhs, tmp.data(), e);
}

// ---------------------------------------------------------------------------------------------- transfers, set-up pieces
// mg_complex.cpp:372-467
void prolong(zcplx* x_fine, zcplx* x_coarse, mg_operator_struct_complex* mg) {
  try {
    glb_context* ctx = glb200_default_context();
    glb_mg_transfer* t = upload_transfer(ctx, mg, mg->curr_level);
    {
      DevVec c(ctx, glb_mg_coarse_size(t), x_coarse), f(ctx, glb_mg_fine_size(t), 0);
      GLBX(glb_mg_prolong(t, f.d, c.d));
      f.to_host(x_fine);
    }
    glb_mg_transfer_destroy(t);
  } catch (const std::exception& e) {
    fail_hard("prolong", e);
  }
}
void restrict(zcplx* x_coarse, zcplx* x_fine, mg_operator_struct_complex* mg) {
  try {
    glb_context* ctx = glb200_default_context();
    glb_mg_transfer* t = upload_transfer(ctx, mg, mg->curr_level);
    {
      DevVec f(ctx, glb_mg_fine_size(t), x_fine), c(ctx, glb_mg_coarse_size(t), 0);
      GLBX(glb_mg_restrict(t, c.d, f.d));
      c.to_host(x_coarse);
    }
    glb_mg_transfer_destroy(t);
  } catch (const std::exception& e) {
    fail_hard("restrict", e);
  }
}

namespace {
// the null vectors of the current level on the device; written back to the host arrays by `sync_back`
struct DevNull {
  glb_context* ctx;
  std::vector<zcplx*> v;
  size_t n;
  DevNull(glb_context* c, mg_operator_struct_complex* mg) : ctx(c), n((size_t)mg->curr_fine_size) {
    const int nv = mg->n_vectors[mg->curr_level];
    v.assign(nv, (zcplx*)0);
    for (int i = 0; i < nv; i++) {
      void* p = 0;
      GLBX(glb_vec_alloc(ctx, GLB_COMPLEX, n, &p));
      v[i] = (zcplx*)p;
      GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, p, mg->null_vectors[mg->curr_level][i]));
    }
  }
  void sync_back(mg_operator_struct_complex* mg) {
    for (size_t i = 0; i < v.size(); i++) GLBX(glb_vec_download(ctx, GLB_COMPLEX, n, mg->null_vectors[mg->curr_level][i], v[i]));
  }
  ~DevNull() {
    for (size_t i = 0; i < v.size(); i++)
      if (v[i]) glb_vec_free(ctx, v[i]);
  }
};
}  // namespace

// mg_complex.cpp:259-370 (ends with block_normalize, :191-256)
void block_orthonormalize(mg_operator_struct_complex* mg) {
  try {
    glb_context* ctx = glb200_default_context();
    DevNull N(ctx, mg);
    GLBX(glb_mg_block_orthonormalize(ctx, mg->curr_x_fine, mg->curr_y_fine, mg->curr_dof_fine, mg->blocksize_x[mg->curr_level],
                                     mg->blocksize_y[mg->curr_level], mg->n_vectors[mg->curr_level], (void* const*)N.v.data()));
    N.sync_back(mg);
  } catch (const std::exception& e) {
    fail_hard("block_orthonormalize", e);
  }
}
// mg_complex.cpp:191-256: every null vector of the current level scaled to unit norm on every block -- the
// orthonormalisation kernel handed one vector at a time (nothing to project out, its closing normalisation remains)
void block_normalize(mg_operator_struct_complex* mg) {
  try {
    glb_context* ctx = glb200_default_context();
    DevNull N(ctx, mg);
    for (int j = 0; j < mg->n_vectors[mg->curr_level]; j++)
      GLBX(glb_mg_block_orthonormalize(ctx, mg->curr_x_fine, mg->curr_y_fine, mg->curr_dof_fine, mg->blocksize_x[mg->curr_level],
                                       mg->blocksize_y[mg->curr_level], 1, (void* const*)(N.v.data() + j)));
    N.sync_back(mg);
  } catch (const std::exception& e) {
    fail_hard("block_normalize", e);
  }
}
// mg_complex.cpp:827-1026: P^dag A P of the fine stencil with the null vectors of the current level, written into the
// (allocated, not yet generated) coarse stencil
void generate_coarse_from_fine_stencil(stencil_2d* coarse, stencil_2d* fine, mg_operator_struct_complex* mg, bool ignore_shifts) {
  if (coarse->generated || fine->stencil_size > 2 || coarse->stencil_size > 2) return;  // :829-832
  try {
    glb_context* ctx = glb200_default_context();
    if (fine->stencil_size != 1 || coarse->stencil_size != 1)
      throw Error("generate_coarse_from_fine_stencil: one-link stencils only on the accelerated path");
    glb_operator* f = upload_stencil(ctx, fine);
    glb_mg_transfer* t = 0;
    glb_operator* c = 0;
    try {
      t = upload_transfer(ctx, mg, mg->curr_level);
      GLBX(glb_mg_galerkin(t, f, ignore_shifts ? 1 : 0, &c));
      GLBX(glb_op_stencil_download(c, coarse->clover, coarse->hopping));
    } catch (...) {
      if (c) glb_op_destroy(c);
      if (t) glb_mg_transfer_destroy(t);
      glb_op_destroy(f);
      throw;
    }
    glb_op_destroy(c);
    glb_mg_transfer_destroy(t);
    glb_op_destroy(f);
    coarse->generated = true;
  } catch (const std::exception& e) {
    fail_hard("generate_coarse_from_fine_stencil", e);
  }
}
void generate_coarse_from_fine_stencil(stencil_2d* coarse, stencil_2d* fine, mg_operator_struct_complex* mg) {
  generate_coarse_from_fine_stencil(coarse, fine, mg, false);
}

// ---------------------------------------------------------------------------------------------- the cycle
// mg_complex.cpp:514-822 with host vectors: upload the hierarchy, run mg_preconditioner_dev, download
void mg_preconditioner(zcplx* lhs, zcplx* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  try {
    mg_precond_struct_complex* p = (mg_precond_struct_complex*)extra_data;
    Hierarchy H(p->mgstruct, p->normal_eqn_smooth || p->normal_eqn_mg);
    H.set_precond(p);
    DevVec l(H.ctx, (size_t)size, lhs), r(H.ctx, (size_t)size, rhs);
    mg_preconditioner_dev(l.d, r.d, size, (void*)&H.pc, verb);
    l.to_host(lhs);
  } catch (const std::exception& e) {
    fail_hard("mg_preconditioner", e);
  }
}

// ---------------------------------------------------------------------------------------------- null vectors
namespace {
// a device set-up struct around the CURRENT level of a host struct
struct DevSetup {
  glb_context* ctx;
  mg_operator_struct_complex_dev mg;
  std::vector<glb_operator*> ops;
  std::vector<glb_mg_transfer*> trs;
  std::vector<zcplx**> table;
  DevNull N;
  DevSetup(mg_operator_struct_complex* host, bool with_operator, bool with_shifts)
      : ctx(glb200_default_context()), N(glb200_default_context(), host) {
    const int n = host->n_refine, lvl = host->curr_level;
    ops.assign(n + 1, (glb_operator*)0);
    trs.assign(n, (glb_mg_transfer*)0);
    table.assign(n, (zcplx**)0);
    table[lvl] = N.v.data();
    mg = mg_operator_struct_complex_dev();
    mg.n_refine = n;
    mg.stencils = ops.data();
    mg.transfers = trs.data();
    mg.curr_level = lvl;
    mg.dslash_count = host->dslash_count;
    mg.x_fine = host->latt[0]->get_lattice_dimension(0);
    mg.y_fine = host->latt[0]->get_lattice_dimension(1);
    mg.blocksize_x = host->blocksize_x;
    mg.blocksize_y = host->blocksize_y;
    mg.n_vectors = host->n_vectors;
    mg.null_vectors = table.data();
    // the partitions only need the lattice of the level: any operator of the right size will do for the context
    ops[lvl] = upload_stencil(ctx, host->stencils[lvl]);
    (void)with_operator;
    if (with_shifts) {
      void (*sx)(zcplx*, zcplx*, void*) = &staggered_symmshift_x;
      void (*sy)(zcplx*, zcplx*, void*) = &staggered_symmshift_y;
      mg.symmshift_x = glb200_operator_from_callback(sx, host->matrix_extra_data);
      mg.symmshift_y = glb200_operator_from_callback(sy, host->matrix_extra_data);
    }
  }
  ~DevSetup() {
    for (size_t i = 0; i < ops.size(); i++)
      if (ops[i]) glb_op_destroy(ops[i]);
    if (mg.symmshift_x) glb_op_destroy(mg.symmshift_x);
    if (mg.symmshift_y) glb_op_destroy(mg.symmshift_y);
  }
};
}  // namespace

void null_partition_staggered(mg_operator_struct_complex* mg, int num_null_vec, blocking_strategy bstrat, Lattice*) {
  if (bstrat == BLOCK_NONE) return;
  try {
    DevSetup S(mg, false, bstrat == BLOCK_TOPO);
    null_partition_staggered_dev(&S.mg, num_null_vec, bstrat);
    S.N.sync_back(mg);
  } catch (const std::exception& e) {
    fail_hard("null_partition_staggered", e);
  }
}
void null_partition_coarse(mg_operator_struct_complex* mg, int num_null_vec, blocking_strategy bstrat) {
  if (bstrat == BLOCK_NONE) return;
  try {
    DevSetup S(mg, false, false);
    null_partition_coarse_dev(&S.mg, num_null_vec, bstrat);
    S.N.sync_back(mg);
  } catch (const std::exception& e) {
    fail_hard("null_partition_coarse", e);
  }
}
void null_generate_free(mg_operator_struct_complex* mg, null_vector_params* nv, bool do_gauge_transform, zcplx* gauge_trans) {
  try {
    DevSetup S(mg, false, nv->bstrat == BLOCK_TOPO && mg->curr_level == 0);
    null_generate_free_dev(&S.mg, nv, do_gauge_transform, gauge_trans);
    S.N.sync_back(mg);
  } catch (const std::exception& e) {
    fail_hard("null_generate_free", e);
  }
}
void null_generate_random_smooth(mg_operator_struct_complex* mg, null_vector_params* nv, inversion_verbose_struct* verb,
                                 std::mt19937* generator) {
  try {
    DevSetup S(mg, true, nv->bstrat == BLOCK_TOPO && mg->curr_level == 0);
    null_vector_params local = *nv;
    local.quiet = false;  // the reference prints its cosines unconditionally
    null_generate_random_smooth_dev(&S.mg, &local, verb, generator);
    S.N.sync_back(mg);
  } catch (const std::exception& e) {
    fail_hard("null_generate_random_smooth", e);
  }
}
