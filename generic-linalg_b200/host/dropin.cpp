// dropin.cpp -- the reference-facing entry points: solvers called with HOST vectors and the
// reference's operator callbacks, and those callbacks themselves.
//
// A host-pointer solve = recognise the callback -> build (or reuse) the device operator ->
// upload phi, phi0 -> run the device shell (dev_solvers.cpp) -> download phi.  The callbacks of
// operators.h / coarse_stencil.h, when called directly, do upload -> device apply -> download.
// There is no CPU compute path: an unrecognised callback is an error unless the explicit parity
// shim (glb200_allow_host_callback_shim) is switched on.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

#include "coarse_stencil.h"
#include "dev_internal.hpp"
#include "generic_eigenvalues.h"
#include "generic_inverters_precond.h"
#include "mg_complex.h"
#include "operators.h"
#include "operators_stencil.h"

using namespace glbx;

namespace {

glb_context* g_default_ctx = 0;
bool g_allow_shim = false;
bool g_shim_from_env_used = false;  // the shim was switched on by GLB200_HOST_CALLBACKS, not by the program
bool g_cache_ops = false;

enum Builtin {
  B_NONE = 0, B_LAPLACE_NC, B_LAPLACE_NC_REAL, B_LAPLACE_U1, B_STAG_FREE, B_STAG_U1, B_GAMMA5, B_STAG_G5_FREE,
  B_STAG_G5_U1, B_STAG_DAGGER_U1, B_STAG_NORMAL_U1, B_STAG_DEO_U1, B_STAG_DOE_U1, B_STAG_M2MDEODOE_U1, B_LAPLACIAN_REAL,
  B_LAPLACIAN_IMAG, B_STENCIL, B_STAG_FREE_REAL,
  // composite views of a stencil_2d (include/glb200.h GLB_SV_*), in that order
  B_SV_M2MDEODOE, B_SV_M2MDTBDBT, B_SV_NORMAL_EO, B_SV_NORMAL_TB, B_SV_DAGGER_EO, B_SV_DAGGER_TB,
  // the rest of operators.h: stencils built on the host from the links (one upload), and the index operator,
  // a composition of three device operators
  B_SYMMSHIFT_X, B_SYMMSHIFT_Y, B_STAG_2LINK, B_STAG_INDEX,
  // the level operators of a host multigrid struct (mg_complex.h): extra_info is a mg_operator_struct_complex*
  B_MG_FINE, B_MG_COARSE, B_MG_FINE_DAGGER, B_MG_COARSE_DAGGER, B_MG_FINE_NORMAL, B_MG_COARSE_NORMAL
};

Builtin classify(void (*fn)(zcplx*, zcplx*, void*)) {
  typedef void (*F)(zcplx*, zcplx*, void*);
  if (fn == (F)&square_laplace) return B_LAPLACE_NC;
  if (fn == (F)&square_laplace_u1) return B_LAPLACE_U1;
  if (fn == (F)&square_staggered) return B_STAG_FREE;
  if (fn == (F)&square_staggered_u1) return B_STAG_U1;
  if (fn == (F)&gamma_5) return B_GAMMA5;
  if (fn == (F)&square_staggered_gamma5) return B_STAG_G5_FREE;
  if (fn == (F)&square_staggered_gamma5_u1) return B_STAG_G5_U1;
  if (fn == (F)&square_staggered_dagger_u1) return B_STAG_DAGGER_U1;
  if (fn == (F)&square_staggered_normal_u1) return B_STAG_NORMAL_U1;
  if (fn == (F)&square_staggered_deo_u1) return B_STAG_DEO_U1;
  if (fn == (F)&square_staggered_doe_u1) return B_STAG_DOE_U1;
  if (fn == (F)&square_staggered_m2mdeodoe_u1) return B_STAG_M2MDEODOE_U1;
  if (fn == (F)&square_laplacian) return B_LAPLACIAN_IMAG;
  if (fn == (F)&apply_stencil_2d) return B_STENCIL;
  if (fn == (F)&staggered_symmshift_x) return B_SYMMSHIFT_X;
  if (fn == (F)&staggered_symmshift_y) return B_SYMMSHIFT_Y;
  if (fn == (F)&square_staggered_2linklaplace_u1) return B_STAG_2LINK;
  if (fn == (F)&staggered_index_operator) return B_STAG_INDEX;
  if (fn == (F)&fine_square_staggered) return B_MG_FINE;
  if (fn == (F)&coarse_square_staggered) return B_MG_COARSE;
  if (fn == (F)&fine_square_staggered_dagger) return B_MG_FINE_DAGGER;
  if (fn == (F)&coarse_square_staggered_dagger) return B_MG_COARSE_DAGGER;
  if (fn == (F)&fine_square_staggered_normal) return B_MG_FINE_NORMAL;
  if (fn == (F)&coarse_square_staggered_normal) return B_MG_COARSE_NORMAL;
  if (fn == (F)&apply_square_staggered_m2mdeodoe_stencil) return B_SV_M2MDEODOE;
  if (fn == (F)&apply_square_staggered_m2mdtbdbt_stencil) return B_SV_M2MDTBDBT;
  if (fn == (F)&apply_square_staggered_normal_eo_stencil) return B_SV_NORMAL_EO;
  if (fn == (F)&apply_square_staggered_normal_tb_stencil) return B_SV_NORMAL_TB;
  if (fn == (F)&apply_square_staggered_dagger_eo_stencil) return B_SV_DAGGER_EO;
  if (fn == (F)&apply_square_staggered_dagger_tb_stencil) return B_SV_DAGGER_TB;
  return B_NONE;
}
Builtin classify(void (*fn)(double*, double*, void*)) {
  typedef void (*F)(double*, double*, void*);
  if (fn == (F)&square_laplace) return B_LAPLACE_NC_REAL;
  if (fn == (F)&square_laplacian) return B_LAPLACIAN_REAL;
  if (fn == (F)&square_staggered) return B_STAG_FREE_REAL;
  return B_NONE;
}

glb_operator* build(Builtin kind, void* extra) {
  glb_context* ctx = glb200_default_context();
  glb_operator* op = 0;
  staggered_u1_op* s = (staggered_u1_op*)extra;
  switch (kind) {
    case B_LAPLACE_NC:  // operators.cpp:28 : diag (4+mass)
      GLBX(glb_op_create_laplace(ctx, GLB_COMPLEX, s->x_fine, s->y_fine, s->Nc, 4 + s->mass, 0.0, &op));
      break;
    case B_LAPLACE_NC_REAL:
      GLBX(glb_op_create_laplace(ctx, GLB_REAL, s->x_fine, s->y_fine, s->Nc, 4 + s->mass, 0.0, &op));
      break;
    case B_STAG_FREE_REAL: GLBX(glb_op_create_staggered_free_real(ctx, s->x_fine, s->y_fine, s->mass, &op)); break;
    case B_LAPLACE_U1: GLBX(glb_op_create_laplace_u1(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, &op)); break;
    case B_STAG_FREE: GLBX(glb_op_create_staggered(ctx, 0, s->x_fine, s->y_fine, s->mass, 0, &op)); break;
    case B_STAG_U1: GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, 0, &op)); break;
    case B_GAMMA5: GLBX(glb_op_create_gamma5(ctx, s->x_fine, s->y_fine, &op)); break;
    case B_STAG_G5_FREE:
      GLBX(glb_op_create_staggered(ctx, 0, s->x_fine, s->y_fine, s->mass, GLB_STAG_GAMMA5, &op));
      break;
    case B_STAG_G5_U1:
      GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, GLB_STAG_GAMMA5, &op));
      break;
    case B_STAG_DAGGER_U1:
      GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, GLB_STAG_DAGGER, &op));
      break;
    case B_STAG_NORMAL_U1:
      GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, GLB_STAG_NORMAL, &op));
      break;
    case B_STAG_DEO_U1:
      GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, GLB_STAG_DEO, &op));
      break;
    case B_STAG_DOE_U1:
      GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, GLB_STAG_DOE, &op));
      break;
    case B_STAG_M2MDEODOE_U1:
      GLBX(glb_op_create_staggered(ctx, s->lattice, s->x_fine, s->y_fine, s->mass, GLB_STAG_M2MDEODOE, &op));
      break;
    case B_LAPLACIAN_REAL: {  // square_laplace.cpp:182 : (4+MASS)
      laplace_op* l = (laplace_op*)extra;
      GLBX(glb_op_create_laplace(ctx, GLB_REAL, l->N, l->N, 1, 4 + l->mass_sq, 0.0, &op));
      break;
    }
    case B_LAPLACIAN_IMAG: {  // imag_laplace.cpp:126 : (4.0+MASS+i)
      laplace_op* l = (laplace_op*)extra;
      GLBX(glb_op_create_laplace(ctx, GLB_COMPLEX, l->N, l->N, 1, 4.0 + l->mass_sq, 1.0, &op));
      break;
    }
    case B_STENCIL: {
      stencil_2d* st = (stencil_2d*)extra;
      if (st->lat->get_nd() != 2) throw Error("apply_stencil_2d: 2-d lattices only");
      const int X = st->lat->get_lattice_dimension(0), Y = st->lat->get_lattice_dimension(1), nc = st->lat->get_nc();
      if (st->sdir != DIR_ALL) {
        // one direction of the stencil (coarse_stencil.cpp:173-393 and the same switch in the partial applies): the
        // reference's set-up and tests use it, no solver does.  Same kernels on a copy that keeps only that plane --
        // the site term DIR_0 carries the three shifts, a hopping / two-link direction none -- so every other term
        // contributes an exact zero.
        const int d = (int)st->sdir;
        if (d < (int)DIR_0 || d > (int)DIR_XP1YM1) throw Error("apply_stencil_2d: unknown stencil direction");
        if (d >= (int)DIR_XP2 && !st->has_two) throw Error("apply_stencil_2d: two-link direction of a one-link stencil");
        const size_t plane = (size_t)nc * nc * X * Y;
        std::vector<zcplx> cl(plane), hp(4 * plane), tw(st->has_two ? 8 * plane : 0);
        if (d == (int)DIR_0) std::copy(st->clover, st->clover + plane, cl.begin());
        else if (d < (int)DIR_XP2) std::copy(st->hopping + (d - (int)DIR_XP1) * plane, st->hopping + (d - (int)DIR_XP1 + 1) * plane, hp.begin() + (d - (int)DIR_XP1) * plane);
        else std::copy(st->two_link + (d - (int)DIR_XP2) * plane, st->two_link + (d - (int)DIR_XP2 + 1) * plane, tw.begin() + (d - (int)DIR_XP2) * plane);
        const bool site = d == (int)DIR_0;
        const double sh[2] = {site ? st->shift.real() : 0.0, site ? st->shift.imag() : 0.0};
        const double eo[2] = {site ? st->eo_shift.real() : 0.0, site ? st->eo_shift.imag() : 0.0};
        const double df[2] = {site ? st->dof_shift.real() : 0.0, site ? st->dof_shift.imag() : 0.0};
        GLBX(glb_op_create_stencil2d(ctx, cl.data(), hp.data(), st->has_two ? tw.data() : 0, X, Y, nc, sh, eo, df, &op));
        break;
      }
      const double sh[2] = {st->shift.real(), st->shift.imag()};
      const double eo[2] = {st->eo_shift.real(), st->eo_shift.imag()};
      const double df[2] = {st->dof_shift.real(), st->dof_shift.imag()};
      GLBX(glb_op_create_stencil2d(ctx, st->clover, st->hopping, st->has_two ? st->two_link : 0, X, Y, nc, sh, eo, df, &op));
      break;
    }
    case B_MG_FINE:
    case B_MG_COARSE: {  // mg_complex.cpp:28-92: the stencil of the current level / of the level below it
      mg_operator_struct_complex* mg = (mg_operator_struct_complex*)extra;
      const int level = mg->curr_level + (kind == B_MG_COARSE ? 1 : 0);
      stencil_2d* st = mg->stencils ? mg->stencils[level] : 0;
      if (st && st->generated) {
        op = glb200_mg_host::upload_stencil(ctx, st);
      } else if (level == 0 && mg->matrix_vector) {  // no stencil on the top level: its function operator (:80-84)
        const Builtin inner = classify(mg->matrix_vector);
        if (inner == B_NONE || (inner >= B_MG_FINE && inner <= B_MG_COARSE_NORMAL) || inner == B_STAG_INDEX)
          throw Error("fine_square_staggered: the top-level operator is not a known device operator");
        op = build(inner, mg->matrix_extra_data);
      } else {
        throw Error("fine_/coarse_square_staggered: the level has no generated stencil");
      }
      break;
    }
    case B_MG_FINE_DAGGER:
    case B_MG_COARSE_DAGGER: {  // mg_complex.cpp:93-133: the top level applies the FUNCTION matrix_vector_dagger, the levels
                                // below it their dagger stencil, else the adjoint of their stencil
      mg_operator_struct_complex* mg = (mg_operator_struct_complex*)extra;
      const int level = mg->curr_level + (kind == B_MG_COARSE_DAGGER ? 1 : 0);
      if (level == 0) {
        const Builtin inner = mg->matrix_vector_dagger ? classify(mg->matrix_vector_dagger) : B_NONE;
        if (inner == B_NONE || (inner >= B_MG_FINE && inner <= B_MG_COARSE_NORMAL) || inner == B_STAG_INDEX)
          throw Error("fine_square_staggered_dagger: matrix_vector_dagger is not a known device operator");
        op = build(inner, mg->matrix_extra_data);
      } else {
        stencil_2d* st = (mg->have_dagger_stencil && mg->dagger_stencils) ? mg->dagger_stencils[level] : 0;
        if (st && st->generated)
          op = glb200_mg_host::upload_stencil(ctx, st);
        else  // the reference projects P^dag D^dag P here (:101-111): the adjoint of the level's stencil
          op = glb200_mg_host::upload_adjoint_stencil(ctx, mg->stencils ? mg->stencils[level] : 0);
      }
      break;
    }
    case B_MG_FINE_NORMAL:
    case B_MG_COARSE_NORMAL:
      break;  // a composition of two device operators, like the index operator: no single operator (lease() builds it)
    case B_SYMMSHIFT_X:
    case B_SYMMSHIFT_Y:
    case B_STAG_2LINK: {
      // nc = 1 stencils of operators.cpp:625-778, entries computed on the host with the reference's products
      // (scaling by 1/2 is exact, so (U/2) psi = (U psi)/2 bit for bit).  Plane order: +x, +y, -x, -y;
      // two-link: +2x, +x+y, +2y, -x+y, -2x, -x-y, -2y, +x-y (coarse_stencil.h:46-60).
      staggered_u1_op* s = (staggered_u1_op*)extra;
      const int X = s->x_fine, Y = s->y_fine, V = X * Y;
      const zcplx* U = s->lattice;
      std::vector<zcplx> cl(V, zcplx(0.0)), hp(4 * (size_t)V, zcplx(0.0)), tl;
      const bool two = (kind == B_STAG_2LINK);
      if (two) tl.assign(8 * (size_t)V, zcplx(0.0));
      for (int i = 0; i < V; i++) {
        const int x = i % X, y = i / X;
        const int xm = (x + X - 1) % X, ym = (y + Y - 1) % Y, xp = (x + 1) % X, yp = (y + 1) % Y;
        const int xmm = (x + X - 2) % X, ymm = (y + Y - 2) % Y;
        const double eta1 = 1 - 2 * (x % 2);
        if (kind == B_SYMMSHIFT_X) {
          hp[i] = 0.5 * U[y * X * 2 + x * 2];
          hp[i + 2 * (size_t)V] = 0.5 * conj(U[y * X * 2 + xm * 2]);
        } else if (kind == B_SYMMSHIFT_Y) {
          hp[i + (size_t)V] = 0.5 * (eta1 * U[y * X * 2 + x * 2 + 1]);
          hp[i + 3 * (size_t)V] = 0.5 * (eta1 * conj(U[ym * X * 2 + x * 2 + 1]));
        } else {
          hp[i] = -0.5 * U[y * X * 2 + x * 2];
          hp[i + (size_t)V] = -0.5 * (eta1 * U[y * X * 2 + x * 2 + 1]);
          hp[i + 2 * (size_t)V] = 0.5 * conj(U[y * X * 2 + xm * 2]);
          hp[i + 3 * (size_t)V] = 0.5 * (eta1 * conj(U[ym * X * 2 + x * 2 + 1]));
          const double w = s->wilson_coeff;
          cl[i] = w * 4.0;
          tl[i] = -(w * U[y * X * 2 + x * 2] * U[y * X * 2 + xp * 2]);                              // +2x
          tl[i + 2 * (size_t)V] = -(w * U[y * X * 2 + x * 2 + 1] * U[yp * X * 2 + x * 2 + 1]);      // +2y
          tl[i + 4 * (size_t)V] = -(w * conj(U[y * X * 2 + xm * 2]) * conj(U[y * X * 2 + xmm * 2]));            // -2x
          tl[i + 6 * (size_t)V] = -(w * conj(U[ym * X * 2 + x * 2 + 1]) * conj(U[ymm * X * 2 + x * 2 + 1]));    // -2y
        }
      }
      const double sh[2] = {two ? s->mass : 0.0, 0.0}, z[2] = {0.0, 0.0};
      GLBX(glb_op_create_stencil2d(ctx, cl.data(), hp.data(), two ? tl.data() : 0, X, Y, 1, sh, z, z, &op));
      break;
    }
    case B_SV_M2MDEODOE:
    case B_SV_M2MDTBDBT:
    case B_SV_NORMAL_EO:
    case B_SV_NORMAL_TB:
    case B_SV_DAGGER_EO:
    case B_SV_DAGGER_TB: {  // the stencil itself, then the view on it (which adopts it)
      glb_operator* base = build(B_STENCIL, extra);
      const int sv = GLB_SV_M2MDEODOE + ((int)kind - (int)B_SV_M2MDEODOE);
      if (glb_op_create_stencil_view(base, sv, 1, &op) != GLB_OK) {
        glb_op_destroy(base);
        throw Error(std::string("stencil view: ") + glb_last_error());
      }
      break;
    }
    default: break;
  }
  return op;
}

// optional reuse of device operators between calls (off by default: the host arrays may change)
struct CacheKey {
  int kind;
  const void* data;
  int X, Y, Nc;
  bool operator<(const CacheKey& o) const {
    if (kind != o.kind) return kind < o.kind;
    if (data != o.data) return data < o.data;
    if (X != o.X) return X < o.X;
    if (Y != o.Y) return Y < o.Y;
    return Nc < o.Nc;
  }
};
std::map<CacheKey, glb_operator*> g_cache;

bool cacheable(Builtin k) { return k >= B_LAPLACE_NC && k <= B_STAG_M2MDEODOE_U1; }

// staggered_index_operator (operators.cpp:782-835) on the device: lhs = i D_0 rhs - (m i/2) S_x S_y rhs
// + (m i/2) S_y S_x rhs with D_0 the massless staggered operator and S the symmetric shifts -- five applies and
// three axpys of existing kernels, in the reference's order.  Usable as a device callback (extra = IndexOp*).
struct Composite {  // a composition of device operators behind ONE device callback (index_apply_dev)
  virtual void apply(zcplx* lhs, zcplx* rhs) = 0;
  virtual ~Composite() {}
};
struct IndexOp : Composite {
  glb_context* ctx;
  glb_operator *D0, *Sx, *Sy;  // D0 belongs to the lease
  zcplx *t1, *t2;
  double mass;
  size_t n;
  void apply(zcplx* lhs, zcplx* rhs);
  ~IndexOp() {
    glb_op_destroy(Sx);
    glb_op_destroy(Sy);
    if (t1) glb_vec_free(ctx, t1);
    if (t2) glb_vec_free(ctx, t2);
  }
};
// fine_ / coarse_square_staggered_normal (mg_complex.cpp:135-172): the level operator into a temporary, its dagger out
struct ChainOp : Composite {
  glb_context* ctx;
  glb_operator *first, *second;  // `first` belongs to the lease
  zcplx* tmp;
  void apply(zcplx* lhs, zcplx* rhs) {
    GLBX(glb_op_apply(first, tmp, rhs));
    GLBX(glb_op_apply(second, lhs, tmp));
  }
  ~ChainOp() {
    glb_op_destroy(second);
    if (tmp) glb_vec_free(ctx, tmp);
  }
};
void index_apply_dev(zcplx* lhs, zcplx* rhs, void* e) { ((Composite*)e)->apply(lhs, rhs); }
void IndexOp::apply(zcplx* lhs, zcplx* rhs) {
  IndexOp* o = this;
  Blas<zcplx> B = {o->ctx, o->n};
  GLBX(glb_op_apply(o->D0, o->t1, rhs));
  B.zero(lhs);
  B.axpy(zcplx(0.0, 1.0), o->t1, lhs);               // lhs = i D_0 rhs
  const zcplx c = o->mass * zcplx(0.0, 0.5);          // mass*cplxId2
  GLBX(glb_op_apply(o->Sy, o->t1, rhs));
  GLBX(glb_op_apply(o->Sx, o->t2, o->t1));
  B.axpy(-c, o->t2, lhs);                             // lhs -= c S_x S_y rhs
  GLBX(glb_op_apply(o->Sx, o->t1, rhs));
  GLBX(glb_op_apply(o->Sy, o->t2, o->t1));
  B.axpy(c, o->t2, lhs);                              // lhs += c S_y S_x rhs
}

struct OpLease {  // operator for the duration of one call
  glb_operator* op;
  bool owned;
  Composite* comp;  // set for a composition: `op` is then its first operator (sizes), the callback is index_apply_dev
  OpLease() : op(0), owned(false), comp(0) {}
  ~OpLease() {
    delete comp;
    if (op && owned) glb_op_destroy(op);
  }
};
// the device callback of a composite lease (complex only)
template <typename T>
struct CompositeCallback {
  static void (*get())(T*, T*, void*) { return 0; }
};
template <>
struct CompositeCallback<zcplx> {
  static void (*get())(zcplx*, zcplx*, void*) { return &index_apply_dev; }
};

void lease(Builtin kind, void* extra, OpLease* out) {
  if (kind == B_MG_FINE_NORMAL || kind == B_MG_COARSE_NORMAL) {
    const bool fine = (kind == B_MG_FINE_NORMAL);
    ChainOp* c = new ChainOp();
    c->ctx = glb200_default_context();
    c->first = c->second = 0;
    c->tmp = 0;
    out->comp = c;
    out->op = c->first = build(fine ? B_MG_FINE : B_MG_COARSE, extra);
    out->owned = true;
    c->second = build(fine ? B_MG_FINE_DAGGER : B_MG_COARSE_DAGGER, extra);
    void* p = 0;
    GLBX(glb_vec_alloc(c->ctx, GLB_COMPLEX, glb_op_local_size(c->first), &p));
    c->tmp = (zcplx*)p;
    return;
  }
  if (kind == B_STAG_INDEX) {
    staggered_u1_op* s = (staggered_u1_op*)extra;
    staggered_u1_op nomass = *s;
    nomass.mass = 0.0;  // operators.cpp:803-808: a massless kernel for i D_st
    IndexOp* c = new IndexOp();
    c->ctx = glb200_default_context();
    c->D0 = c->Sx = c->Sy = 0;
    c->t1 = c->t2 = 0;
    c->mass = s->mass;
    out->comp = c;
    out->op = c->D0 = build(B_STAG_U1, &nomass);
    out->owned = true;
    c->Sx = build(B_SYMMSHIFT_X, extra);
    c->Sy = build(B_SYMMSHIFT_Y, extra);
    c->n = glb_op_local_size(c->D0);
    void* p = 0;
    GLBX(glb_vec_alloc(c->ctx, GLB_COMPLEX, c->n, &p));
    c->t1 = (zcplx*)p;
    GLBX(glb_vec_alloc(c->ctx, GLB_COMPLEX, c->n, &p));
    c->t2 = (zcplx*)p;
    return;
  }
  if (g_cache_ops && cacheable(kind)) {
    staggered_u1_op* s = (staggered_u1_op*)extra;
    CacheKey key = {(int)kind, (const void*)s->lattice, s->x_fine, s->y_fine, s->Nc};
    std::map<CacheKey, glb_operator*>::iterator it = g_cache.find(key);
    if (it == g_cache.end()) it = g_cache.insert(std::make_pair(key, build(kind, extra))).first;
    glb_op_set_mass(it->second, s->mass);
    if (kind == B_LAPLACE_NC || kind == B_LAPLACE_NC_REAL) {  // diag depends on the mass
      glb_op_destroy(it->second);
      it->second = build(kind, extra);
    }
    out->op = it->second;
    out->owned = false;
    return;
  }
  out->op = build(kind, extra);
  out->owned = true;
}

// ---- parity shim for unknown host callbacks: download -> call -> upload
template <typename T>
struct Shim {
  void (*fn)(T*, T*, void*);
  void* extra;
  size_t n;
  std::vector<T> in, out;
};
template <typename T>
void shim_cb(T* d_lhs, T* d_rhs, void* e) {
  Shim<T>* s = (Shim<T>*)e;
  glb_context* ctx = glb200_default_context();
  GLBX(glb_vec_download(ctx, Traits<T>::dtype, s->n, s->in.data(), d_rhs));
  s->fn(s->out.data(), s->in.data(), s->extra);
  GLBX(glb_vec_upload(ctx, Traits<T>::dtype, s->n, d_lhs, s->out.data()));
}

// One host-pointer solve: resolve, upload, run `body(d_phi, d_b, cb, cb_extra)`, download.
template <typename T, typename Body>
inversion_info host_solve(const char* alg, T* phi, T* phi0, int size, void (*mv)(T*, T*, void*), void* extra,
                          Body body) {
  try {
    glb_context* ctx = glb200_default_context();
    const Builtin kind = classify(mv);
    OpLease L;
    Shim<T> shim;
    void (*cb)(T*, T*, void*) = 0;
    void* cb_extra = 0;
    if (kind != B_NONE) {
      lease(kind, extra, &L);
      if ((size_t)size != glb_op_local_size(L.op))
        throw Error(glb_comm_size(ctx) == 1 ? "`size` does not match the operator's lattice"
                                            : "`size` must be this rank's share of the lattice (X * Yloc * Nc): with a "
                                              "communicator every rank passes its own rows of the vectors");
      cb = &glb200_apply_dev;
      cb_extra = L.op;
      if (L.comp) {  // the index operator: a composition of device operators behind a device callback
        cb = CompositeCallback<T>::get();
        cb_extra = L.comp;
      }
    } else if (g_allow_shim) {
      if (g_shim_from_env_used)
        std::cerr << "[glb200] WARNING: " << alg << ": the operator callback is an unknown HOST function and "
                     "GLB200_HOST_CALLBACKS=1 is set: every apply goes device -> host function -> device (a parity aid, "
                     "not a GPU path; call glb200_allow_host_callback_shim() from the program instead, or use the "
                     "operators of operators.h / a device callback)" << std::endl;
      shim.fn = mv;
      shim.extra = extra;
      shim.n = size;
      shim.in.resize(size);
      shim.out.resize(size);
      cb = &shim_cb<T>;
      cb_extra = &shim;
    } else {
      throw Error("operator callback is not a known device operator (operators.h / coarse_stencil.h); "
                  "this library has no CPU path -- see glb200_device.h");
    }
    Blas<T> B = {ctx, (size_t)size};
    Work<T> W(B);
    T* d_phi = W.get();
    T* d_b = W.get();
    GLBX(glb_vec_upload(ctx, Traits<T>::dtype, size, d_phi, phi));
    GLBX(glb_vec_upload(ctx, Traits<T>::dtype, size, d_b, phi0));
    inversion_info inf = body(d_phi, d_b, cb, cb_extra);
    GLBX(glb_vec_download(ctx, Traits<T>::dtype, size, phi, d_phi));
    return inf;
  } catch (const std::exception& e) {
    std::cerr << "[glb200] " << alg << " aborted: " << e.what() << std::endl;
    inversion_info inf;
    inf.name = alg;
    return inf;
  }
}

template <typename T>
void direct_apply(Builtin kind, T* lhs, T* rhs, void* extra) {
  try {
    glb_context* ctx = glb200_default_context();
    OpLease L;
    lease(kind, extra, &L);
    const size_t n = glb_op_local_size(L.op);
    Blas<T> B = {ctx, n};
    Work<T> W(B);
    T* d_in = W.get();
    T* d_out = W.get();
    GLBX(glb_vec_upload(ctx, Traits<T>::dtype, n, d_in, rhs));
    if (L.comp)
      CompositeCallback<T>::get()(d_out, d_in, L.comp);
    else
      GLBX(glb_op_apply(L.op, d_out, d_in));
    GLBX(glb_vec_download(ctx, Traits<T>::dtype, n, lhs, d_out));
  } catch (const std::exception& e) {
    std::cerr << "[glb200] operator apply failed: " << e.what() << std::endl;
    std::abort();  // the callback contract has no error channel
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ context
glb_context* glb200_default_context() {
  if (!g_default_ctx) {
    int dev = 0;
    if (const char* e = std::getenv("GLB200_DEVICE"))
      dev = std::atoi(e);
    else if (const char* e2 = std::getenv("LOCAL_RANK"))
      dev = std::atoi(e2);
    glb_context* ctx = 0;
    if (glb_create(dev, &ctx) != GLB_OK) throw Error(std::string("cannot create device context: ") + glb_last_error());
    g_default_ctx = ctx;
  }
  return g_default_ctx;
}
void glb200_set_default_context(glb_context* ctx) { g_default_ctx = ctx; }
void glb200_allow_host_callback_shim(bool allow) { g_allow_shim = allow; }
namespace {
// GLB200_HOST_CALLBACKS=1 in the environment switches the shim on without a source change: programs whose operator is
// their OWN host function (typically a composition of the operators of operators.h, e.g. level_crossing.cpp:366) then
// run unmodified -- the solver's vectors stay on the device and every apply goes download -> user function -> upload.
struct ShimFromEnv {
  ShimFromEnv() {
    const char* e = std::getenv("GLB200_HOST_CALLBACKS");
    if (e && e[0] == '1') {
      g_allow_shim = true;
      g_shim_from_env_used = true;  // every solve that falls back on it says so on stderr
    }
  }
} g_shim_from_env;
}  // namespace
extern "C" void glb200_cache_operators(int on) {
  g_cache_ops = on != 0;
  if (!on) {
    for (std::map<CacheKey, glb_operator*>::iterator it = g_cache.begin(); it != g_cache.end(); ++it)
      glb_op_destroy(it->second);
    g_cache.clear();
  }
}

void glb200_apply_dev(double* d_lhs, double* d_rhs, void* h) { GLBX(glb_op_apply((glb_operator*)h, d_lhs, d_rhs)); }
void glb200_apply_dev(zcplx* d_lhs, zcplx* d_rhs, void* h) { GLBX(glb_op_apply((glb_operator*)h, d_lhs, d_rhs)); }

glb_operator* glb200_operator_from_callback(void (*mv)(double*, double*, void*), void* extra) {
  const Builtin k = classify(mv);
  return k == B_NONE ? 0 : build(k, extra);
}
glb_operator* glb200_operator_from_callback(void (*mv)(zcplx*, zcplx*, void*), void* extra) {
  const Builtin k = classify(mv);
  return k == B_NONE ? 0 : build(k, extra);
}

// ------------------------------------------------------------------------------------------ operator callbacks
int get_stencil_size(op_type opt) {  // operators.cpp:9-25
  switch (opt) {
    case STAGGERED:
    case LAPLACE:
    case LAPLACE_NC2:
    case G5_STAGGERED: return 1;
    case STAGGERED_INDEX:
    case STAGGERED_NORMAL: return 2;
  }
  return 0;
}
void square_laplace(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_LAPLACE_NC, lhs, rhs, e); }
void square_laplace(double* lhs, double* rhs, void* e) { direct_apply<double>(B_LAPLACE_NC_REAL, lhs, rhs, e); }
void square_laplace_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_LAPLACE_U1, lhs, rhs, e); }
void square_staggered(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_FREE, lhs, rhs, e); }
void square_staggered(double* lhs, double* rhs, void* e) { direct_apply<double>(B_STAG_FREE_REAL, lhs, rhs, e); }
void square_staggered_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_U1, lhs, rhs, e); }
void gamma_5(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_GAMMA5, lhs, rhs, e); }
void square_staggered_gamma5(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_G5_FREE, lhs, rhs, e); }
void square_staggered_gamma5_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_G5_U1, lhs, rhs, e); }
void square_staggered_dagger_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_DAGGER_U1, lhs, rhs, e); }
void square_staggered_normal_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_NORMAL_U1, lhs, rhs, e); }
void square_staggered_deo_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_DEO_U1, lhs, rhs, e); }
void square_staggered_doe_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_DOE_U1, lhs, rhs, e); }
void square_staggered_2linklaplace_u1(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_2LINK, lhs, rhs, e); }
void staggered_symmshift_x(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SYMMSHIFT_X, lhs, rhs, e); }
void staggered_symmshift_y(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SYMMSHIFT_Y, lhs, rhs, e); }
void staggered_index_operator(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STAG_INDEX, lhs, rhs, e); }
void square_staggered_m2mdeodoe_u1(zcplx* lhs, zcplx* rhs, void* e) {
  direct_apply<zcplx>(B_STAG_M2MDEODOE_U1, lhs, rhs, e);
}
// operators.cpp:528-545 / :574-598 with host vectors: upload, one device pass each, download
void square_staggered_eoprec_prepare(zcplx* rhs_e, zcplx* rhs_orig, void* e) {
  try {
    glb_context* ctx = glb200_default_context();
    OpLease L;
    lease(B_STAG_U1, e, &L);
    const size_t n = glb_op_local_size(L.op);
    Blas<zcplx> B = {ctx, n};
    Work<zcplx> W(B);
    zcplx *d_in = W.get(), *d_out = W.get();
    GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d_in, rhs_orig));
    GLBX(glb_stag_eoprec_prepare(L.op, d_out, d_in));
    GLBX(glb_vec_download(ctx, GLB_COMPLEX, n, rhs_e, d_out));
  } catch (const std::exception& ex) {
    std::cerr << "[glb200] square_staggered_eoprec_prepare failed: " << ex.what() << std::endl;
    std::abort();
  }
}
void square_staggered_eoprec_reconstruct(zcplx* lhs_full, zcplx* lhs_e, zcplx* rhs_o, void* e) {
  try {
    glb_context* ctx = glb200_default_context();
    OpLease L;
    lease(B_STAG_U1, e, &L);
    const size_t n = glb_op_local_size(L.op);
    Blas<zcplx> B = {ctx, n};
    Work<zcplx> W(B);
    zcplx *d_e = W.get(), *d_o = W.get(), *d_out = W.get();
    GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d_e, lhs_e));
    GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d_o, rhs_o));
    GLBX(glb_stag_eoprec_reconstruct(L.op, d_out, d_e, d_o));
    GLBX(glb_vec_download(ctx, GLB_COMPLEX, n, lhs_full, d_out));
  } catch (const std::exception& ex) {
    std::cerr << "[glb200] square_staggered_eoprec_reconstruct failed: " << ex.what() << std::endl;
    std::abort();
  }
}
void square_laplacian(double* lhs, double* rhs, void* e) { direct_apply<double>(B_LAPLACIAN_REAL, lhs, rhs, e); }
void square_laplacian(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_LAPLACIAN_IMAG, lhs, rhs, e); }
void apply_stencil_2d(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_STENCIL, lhs, rhs, e); }
void apply_square_staggered_m2mdeodoe_stencil(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SV_M2MDEODOE, lhs, rhs, e); }
void apply_square_staggered_m2mdtbdbt_stencil(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SV_M2MDTBDBT, lhs, rhs, e); }
void apply_square_staggered_normal_eo_stencil(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SV_NORMAL_EO, lhs, rhs, e); }
void apply_square_staggered_normal_tb_stencil(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SV_NORMAL_TB, lhs, rhs, e); }
void apply_square_staggered_dagger_eo_stencil(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SV_DAGGER_EO, lhs, rhs, e); }
void apply_square_staggered_dagger_tb_stencil(zcplx* lhs, zcplx* rhs, void* e) { direct_apply<zcplx>(B_SV_DAGGER_TB, lhs, rhs, e); }
// operators_stencil.cpp:179-193, :217-236; mg_complex.cpp:1211-1225, :1252-1272 with host vectors: one device pass each
static void stencil_prec(int tb, bool reconstruct, zcplx* out, zcplx* a, zcplx* b, stencil_2d* st) {
  try {
    glb_context* ctx = glb200_default_context();
    OpLease L;
    lease(B_STENCIL, (void*)st, &L);
    const size_t n = glb_op_local_size(L.op);
    Blas<zcplx> B = {ctx, n};
    Work<zcplx> W(B);
    zcplx *d_a = W.get(), *d_b = W.get(), *d_out = W.get();
    GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d_a, a));
    if (reconstruct) {
      GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d_b, b));
      GLBX(glb_stencil_prec_reconstruct(L.op, tb, d_out, d_a, d_b));
    } else {
      GLBX(glb_stencil_prec_prepare(L.op, tb, d_out, d_a));
    }
    GLBX(glb_vec_download(ctx, GLB_COMPLEX, n, out, d_out));
  } catch (const std::exception& ex) {
    std::cerr << "[glb200] preconditioned stencil prepare / reconstruct failed: " << ex.what() << std::endl;
    std::abort();
  }
}
void apply_square_staggered_eoprec_prepare_stencil(zcplx* rhs_e, zcplx* rhs_orig, stencil_2d* st) {
  stencil_prec(0, false, rhs_e, rhs_orig, 0, st);
}
void apply_square_staggered_eoprec_reconstruct_stencil(zcplx* lhs_full, zcplx* lhs_e, zcplx* rhs_o, stencil_2d* st) {
  stencil_prec(0, true, lhs_full, lhs_e, rhs_o, st);
}
void apply_square_staggered_tbprec_prepare_stencil(zcplx* rhs_t, zcplx* rhs_orig, stencil_2d* st) {
  stencil_prec(1, false, rhs_t, rhs_orig, 0, st);
}
void apply_square_staggered_tbprec_reconstruct_stencil(zcplx* lhs_full, zcplx* lhs_t, zcplx* rhs_b, stencil_2d* st) {
  stencil_prec(1, true, lhs_full, lhs_t, rhs_b, st);
}
// coarse_stencil.cpp:1515-1640 by comb probing (see coarse_stencil.h)
void generate_stencil_2d(stencil_2d* st, void (*mv)(zcplx*, zcplx*, void*), void* extra) {
  if (st->generated || st->stencil_size > 2) return;
  Lattice* lat = st->lat;
  const int X = lat->get_lattice_dimension(0), Y = lat->get_lattice_dimension(1), nc = lat->get_nc();
  const int L = lat->get_lattice_size();
  const int reach = 2 * st->stencil_size + 1;
  // smallest period >= reach that divides the extent; the whole extent (one source per row / column) otherwise
  struct Period {
    static int of(int n, int reach) {
      for (int p = reach; p < n; p++)
        if (n % p == 0) return p;
      return n;
    }
  };
  const int px = Period::of(X, reach), py = Period::of(Y, reach);
  std::vector<zcplx> rhs(L), lhs(L);
  const size_t plane = (size_t)nc * L;
  // where the operator's row of the target site t = s - offset picks up the source s: plane index and offset
  static const int hop[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
  static const int two[8][2] = {{2, 0}, {1, 1}, {0, 2}, {-1, 1}, {-2, 0}, {-1, -1}, {0, -2}, {1, -1}};
  for (int color = 0; color < nc; color++)
    for (int cy = 0; cy < py; cy++)
      for (int cx = 0; cx < px; cx++) {
        std::fill(rhs.begin(), rhs.end(), zcplx(0.0));
        for (int y = cy; y < Y; y += py)
          for (int x = cx; x < X; x += px) rhs[((size_t)y * X + x) * nc + color] = 1.0;
        std::fill(lhs.begin(), lhs.end(), zcplx(0.0));
        mv(lhs.data(), rhs.data(), extra);
        for (int y = cy; y < Y; y += py)
          for (int x = cx; x < X; x += px)
            for (int c = 0; c < nc; c++) {
              const size_t self = ((size_t)y * X + x) * nc + c;
              st->clover[color + nc * self] += lhs[self];
              for (int d = 0; d < 4; d++) {  // the site whose direction-d neighbour is the source
                const int tx = ((x - hop[d][0]) % X + X) % X, ty = ((y - hop[d][1]) % Y + Y) % Y;
                const size_t t = ((size_t)ty * X + tx) * nc + c;
                st->hopping[color + nc * t + d * plane] += lhs[t];
              }
              if (st->has_two)
                for (int d = 0; d < 8; d++) {
                  const int tx = ((x - two[d][0]) % X + X) % X, ty = ((y - two[d][1]) % Y + Y) % Y;
                  const size_t t = ((size_t)ty * X + tx) * nc + c;
                  st->two_link[color + nc * t + d * plane] += lhs[t];
                }
            }
      }
  st->generated = true;
}

static void stencil_part(zcplx* lhs, zcplx* rhs, void* e, int part) {
  try {
    glb_context* ctx = glb200_default_context();
    OpLease L;
    lease(B_STENCIL, e, &L);
    const size_t n = glb_op_local_size(L.op);
    Blas<zcplx> B = {ctx, n};
    Work<zcplx> W(B);
    zcplx *d_in = W.get(), *d_out = W.get();
    GLBX(glb_vec_upload(ctx, GLB_COMPLEX, n, d_in, rhs));
    GLBX(glb_op_apply_part(L.op, d_out, d_in, part));
    GLBX(glb_vec_download(ctx, GLB_COMPLEX, n, lhs, d_out));
  } catch (const std::exception& ex) {
    std::cerr << "[glb200] partial stencil apply failed: " << ex.what() << std::endl;
    std::abort();
  }
}
void apply_stencil_2d_eo(zcplx* lhs, zcplx* rhs, void* e) { stencil_part(lhs, rhs, e, GLB_PART_EO); }
void apply_stencil_2d_oe(zcplx* lhs, zcplx* rhs, void* e) { stencil_part(lhs, rhs, e, GLB_PART_OE); }
void apply_stencil_2d_tb(zcplx* lhs, zcplx* rhs, void* e) { stencil_part(lhs, rhs, e, GLB_PART_TB); }
void apply_stencil_2d_bt(zcplx* lhs, zcplx* rhs, void* e) { stencil_part(lhs, rhs, e, GLB_PART_BT); }

// operators_stencil.cpp:14-63 / :65-118 / :120-170 : hopping[+x] = -U_x/2, [+y] = -eta U_y/2,
// [-x] = +conj U_x(x-1)/2, [-y] = +eta conj U_y(y-1)/2 ; gamma5 variant carries the site parity
// sign and moves the mass to eo_shift; dagger variant flips the hopping sign.
static void fill_stag_stencil(stencil_2d* st, staggered_u1_op* s, double hop, bool g5) {
  if (st->generated || st->lat->get_nc() != 1) return;
  const int X = st->lat->get_lattice_dimension(0), Y = st->lat->get_lattice_dimension(1);
  const int V = X * Y;
  for (int i = 0; i < V; i++) {
    const int x = i % X, y = i / X;
    const int eta1 = 1 - 2 * (x % 2);
    const int eo = g5 ? (((x + y) % 2 == 0) ? 1 : -1) : 1;
    const int xm = (x - 1 + X) % X, ym = (y - 1 + Y) % Y;
    if (g5) {
      st->hopping[i] = -0.5 * eo * s->lattice[2 * i];
      st->hopping[i + V] = -0.5 * eo * eta1 * s->lattice[2 * i + 1];
      st->hopping[i + 2 * V] = 0.5 * eo * conj(s->lattice[2 * (y * X + xm)]);
      st->hopping[i + 3 * V] = 0.5 * eo * eta1 * conj(s->lattice[2 * (ym * X + x) + 1]);
    } else {
      st->hopping[i] = (hop * -0.5) * s->lattice[2 * i];
      st->hopping[i + V] = (hop * -0.5) * eta1 * s->lattice[2 * i + 1];
      st->hopping[i + 2 * V] = (hop * 0.5) * conj(s->lattice[2 * (y * X + xm)]);
      st->hopping[i + 3 * V] = (hop * 0.5) * eta1 * conj(s->lattice[2 * (ym * X + x) + 1]);
    }
  }
  st->shift = g5 ? 0.0 : s->mass;
  st->eo_shift = g5 ? s->mass : 0.0;
  st->dof_shift = 0.0;
  st->generated = true;
}
void get_square_staggered_u1_stencil(stencil_2d* st, staggered_u1_op* s) { fill_stag_stencil(st, s, 1.0, false); }
void get_square_staggered_gamma5_u1_stencil(stencil_2d* st, staggered_u1_op* s) { fill_stag_stencil(st, s, 1.0, true); }
void get_square_staggered_dagger_u1_stencil(stencil_2d* st, staggered_u1_op* s) { fill_stag_stencil(st, s, -1.0, false); }

// ------------------------------------------------------------------------------------------ solvers, host vectors
#define GLB200_HOST_BASIC(NAME, DEVNAME, ALG)                                                                      \
  inversion_info NAME(double* phi, double* phi0, int size, int max_iter, double res,                               \
                      void (*mv)(double*, double*, void*), void* extra, inversion_verbose_struct* verb) {          \
    return host_solve<double>(ALG, phi, phi0, size, mv, extra,                                                     \
                              [&](double* dp, double* db, void (*cb)(double*, double*, void*), void* ce) {         \
                                return DEVNAME(dp, db, size, max_iter, res, cb, ce, verb);                         \
                              });                                                                                  \
  }                                                                                                                \
  inversion_info NAME(zcplx* phi, zcplx* phi0, int size, int max_iter, double res, void (*mv)(zcplx*, zcplx*, void*), \
                      void* extra, inversion_verbose_struct* verb) {                                               \
    return host_solve<zcplx>(ALG, phi, phi0, size, mv, extra,                                                      \
                             [&](zcplx* dp, zcplx* db, void (*cb)(zcplx*, zcplx*, void*), void* ce) {              \
                               return DEVNAME(dp, db, size, max_iter, res, cb, ce, verb);                          \
                             });                                                                                   \
  }
#define GLB200_HOST_RESTART(NAME, DEVNAME, ALG)                                                                    \
  inversion_info NAME(double* phi, double* phi0, int size, int max_iter, double res, int rf,                       \
                      void (*mv)(double*, double*, void*), void* extra, inversion_verbose_struct* verb) {          \
    return host_solve<double>(ALG, phi, phi0, size, mv, extra,                                                     \
                              [&](double* dp, double* db, void (*cb)(double*, double*, void*), void* ce) {         \
                                return DEVNAME(dp, db, size, max_iter, res, rf, cb, ce, verb);                     \
                              });                                                                                  \
  }                                                                                                                \
  inversion_info NAME(zcplx* phi, zcplx* phi0, int size, int max_iter, double res, int rf,                         \
                      void (*mv)(zcplx*, zcplx*, void*), void* extra, inversion_verbose_struct* verb) {            \
    return host_solve<zcplx>(ALG, phi, phi0, size, mv, extra,                                                      \
                             [&](zcplx* dp, zcplx* db, void (*cb)(zcplx*, zcplx*, void*), void* ce) {              \
                               return DEVNAME(dp, db, size, max_iter, res, rf, cb, ce, verb);                      \
                             });                                                                                   \
  }

// generic_sor.cpp:24,122; generic_minres.cpp:22,120,128,236
#define GLB200_HOST_RELAX(T)                                                                                       \
  inversion_info minv_vector_sor(T* phi, T* phi0, int size, int max_iter, double eps, double omega,                \
                                 void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {         \
    return host_solve<T>("SOR", phi, phi0, size, mv, extra, [&](T* dp, T* db, void (*cb)(T*, T*, void*), void* ce) { \
      return minv_vector_sor_dev(dp, db, size, max_iter, eps, omega, cb, ce, verb);                                \
    });                                                                                                            \
  }                                                                                                                \
  inversion_info minv_vector_minres(T* phi, T* phi0, int size, int max_iter, double eps, double omega,             \
                                    void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {      \
    return host_solve<T>("MinRes", phi, phi0, size, mv, extra,                                                     \
                         [&](T* dp, T* db, void (*cb)(T*, T*, void*), void* ce) {                                  \
                           return minv_vector_minres_dev(dp, db, size, max_iter, eps, omega, cb, ce, verb);        \
                         });                                                                                       \
  }                                                                                                                \
  inversion_info minv_vector_minres(T* phi, T* phi0, int size, int max_iter, double eps, void (*mv)(T*, T*, void*), \
                                    void* extra, inversion_verbose_struct* verb) {                                 \
    return minv_vector_minres(phi, phi0, size, max_iter, eps, 1.0, mv, extra, verb);                               \
  }
GLB200_HOST_RELAX(double)
GLB200_HOST_RELAX(zcplx)

// generic_poweriter.cpp:23 with host vectors: phi0 goes up, the estimate comes back
eigenvalue_info eig_vector_poweriter(double* eig, double* phi0, int size, int max_iter, double relres,
                                     void (*mv)(double*, double*, void*), void* extra) {
  eigenvalue_info out;
  out.relative_diff = 0.0;
  out.iter = 0;
  out.success = false;
  out.name = "Power Iteration";
  std::vector<double> unused(size, 0.0);  // host_solve moves an in/out vector and a rhs: the start vector is the rhs
  host_solve<double>("Power Iteration", unused.data(), phi0, size, mv, extra,
                     [&](double*, double* d_phi0, void (*cb)(double*, double*, void*), void* ce) {
                       out = eig_vector_poweriter_dev(eig, d_phi0, size, max_iter, relres, cb, ce);
                       return inversion_info();
                     });
  return out;
}

GLB200_HOST_BASIC(minv_vector_cg, minv_vector_cg_dev, "CG")
GLB200_HOST_RESTART(minv_vector_cg_restart, minv_vector_cg_restart_dev, "CG")
GLB200_HOST_BASIC(minv_vector_cr, minv_vector_cr_dev, "CR")
GLB200_HOST_RESTART(minv_vector_cr_restart, minv_vector_cr_restart_dev, "CR")
GLB200_HOST_BASIC(minv_vector_gcr, minv_vector_gcr_dev, "GCR")
GLB200_HOST_RESTART(minv_vector_gcr_restart, minv_vector_gcr_restart_dev, "GCR")
GLB200_HOST_BASIC(minv_vector_bicgstab, minv_vector_bicgstab_dev, "BiCGStab")
GLB200_HOST_RESTART(minv_vector_bicgstab_restart, minv_vector_bicgstab_restart_dev, "BiCGStab")
GLB200_HOST_BASIC(minv_vector_gmres, minv_vector_gmres_dev, "GMRES")
GLB200_HOST_RESTART(minv_vector_gmres_restart, minv_vector_gmres_restart_dev, "GMRES")

#define GLB200_HOST_BICGL(T)                                                                                       \
  inversion_info minv_vector_bicgstab_l(T* phi, T* phi0, int size, int max_iter, double res, int l,                \
                                        void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {  \
    return host_solve<T>("BiCGStab-l", phi, phi0, size, mv, extra,                                                 \
                         [&](T* dp, T* db, void (*cb)(T*, T*, void*), void* ce) {                                  \
                           return minv_vector_bicgstab_l_dev(dp, db, size, max_iter, res, l, cb, ce, verb);        \
                         });                                                                                       \
  }                                                                                                                \
  inversion_info minv_vector_bicgstab_l_restart(T* phi, T* phi0, int size, int max_iter, double res, int rf, int l, \
                                                void (*mv)(T*, T*, void*), void* extra,                            \
                                                inversion_verbose_struct* verb) {                                  \
    return host_solve<T>("BiCGStab-l", phi, phi0, size, mv, extra,                                                 \
                         [&](T* dp, T* db, void (*cb)(T*, T*, void*), void* ce) {                                  \
                           return minv_vector_bicgstab_l_restart_dev(dp, db, size, max_iter, res, rf, l, cb, ce,   \
                                                                     verb);                                        \
                         });                                                                                       \
  }
GLB200_HOST_BICGL(double)
GLB200_HOST_BICGL(zcplx)

// multishift: phi[] are n_shift host vectors; the device copies are permuted by the solver exactly
// like the reference permutes the caller's pointers, and restored before download.
template <typename T>
struct MultiDev {  // signature shared by minv_vector_{cg,cr,bicgstab}_m_dev
  typedef inversion_info (*fn)(T**, T*, int, int, int, int, double, double*, void (*)(T*, T*, void*), void*, bool,
                               inversion_verbose_struct*);
};
template <typename T>
static inversion_info multi_host(const char* alg, typename MultiDev<T>::fn dev, T** phi, T* phi0, int n_shift, int size,
                                 int rfc, int max_iter, double eps, double* shifts, void (*mv)(T*, T*, void*),
                                 void* extra, bool worst_first, inversion_verbose_struct* verb) {
  try {
    glb_context* ctx = glb200_default_context();
    const Builtin kind = classify(mv);
    OpLease L;
    Shim<T> shim;
    void (*cb)(T*, T*, void*) = &glb200_apply_dev;
    void* cb_extra = 0;
    if (kind != B_NONE) {
      lease(kind, extra, &L);
      cb_extra = (void*)L.op;
      if (L.comp) {
        cb = CompositeCallback<T>::get();
        cb_extra = (void*)L.comp;
      }
    } else if (g_allow_shim) {  // the caller's own host function (tests/multishift/multishift.cpp:634)
      if (g_shim_from_env_used)
        std::cerr << "[glb200] WARNING: " << alg << ": unknown HOST operator callback served through the "
                     "GLB200_HOST_CALLBACKS=1 shim (device -> host function -> device per apply): a parity aid, not a "
                     "GPU path" << std::endl;
      shim.fn = mv;
      shim.extra = extra;
      shim.n = size;
      shim.in.resize(size);
      shim.out.resize(size);
      cb = &shim_cb<T>;
      cb_extra = &shim;
    } else {
      throw Error("operator callback is not a known device operator");
    }
    Blas<T> B = {ctx, (size_t)size};
    Work<T> W(B);
    T* d_b = W.get();
    std::vector<T*> d_phi(n_shift);
    for (int s = 0; s < n_shift; s++) d_phi[s] = W.get();
    GLBX(glb_vec_upload(ctx, Traits<T>::dtype, size, d_b, phi0));
    inversion_info inf = dev(d_phi.data(), d_b, n_shift, size, rfc, max_iter, eps, shifts, cb, cb_extra, worst_first, verb);
    for (int s = 0; s < n_shift; s++) GLBX(glb_vec_download(ctx, Traits<T>::dtype, size, phi[s], d_phi[s]));
    return inf;
  } catch (const std::exception& e) {
    std::cerr << "[glb200] " << alg << " aborted: " << e.what() << std::endl;
    inversion_info inf(n_shift > 0 ? n_shift : 1);  // callers index resSqmrhs[] without looking at `success`
    for (int s = 0; s < inf.n_rhs; s++) inf.resSqmrhs[s] = 0.0;
    inf.name = alg;
    return inf;
  }
}
template <typename T>
static inversion_info cg_m_host(T** phi, T* phi0, int n_shift, int size, int rfc, int max_iter, double eps,
                                double* shifts, void (*mv)(T*, T*, void*), void* extra, bool worst_first,
                                inversion_verbose_struct* verb) {
  typename MultiDev<T>::fn dev = &minv_vector_cg_m_dev;
  return multi_host<T>("CG-M", dev, phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra, worst_first, verb);
}
#define GLB200_HOST_MULTI(T)                                                                                        \
  inversion_info minv_vector_cr_m(T** phi, T* phi0, int n_shift, int size, int rfc, int max_iter, double eps,       \
                                  double* shifts, void (*mv)(T*, T*, void*), void* extra, bool worst_first,         \
                                  inversion_verbose_struct* verb) {                                                 \
    MultiDev<T>::fn dev = &minv_vector_cr_m_dev;                                                                    \
    return multi_host<T>("CR-M", dev, phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra, worst_first, \
                         verb);                                                                                     \
  }                                                                                                                 \
  inversion_info minv_vector_bicgstab_m(T** phi, T* phi0, int n_shift, int size, int rfc, int max_iter, double eps, \
                                        double* shifts, void (*mv)(T*, T*, void*), void* extra, bool worst_first,   \
                                        inversion_verbose_struct* verb) {                                           \
    MultiDev<T>::fn dev = &minv_vector_bicgstab_m_dev;                                                              \
    return multi_host<T>("BICGSTAB-M", dev, phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra,        \
                         worst_first, verb);                                                                        \
  }
GLB200_HOST_MULTI(double)
GLB200_HOST_MULTI(zcplx)

// ------------------------------------------------------------------------------------------ preconditioned family
// A host preconditioner callback is mapped to its device form: the stock ones of generic_precond.h are
// recognised (identity_preconditioner; gcr_preconditioner on a known operator); anything else is an error
// unless the parity shim is on (download -> host callback -> upload around every application).
template <typename T>
struct GcrStruct;
template <>
struct GcrStruct<double> {
  typedef gcr_precond_struct_real type;
};
template <>
struct GcrStruct<zcplx> {
  typedef gcr_precond_struct_complex type;
};
template <typename T>
struct PrecondShim {
  void (*fn)(T*, T*, int, void*, inversion_verbose_struct*);
  void* extra;
  std::vector<T> in, out;
};
template <typename T>
void precond_shim_cb(T* d_lhs, T* d_rhs, int size, void* e, inversion_verbose_struct* verb) {
  PrecondShim<T>* s = (PrecondShim<T>*)e;
  glb_context* ctx = glb200_default_context();
  s->in.resize(size);
  s->out.resize(size);
  GLBX(glb_vec_download(ctx, Traits<T>::dtype, size, s->in.data(), d_rhs));
  GLBX(glb_vec_download(ctx, Traits<T>::dtype, size, s->out.data(), d_lhs));
  s->fn(s->out.data(), s->in.data(), size, s->extra, verb);
  GLBX(glb_vec_upload(ctx, Traits<T>::dtype, size, d_lhs, s->out.data()));
}
// mg_preconditioner exists for complex vectors only
template <typename T>
struct MgPrecond {
  typedef void (*pfn)(T*, T*, int, void*, inversion_verbose_struct*);
  static bool is(pfn) { return false; }
  static pfn dev() { return 0; }
};
template <>
struct MgPrecond<zcplx> {
  typedef void (*pfn)(zcplx*, zcplx*, int, void*, inversion_verbose_struct*);
  static bool is(pfn f) {
    pfn mgp = &mg_preconditioner;
    return f == mgp;
  }
  static pfn dev() { return &mg_preconditioner_dev; }
};
template <typename T>
struct PrecondMap {
  typedef void (*pfn)(T*, T*, int, void*, inversion_verbose_struct*);
  pfn dev;
  void* info;
  OpLease L;
  typename GcrStruct<T>::type gcr;
  PrecondShim<T> shim;
  Shim<T> opshim;  // the stock preconditioner's own operator when that is a user host function (shim enabled)
  glb200_mg_host::Hierarchy* mgh;  // mg_preconditioner: the host hierarchy uploaded once for the whole solve
  ~PrecondMap() { delete mgh; }
  PrecondMap(pfn host, void* host_info, int size) : dev(0), info(0), mgh(0) {
    if (MgPrecond<T>::is(host)) {
      mg_precond_struct_complex* p = (mg_precond_struct_complex*)host_info;
      mgh = new glb200_mg_host::Hierarchy(p->mgstruct, p->normal_eqn_smooth || p->normal_eqn_mg);
      mgh->set_precond(p);
      dev = MgPrecond<T>::dev();
      info = &mgh->pc;
      return;
    }
    pfn ident = &identity_preconditioner;
    pfn gcrp = &gcr_preconditioner;
    pfn mrp = &minres_preconditioner;
    if (host == ident) {
      dev = &identity_preconditioner_dev;
    } else if (host == gcrp || host == mrp) {  // the two structs have the same layout (generic_precond.h:27-70)
      typename GcrStruct<T>::type* g = (typename GcrStruct<T>::type*)host_info;
      const Builtin kind = classify(g->matrix_vector);
      gcr.n_step = g->n_step;
      gcr.rel_res = g->rel_res;
      if (kind != B_NONE) {
        lease(kind, g->matrix_extra_data, &L);
        gcr.matrix_vector = &glb200_apply_dev;
        gcr.matrix_extra_data = L.op;
        if (L.comp) {
          gcr.matrix_vector = CompositeCallback<T>::get();
          gcr.matrix_extra_data = L.comp;
        }
      } else if (g_allow_shim) {
        opshim.fn = g->matrix_vector;
        opshim.extra = g->matrix_extra_data;
        opshim.n = size;
        opshim.in.resize(size);
        opshim.out.resize(size);
        gcr.matrix_vector = &shim_cb<T>;
        gcr.matrix_extra_data = &opshim;
      } else {
        throw Error("gcr_ / minres_preconditioner: its operator callback is not a known device operator");
      }
      dev = (host == gcrp) ? (pfn)&gcr_preconditioner_dev : (pfn)&minres_preconditioner_dev;
      info = &gcr;
    } else if (g_allow_shim) {
      shim.fn = host;
      shim.extra = host_info;
      dev = &precond_shim_cb<T>;
      info = &shim;
    } else {
      throw Error("preconditioner callback is not one of generic_precond.h (identity_, gcr_, minres_preconditioner); "
                  "use the _dev entry points with a device preconditioner");
    }
  }
};

#define GLB200_HOST_PRECOND(T, NAME, DEVNAME, ALG)                                                                  \
  inversion_info NAME(T* phi, T* phi0, int size, int max_iter, double eps, void (*mv)(T*, T*, void*), void* extra,  \
                      void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), void* pci,                         \
                      inversion_verbose_struct* verb) {                                                             \
    return host_solve<T>(ALG, phi, phi0, size, mv, extra, [&](T* dp, T* db, void (*cb)(T*, T*, void*), void* ce) {  \
      PrecondMap<T> pm(pc, pci, size);                                                                              \
      return DEVNAME(dp, db, size, max_iter, eps, cb, ce, pm.dev, pm.info, verb);                                   \
    });                                                                                                             \
  }
#define GLB200_HOST_PRECOND_RESTART(T, NAME, DEVNAME, ALG)                                                          \
  inversion_info NAME(T* phi, T* phi0, int size, int max_iter, double res, int rf, void (*mv)(T*, T*, void*),       \
                      void* extra, void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), void* pci,            \
                      inversion_verbose_struct* verb) {                                                             \
    return host_solve<T>(ALG, phi, phi0, size, mv, extra, [&](T* dp, T* db, void (*cb)(T*, T*, void*), void* ce) {  \
      PrecondMap<T> pm(pc, pci, size);                                                                              \
      return DEVNAME(dp, db, size, max_iter, res, rf, cb, ce, pm.dev, pm.info, verb);                               \
    });                                                                                                             \
  }
#define GLB200_HOST_PRECOND_ALL(T)                                                                                  \
  GLB200_HOST_PRECOND(T, minv_vector_cg_precond, minv_vector_cg_precond_dev, "PCG")                                 \
  GLB200_HOST_PRECOND(T, minv_vector_cg_flex_precond, minv_vector_cg_flex_precond_dev, "FPCG")                      \
  GLB200_HOST_PRECOND_RESTART(T, minv_vector_cg_flex_precond_restart, minv_vector_cg_flex_precond_restart_dev, "FPCG") \
  GLB200_HOST_PRECOND(T, minv_vector_gcr_var_precond, minv_vector_gcr_var_precond_dev, "VPGCR")                     \
  GLB200_HOST_PRECOND_RESTART(T, minv_vector_gcr_var_precond_restart, minv_vector_gcr_var_precond_restart_dev, "VPGCR") \
  GLB200_HOST_PRECOND(T, minv_vector_bicgstab_precond, minv_vector_bicgstab_precond_dev, "Preconditioned BiCGStab") \
  GLB200_HOST_PRECOND_RESTART(T, minv_vector_bicgstab_precond_restart, minv_vector_bicgstab_precond_restart_dev,    \
                              "Preconditioned BiCGStab")
GLB200_HOST_PRECOND_ALL(double)
GLB200_HOST_PRECOND_ALL(zcplx)

// the stock preconditioners called directly with host vectors (generic_precond.cpp:23-77)
void identity_preconditioner(double* lhs, double* rhs, int size, void*, inversion_verbose_struct*) {
  std::memcpy(lhs, rhs, sizeof(double) * (size_t)size);
}
void identity_preconditioner(zcplx* lhs, zcplx* rhs, int size, void*, inversion_verbose_struct*) {
  std::memcpy((void*)lhs, (const void*)rhs, sizeof(zcplx) * (size_t)size);
}
void gcr_preconditioner(double* lhs, double* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  gcr_precond_struct_real* g = (gcr_precond_struct_real*)extra_data;
  minv_vector_gcr(lhs, rhs, size, g->n_step, g->rel_res, g->matrix_vector, g->matrix_extra_data, verb);
}
void gcr_preconditioner(zcplx* lhs, zcplx* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  gcr_precond_struct_complex* g = (gcr_precond_struct_complex*)extra_data;
  minv_vector_gcr(lhs, rhs, size, g->n_step, g->rel_res, g->matrix_vector, g->matrix_extra_data, verb);
}

void minres_preconditioner(double* lhs, double* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  minres_precond_struct_real* m = (minres_precond_struct_real*)extra_data;
  minv_vector_minres(lhs, rhs, size, m->n_step, m->rel_res, m->matrix_vector, m->matrix_extra_data, verb);
}
void minres_preconditioner(zcplx* lhs, zcplx* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  minres_precond_struct_complex* m = (minres_precond_struct_complex*)extra_data;
  minv_vector_minres(lhs, rhs, size, m->n_step, m->rel_res, m->matrix_vector, m->matrix_extra_data, verb);
}

// generic_inverter_precond.cpp:18-114
template <typename T>
static inversion_info dispatch_precond(T* lhs, T* rhs, int size, minv_inverter_precond type,
                                       minv_inverter_precond_params& p, void (*mv)(T*, T*, void*), void* extra,
                                       void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), void* pci,
                                       inversion_verbose_struct* verb) {
  switch (type) {
    case MINV_PRE_CG:
      return minv_vector_cg_precond(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    case MINV_PRE_FPCG:
      return p.restart ? minv_vector_cg_flex_precond_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra,
                                                             pc, pci, verb)
                       : minv_vector_cg_flex_precond(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    case MINV_PRE_VPGCR:
      return p.restart ? minv_vector_gcr_var_precond_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra,
                                                             pc, pci, verb)
                       : minv_vector_gcr_var_precond(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    case MINV_PRE_BICGSTAB:
      return p.restart ? minv_vector_bicgstab_precond_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv,
                                                              extra, pc, pci, verb)
                       : minv_vector_bicgstab_precond(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    default:
      return inversion_info();
  }
}
inversion_info minv_preconditioned(double* lhs, double* rhs, int size, minv_inverter_precond type,
                                   minv_inverter_precond_params& p, void (*mv)(double*, double*, void*), void* extra,
                                   void (*pc)(double*, double*, int, void*, inversion_verbose_struct*), void* pci,
                                   inversion_verbose_struct* verb) {
  return dispatch_precond<double>(lhs, rhs, size, type, p, mv, extra, pc, pci, verb);
}
inversion_info minv_preconditioned(zcplx* lhs, zcplx* rhs, int size, minv_inverter_precond type,
                                   minv_inverter_precond_params& p, void (*mv)(zcplx*, zcplx*, void*), void* extra,
                                   void (*pc)(zcplx*, zcplx*, int, void*, inversion_verbose_struct*), void* pci,
                                   inversion_verbose_struct* verb) {
  return dispatch_precond<zcplx>(lhs, rhs, size, type, p, mv, extra, pc, pci, verb);
}
inversion_info minv_vector_cg_m(double** phi, double* phi0, int n_shift, int size, int rfc, int max_iter, double eps,
                                double* shifts, void (*mv)(double*, double*, void*), void* extra, bool worst_first,
                                inversion_verbose_struct* verb) {
  return cg_m_host<double>(phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra, worst_first, verb);
}
inversion_info minv_vector_cg_m(zcplx** phi, zcplx* phi0, int n_shift, int size, int rfc, int max_iter, double eps,
                                double* shifts, void (*mv)(zcplx*, zcplx*, void*), void* extra, bool worst_first,
                                inversion_verbose_struct* verb) {
  return cg_m_host<zcplx>(phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra, worst_first, verb);
}

// generic_inverter.cpp:18-190 : enum dispatch
template <typename T>
static inversion_info dispatch(T* lhs, T* rhs, int size, minv_inverter type, minv_inverter_params& p,
                               void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {
  switch (type) {
    case MINV_CG:
      return p.restart ? minv_vector_cg_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_cg(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_CR:
      return p.restart ? minv_vector_cr_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_cr(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_GCR:
      return p.restart ? minv_vector_gcr_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_gcr(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_BICGSTAB:
      return p.restart
                 ? minv_vector_bicgstab_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                 : minv_vector_bicgstab(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_BICGSTAB_L:
      return p.restart ? minv_vector_bicgstab_l_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq,
                                                        p.bicgstabl_l, mv, extra, verb)
                       : minv_vector_bicgstab_l(lhs, rhs, size, p.max_iters, p.tol, p.bicgstabl_l, mv, extra, verb);
    case MINV_GMRES:
      return p.restart ? minv_vector_gmres_restart(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_gmres(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_SOR:  // generic_inverter.cpp:86-95: no restarts for these two
      return minv_vector_sor(lhs, rhs, size, p.max_iters, p.tol, p.sor_omega, mv, extra, verb);
    case MINV_MINRES:
      return minv_vector_minres(lhs, rhs, size, p.max_iters, p.tol, p.minres_omega, mv, extra, verb);
    default:
      return inversion_info();
  }
}
inversion_info minv_unpreconditioned(double* lhs, double* rhs, int size, minv_inverter type, minv_inverter_params& p,
                                     void (*mv)(double*, double*, void*), void* extra,
                                     inversion_verbose_struct* verb) {
  return dispatch<double>(lhs, rhs, size, type, p, mv, extra, verb);
}
inversion_info minv_unpreconditioned(zcplx* lhs, zcplx* rhs, int size, minv_inverter type, minv_inverter_params& p,
                                     void (*mv)(zcplx*, zcplx*, void*), void* extra, inversion_verbose_struct* verb) {
  return dispatch<zcplx>(lhs, rhs, size, type, p, mv, extra, verb);
}
