// capi_solvers.cpp -- plain-C view of the C++ drop-in API, for harnesses that cannot speak C++
// (tests/ and bench.py load it through ctypes).  Every function here only forwards to the public
// C++ entry points of generic_inverters.h / glb200_device.h / operators.h -- the calls a C++ user
// of the reference would make -- and flattens inversion_info into a POD.
#include <chrono>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "coarse_stencil.h"
#include "dev_internal.hpp"
#include "generic_inverters_precond.h"
#include "mg_complex.h"
#include "null_gen.h"
#include "operators.h"
#include "operators_stencil.h"
#include "u1_utils.h"

typedef std::complex<double> zc;

extern "C" {

typedef struct glbx_result {
  double resSq;
  int iter;
  int success;
  int ops_count;
  int n_rhs;
  double resSqmrhs[32];
  char name[64];
} glbx_result;

// operator selector (same numbering as oracle/oracle_api.h so the tests can share tables)
enum {
  GX_LAPLACE_REAL = 0, GX_LAPLACE_IMAG = 1, GX_LAPLACE_NC = 2, GX_LAPLACE_U1 = 3, GX_STAG_FREE = 4, GX_STAG_U1 = 5,
  GX_STAG_GAMMA5_U1 = 6, GX_STAG_DAGGER_U1 = 7, GX_STAG_NORMAL_U1 = 8, GX_GAMMA5 = 9, GX_STENCIL = 10,
  GX_STENCIL_FROM_STAG = 11, GX_STAG_GAMMA5_FREE = 12, GX_LAPLACE_REAL_NC = 13, GX_STAG_FREE_REAL = 14,
  GX_STAG_DEO_U1 = 15, GX_STAG_DOE_U1 = 16, GX_STAG_M2MDEODOE_U1 = 17, GX_SYMMSHIFT_X = 18, GX_SYMMSHIFT_Y = 19,
  GX_STAG_2LINK_U1 = 20, GX_STAG_INDEX = 21
};
enum {
  GX_CG = 0, GX_CG_RESTART = 1, GX_CR = 2, GX_CR_RESTART = 3, GX_GCR = 4, GX_GCR_RESTART = 5, GX_BICGSTAB = 6,
  GX_BICGSTAB_RESTART = 7, GX_BICGSTAB_L = 8, GX_BICGSTAB_L_RESTART = 9, GX_GMRES = 10, GX_GMRES_RESTART = 11
};

typedef struct glbx_opdesc {
  int kind;
  int X, Y, Nc;
  double mass;
  const void* links;
  const void* clover;
  const void* hopping;
  const void* two_link;
  int has_two;
  double shift[2], eo_shift[2], dof_shift[2];
  int view;  // stencil kinds: 0 apply_stencil_2d, GLB_SV_* the composite callback on the stencil
  double wilson_coeff;  // staggered_u1_op::wilson_coeff (square_staggered_2linklaplace_u1)
} glbx_opdesc;

}  // extern "C"

namespace {

void flatten(const inversion_info& inf, glbx_result* out) {
  std::memset(out, 0, sizeof(*out));
  out->resSq = inf.resSq;
  out->iter = inf.iter;
  out->success = inf.success ? 1 : 0;
  out->ops_count = inf.ops_count;
  out->n_rhs = inf.n_rhs;
  if (inf.resSqmrhs && inf.n_rhs > 0)
    for (int i = 0; i < inf.n_rhs && i < 32; i++) out->resSqmrhs[i] = inf.resSqmrhs[i];
  std::strncpy(out->name, inf.name.c_str(), sizeof(out->name) - 1);
}

void make_verb(int level, inversion_verbose_struct* v) {
  v->verbosity = (inversion_verbose_level)level;
  v->verb_prefix = "[glb200] ";
  v->precond_verbosity = VERB_NONE;
  v->precond_verb_prefix = "";
}

// the reference-side objects a user would hold for each operator
struct HostOp {
  staggered_u1_op stag;
  laplace_op lap;
  Lattice* lat;
  stencil_2d* st;
  void (*cz)(zc*, zc*, void*);
  void (*cd)(double*, double*, void*);
  void* extra;
  int size;
  HostOp() : lat(0), st(0), cz(0), cd(0), extra(0), size(0) {}
  ~HostOp() {
    delete st;
    delete lat;
  }
};

bool build_host_op(const glbx_opdesc* d, HostOp* h) {
  h->stag.lattice = (zc*)d->links;
  h->stag.mass = d->mass;
  h->stag.x_fine = d->X;
  h->stag.y_fine = d->Y;
  h->stag.Nc = d->Nc > 0 ? d->Nc : 1;
  h->stag.wilson_coeff = d->wilson_coeff;
  h->lap.N = d->X;
  h->lap.mass_sq = d->mass;
  h->extra = &h->stag;
  h->size = d->X * d->Y;
  switch (d->kind) {
    case GX_LAPLACE_REAL: h->cd = &square_laplacian; h->extra = &h->lap; break;
    case GX_LAPLACE_IMAG: h->cz = &square_laplacian; h->extra = &h->lap; break;
    case GX_LAPLACE_NC: h->cz = &square_laplace; h->size *= h->stag.Nc; break;
    case GX_LAPLACE_REAL_NC: h->cd = &square_laplace; h->size *= h->stag.Nc; break;
    case GX_STAG_FREE_REAL: h->cd = &square_staggered; break;
    case GX_LAPLACE_U1: h->cz = &square_laplace_u1; break;
    case GX_STAG_FREE: h->cz = &square_staggered; break;
    case GX_STAG_U1: h->cz = &square_staggered_u1; break;
    case GX_STAG_GAMMA5_U1: h->cz = &square_staggered_gamma5_u1; break;
    case GX_STAG_GAMMA5_FREE: h->cz = &square_staggered_gamma5; break;
    case GX_STAG_DAGGER_U1: h->cz = &square_staggered_dagger_u1; break;
    case GX_STAG_NORMAL_U1: h->cz = &square_staggered_normal_u1; break;
    case GX_GAMMA5: h->cz = &gamma_5; break;
    case GX_STAG_DEO_U1: h->cz = &square_staggered_deo_u1; break;
    case GX_STAG_DOE_U1: h->cz = &square_staggered_doe_u1; break;
    case GX_STAG_M2MDEODOE_U1: h->cz = &square_staggered_m2mdeodoe_u1; break;
    case GX_SYMMSHIFT_X: h->cz = &staggered_symmshift_x; break;
    case GX_SYMMSHIFT_Y: h->cz = &staggered_symmshift_y; break;
    case GX_STAG_2LINK_U1: h->cz = &square_staggered_2linklaplace_u1; break;
    case GX_STAG_INDEX: h->cz = &staggered_index_operator; break;
    case GX_STENCIL:
    case GX_STENCIL_FROM_STAG: {
      int dims[2] = {d->X, d->Y};
      const int nc = (d->kind == GX_STENCIL_FROM_STAG) ? 1 : h->stag.Nc;
      h->lat = new Lattice(2, dims, nc);
      if (d->kind == GX_STENCIL_FROM_STAG) {
        h->st = new stencil_2d(h->lat, 1);
        get_square_staggered_u1_stencil(h->st, &h->stag);
      } else {
        h->st = new stencil_2d(h->lat, d->has_two ? 2 : 1, zc(d->shift[0], d->shift[1]),
                               zc(d->eo_shift[0], d->eo_shift[1]), zc(d->dof_shift[0], d->dof_shift[1]));
        const size_t m = (size_t)d->X * d->Y * nc * nc;
        std::memcpy((void*)h->st->clover, d->clover, m * sizeof(zc));
        std::memcpy((void*)h->st->hopping, d->hopping, 4 * m * sizeof(zc));
        if (d->has_two) std::memcpy((void*)h->st->two_link, d->two_link, 8 * m * sizeof(zc));
        h->st->generated = true;
      }
      switch (d->view) {
        case 0: h->cz = &apply_stencil_2d; break;
        case GLB_SV_M2MDEODOE: h->cz = &apply_square_staggered_m2mdeodoe_stencil; break;
        case GLB_SV_M2MDTBDBT: h->cz = &apply_square_staggered_m2mdtbdbt_stencil; break;
        case GLB_SV_NORMAL_EO: h->cz = &apply_square_staggered_normal_eo_stencil; break;
        case GLB_SV_NORMAL_TB: h->cz = &apply_square_staggered_normal_tb_stencil; break;
        case GLB_SV_DAGGER_EO: h->cz = &apply_square_staggered_dagger_eo_stencil; break;
        case GLB_SV_DAGGER_TB: h->cz = &apply_square_staggered_dagger_tb_stencil; break;
        default: return false;
      }
      h->extra = h->st;
      h->size = d->X * d->Y * nc;
      break;
    }
    default: return false;
  }
  return true;
}

template <typename T>
inversion_info run_host(int solver, T* phi, T* b, int size, int max_iter, double eps, int rf, int l,
                        void (*cb)(T*, T*, void*), void* extra, inversion_verbose_struct* v) {
  switch (solver) {
    case GX_CG: return minv_vector_cg(phi, b, size, max_iter, eps, cb, extra, v);
    case GX_CG_RESTART: return minv_vector_cg_restart(phi, b, size, max_iter, eps, rf, cb, extra, v);
    case GX_CR: return minv_vector_cr(phi, b, size, max_iter, eps, cb, extra, v);
    case GX_CR_RESTART: return minv_vector_cr_restart(phi, b, size, max_iter, eps, rf, cb, extra, v);
    case GX_GCR: return minv_vector_gcr(phi, b, size, max_iter, eps, cb, extra, v);
    case GX_GCR_RESTART: return minv_vector_gcr_restart(phi, b, size, max_iter, eps, rf, cb, extra, v);
    case GX_BICGSTAB: return minv_vector_bicgstab(phi, b, size, max_iter, eps, cb, extra, v);
    case GX_BICGSTAB_RESTART: return minv_vector_bicgstab_restart(phi, b, size, max_iter, eps, rf, cb, extra, v);
    case GX_BICGSTAB_L: return minv_vector_bicgstab_l(phi, b, size, max_iter, eps, l, cb, extra, v);
    case GX_BICGSTAB_L_RESTART: return minv_vector_bicgstab_l_restart(phi, b, size, max_iter, eps, rf, l, cb, extra, v);
    case GX_GMRES: return minv_vector_gmres(phi, b, size, max_iter, eps, cb, extra, v);
    case GX_GMRES_RESTART: return minv_vector_gmres_restart(phi, b, size, max_iter, eps, rf, cb, extra, v);
  }
  return inversion_info();
}

template <typename T>
inversion_info run_dev(int solver, T* phi, T* b, int size, int max_iter, double eps, int rf, int l, void* op,
                       inversion_verbose_struct* v) {
  void (*cb)(T*, T*, void*) = &glb200_apply_dev;
  switch (solver) {
    case GX_CG: return minv_vector_cg_dev(phi, b, size, max_iter, eps, cb, op, v);
    case GX_CG_RESTART: return minv_vector_cg_restart_dev(phi, b, size, max_iter, eps, rf, cb, op, v);
    case GX_CR: return minv_vector_cr_dev(phi, b, size, max_iter, eps, cb, op, v);
    case GX_CR_RESTART: return minv_vector_cr_restart_dev(phi, b, size, max_iter, eps, rf, cb, op, v);
    case GX_GCR: return minv_vector_gcr_dev(phi, b, size, max_iter, eps, cb, op, v);
    case GX_GCR_RESTART: return minv_vector_gcr_restart_dev(phi, b, size, max_iter, eps, rf, cb, op, v);
    case GX_BICGSTAB: return minv_vector_bicgstab_dev(phi, b, size, max_iter, eps, cb, op, v);
    case GX_BICGSTAB_RESTART: return minv_vector_bicgstab_restart_dev(phi, b, size, max_iter, eps, rf, cb, op, v);
    case GX_BICGSTAB_L: return minv_vector_bicgstab_l_dev(phi, b, size, max_iter, eps, l, cb, op, v);
    case GX_BICGSTAB_L_RESTART: return minv_vector_bicgstab_l_restart_dev(phi, b, size, max_iter, eps, rf, l, cb, op, v);
    case GX_GMRES: return minv_vector_gmres_dev(phi, b, size, max_iter, eps, cb, op, v);
    case GX_GMRES_RESTART: return minv_vector_gmres_restart_dev(phi, b, size, max_iter, eps, rf, cb, op, v);
  }
  return inversion_info();
}

}  // namespace

extern "C" {

void glb200_cache_operators(int on);

// the process-wide context of the C++ layer (so a harness can allocate device vectors on it)
glb_context* glbx_default_context(void) {
  try {
    return glb200_default_context();
  } catch (const std::exception& e) {
    std::cerr << "[glb200] " << e.what() << std::endl;
    return 0;
  }
}
void glbx_set_default_context(glb_context* ctx) { glb200_set_default_context(ctx); }
void glbx_force_host_scalars(int on) { glb200_force_host_scalars(on != 0); }
void glbx_allow_host_callback_shim(int on) { glb200_allow_host_callback_shim(on != 0); }
void glbx_cache_operators(int on) { glb200_cache_operators(on); }

// The synthetic inputs BASELINE.md section 3 prescribes, drawn with the drop-in's own host helpers the way the
// reference's drivers do it (one std::mt19937(seed) stream: gauss_gauge_u1(beta) of u1_utils.h, then gaussian() of
// generic_vector.h).  links: 2*X*Y complex (lattice[y*X*2 + x*2 + mu]), rhs: X*Y complex.  Host-side input
// preparation for bench.py and the examples; not on the solver path.
int glbx_synthetic_inputs(unsigned seed, int X, int Y, double beta, void* links, void* rhs) {
  if (X < 1 || Y < 1 || !links) return GLB_ERR_ARG;
  std::mt19937 gen(seed);
  gauss_gauge_u1((zc*)links, X, Y, gen, beta);
  if (rhs) {
    // gaussian() of generic_vector.h:48-60 (the library itself never includes that driver-side header): unit normal
    // entries, the imaginary part drawn first
    std::normal_distribution<> unit_normal(0.0, 1.0);
    zc* v = (zc*)rhs;
    for (long long i = 0; i < (long long)X * Y; i++) {
      const double im = unit_normal(gen);
      const double re = unit_normal(gen);
      v[i] = zc(re, im);
    }
  }
  return GLB_OK;
}

// lhs = A rhs through the reference-named host callback (upload, device apply, download)
int glbx_host_apply(const glbx_opdesc* d, void* lhs, const void* rhs) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  if (h.cz)
    h.cz((zc*)lhs, (zc*)rhs, h.extra);
  else
    h.cd((double*)lhs, (double*)rhs, h.extra);
  return GLB_OK;
}

// apply_stencil_2d_{eo,oe,tb,bt} through the reference-named host functions (d: a stencil descriptor); part 1..4
int glbx_host_stencil_part(const glbx_opdesc* d, int part, void* lhs, const void* rhs) {
  HostOp h;
  if (!build_host_op(d, &h) || !h.st) return GLB_ERR_ARG;
  if (part == GLB_PART_EO) apply_stencil_2d_eo((zc*)lhs, (zc*)rhs, h.st);
  else if (part == GLB_PART_OE) apply_stencil_2d_oe((zc*)lhs, (zc*)rhs, h.st);
  else if (part == GLB_PART_TB) apply_stencil_2d_tb((zc*)lhs, (zc*)rhs, h.st);
  else if (part == GLB_PART_BT) apply_stencil_2d_bt((zc*)lhs, (zc*)rhs, h.st);
  else return GLB_ERR_ARG;
  return GLB_OK;
}

// operators.cpp:528 / :574 through the reference-named host functions (d: any gauged staggered descriptor)
int glbx_host_eoprec_prepare(const glbx_opdesc* d, void* rhs_e, const void* rhs_orig) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  square_staggered_eoprec_prepare((zc*)rhs_e, (zc*)rhs_orig, &h.stag);
  return GLB_OK;
}
int glbx_host_eoprec_reconstruct(const glbx_opdesc* d, void* lhs_full, const void* lhs_e, const void* rhs_o) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  square_staggered_eoprec_reconstruct((zc*)lhs_full, (zc*)lhs_e, (zc*)rhs_o, &h.stag);
  return GLB_OK;
}

// the reference's call: solver(phi, phi0, size, ..., callback, extra_info, verbosity) on HOST vectors
int glbx_host_solve(int solver, const glbx_opdesc* d, void* phi, const void* phi0, int max_iter, double eps,
                    int restart_freq, int l, int verbosity, glbx_result* out) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  {
    // slab form of the reference call: with a communicator attached to the default context every rank passes ITS rows
    // of phi / phi0 (size = X * Yloc * Nc) together with the global operator description, of which it reads its rows
    glb_context* ctx = glb200_default_context();
    if (glb_comm_size(ctx) > 1) {
      int y0 = 0, yl = 0;
      if (glb_slab_bounds(ctx, d->Y, &y0, &yl) != GLB_OK) return GLB_ERR_ARG;
      h.size = (int)((long long)h.size / d->Y * yl);
    }
  }
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  inversion_info inf;
  if (h.cz)
    inf = run_host<zc>(solver, (zc*)phi, (zc*)phi0, h.size, max_iter, eps, restart_freq, l, h.cz, h.extra, &v);
  else
    inf = run_host<double>(solver, (double*)phi, (double*)phi0, h.size, max_iter, eps, restart_freq, l, h.cd, h.extra,
                           &v);
  flatten(inf, out);
  return GLB_OK;
}

int glbx_host_solve_cg_m(const glbx_opdesc* d, void** phi, const void* phi0, int n_shift, int resid_freq_check,
                         int max_iter, double eps, double* shifts, int worst_first, int verbosity, glbx_result* out) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  if (h.cz) {
    inversion_info inf = minv_vector_cg_m((zc**)phi, (zc*)phi0, n_shift, h.size, resid_freq_check, max_iter, eps,
                                          shifts, h.cz, h.extra, worst_first != 0, &v);
    flatten(inf, out);
  } else {
    inversion_info inf = minv_vector_cg_m((double**)phi, (double*)phi0, n_shift, h.size, resid_freq_check, max_iter,
                                          eps, shifts, h.cd, h.extra, worst_first != 0, &v);
    flatten(inf, out);
  }
  return GLB_OK;
}

// apply_square_staggered_{eo,tb}prec_{prepare,reconstruct}_stencil with host vectors (operators_stencil.h:31-38)
int glbx_host_stencil_prec(const glbx_opdesc* d, int top_bottom, int reconstruct, void* out, void* a, void* b) {
  HostOp h;
  if (!build_host_op(d, &h) || !h.st) return GLB_ERR_ARG;
  if (!reconstruct) {
    if (top_bottom)
      apply_square_staggered_tbprec_prepare_stencil((zc*)out, (zc*)a, h.st);
    else
      apply_square_staggered_eoprec_prepare_stencil((zc*)out, (zc*)a, h.st);
  } else {
    if (top_bottom)
      apply_square_staggered_tbprec_reconstruct_stencil((zc*)out, (zc*)a, (zc*)b, h.st);
    else
      apply_square_staggered_eoprec_reconstruct_stencil((zc*)out, (zc*)a, (zc*)b, h.st);
  }
  return GLB_OK;
}

// minv_vector_sor (which = 0) / minv_vector_minres (which = 1) with their relaxation parameter, host vectors
int glbx_host_solve_relax(int which, const glbx_opdesc* d, void* phi, const void* phi0, int max_iter, double eps,
                          double omega, int verbosity, glbx_result* out) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  inversion_info inf;
  if (h.cz)
    inf = which == 0 ? minv_vector_sor((zc*)phi, (zc*)phi0, h.size, max_iter, eps, omega, h.cz, h.extra, &v)
                     : minv_vector_minres((zc*)phi, (zc*)phi0, h.size, max_iter, eps, omega, h.cz, h.extra, &v);
  else
    inf = which == 0 ? minv_vector_sor((double*)phi, (double*)phi0, h.size, max_iter, eps, omega, h.cd, h.extra, &v)
                     : minv_vector_minres((double*)phi, (double*)phi0, h.size, max_iter, eps, omega, h.cd, h.extra, &v);
  flatten(inf, out);
  return GLB_OK;
}
// the same on device vectors
int glbx_dev_solve_relax(int which, glb_operator* op, void* d_phi, void* d_phi0, int max_iter, double eps, double omega,
                         int verbosity, glbx_result* out) {
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  const int size = (int)glb_op_local_size(op);
  inversion_info inf;
  if (glb_op_dtype(op) == GLB_COMPLEX) {
    void (*cb)(zc*, zc*, void*) = &glb200_apply_dev;
    inf = which == 0 ? minv_vector_sor_dev((zc*)d_phi, (zc*)d_phi0, size, max_iter, eps, omega, cb, op, &v)
                     : minv_vector_minres_dev((zc*)d_phi, (zc*)d_phi0, size, max_iter, eps, omega, cb, op, &v);
  } else {
    void (*cb)(double*, double*, void*) = &glb200_apply_dev;
    inf = which == 0 ? minv_vector_sor_dev((double*)d_phi, (double*)d_phi0, size, max_iter, eps, omega, cb, op, &v)
                     : minv_vector_minres_dev((double*)d_phi, (double*)d_phi0, size, max_iter, eps, omega, cb, op, &v);
  }
  flatten(inf, out);
  return GLB_OK;
}

// device vectors + a glb_operator handle (the device variant of the callback contract)
int glbx_dev_solve(int solver, glb_operator* op, void* d_phi, void* d_phi0, int max_iter, double eps, int restart_freq,
                   int l, int verbosity, glbx_result* out) {
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  const int size = (int)glb_op_local_size(op);
  inversion_info inf;
  if (glb_op_dtype(op) == GLB_COMPLEX)
    inf = run_dev<zc>(solver, (zc*)d_phi, (zc*)d_phi0, size, max_iter, eps, restart_freq, l, op, &v);
  else
    inf = run_dev<double>(solver, (double*)d_phi, (double*)d_phi0, size, max_iter, eps, restart_freq, l, op, &v);
  flatten(inf, out);
  return GLB_OK;
}

int glbx_dev_solve_cg_m(glb_operator* op, void** d_phi, void* d_phi0, int n_shift, int resid_freq_check, int max_iter,
                        double eps, double* shifts, int worst_first, int verbosity, glbx_result* out) {
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  const int size = (int)glb_op_local_size(op);
  if (glb_op_dtype(op) == GLB_COMPLEX) {
    void (*cb)(zc*, zc*, void*) = &glb200_apply_dev;
    inversion_info inf = minv_vector_cg_m_dev((zc**)d_phi, (zc*)d_phi0, n_shift, size, resid_freq_check, max_iter, eps,
                                              shifts, cb, (void*)op, worst_first != 0, &v);
    flatten(inf, out);
  } else {
    void (*cb)(double*, double*, void*) = &glb200_apply_dev;
    inversion_info inf = minv_vector_cg_m_dev((double**)d_phi, (double*)d_phi0, n_shift, size, resid_freq_check,
                                              max_iter, eps, shifts, cb, (void*)op, worst_first != 0, &v);
    flatten(inf, out);
  }
  return GLB_OK;
}


// ---- multishift CR / BiCGStab and the preconditioned family through the reference's own calls
// which: 0 minv_vector_cg_m, 1 minv_vector_cr_m, 2 minv_vector_bicgstab_m   (HOST vectors)
int glbx_host_solve_multi(int which, const glbx_opdesc* d, void** phi, const void* phi0, int n_shift, int resid_freq_check,
                          int max_iter, double eps, double* shifts, int worst_first, int verbosity, glbx_result* out) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  inversion_info inf(n_shift);
  if (h.cz) {
    zc** p = (zc**)phi;
    zc* b = (zc*)phi0;
    if (which == 0) inf = minv_vector_cg_m(p, b, n_shift, h.size, resid_freq_check, max_iter, eps, shifts, h.cz, h.extra, worst_first != 0, &v);
    if (which == 1) inf = minv_vector_cr_m(p, b, n_shift, h.size, resid_freq_check, max_iter, eps, shifts, h.cz, h.extra, worst_first != 0, &v);
    if (which == 2) inf = minv_vector_bicgstab_m(p, b, n_shift, h.size, resid_freq_check, max_iter, eps, shifts, h.cz, h.extra, worst_first != 0, &v);
  } else {
    double** p = (double**)phi;
    double* b = (double*)phi0;
    if (which == 0) inf = minv_vector_cg_m(p, b, n_shift, h.size, resid_freq_check, max_iter, eps, shifts, h.cd, h.extra, worst_first != 0, &v);
    if (which == 1) inf = minv_vector_cr_m(p, b, n_shift, h.size, resid_freq_check, max_iter, eps, shifts, h.cd, h.extra, worst_first != 0, &v);
    if (which == 2) inf = minv_vector_bicgstab_m(p, b, n_shift, h.size, resid_freq_check, max_iter, eps, shifts, h.cd, h.extra, worst_first != 0, &v);
  }
  flatten(inf, out);
  return GLB_OK;
}
}  // extern "C"

namespace {
// solver: 0 PCG, 1 FPCG, 2 FPCG restart, 3 VPGCR, 4 VPGCR restart, 5 PBiCGStab, 6 PBiCGStab restart
// precond: 0 identity_preconditioner, 1 gcr_preconditioner (n_step iterations to rel_res on the same operator)
template <typename T, typename G>
inversion_info run_host_precond(int solver, T* phi, T* b, int size, int max_iter, double eps, int rf,
                                void (*cb)(T*, T*, void*), void* extra, int precond, int n_step, double rel_res,
                                inversion_verbose_struct* v) {
  G g;
  g.n_step = n_step;
  g.rel_res = rel_res;
  g.matrix_vector = cb;
  g.matrix_extra_data = extra;
  typedef void (*pfn)(T*, T*, int, void*, inversion_verbose_struct*);
  // precond 2: minres_preconditioner; its struct has the layout of the GCR one (generic_precond.h:27-70)
  pfn pc_gcr = &gcr_preconditioner, pc_id = &identity_preconditioner, pc_mr = &minres_preconditioner;
  pfn pc = (precond == 1) ? pc_gcr : (precond == 2) ? pc_mr : pc_id;
  void* pci = (precond != 0) ? (void*)&g : 0;
  switch (solver) {
    case 0: return minv_vector_cg_precond(phi, b, size, max_iter, eps, cb, extra, pc, pci, v);
    case 1: return minv_vector_cg_flex_precond(phi, b, size, max_iter, eps, cb, extra, pc, pci, v);
    case 2: return minv_vector_cg_flex_precond_restart(phi, b, size, max_iter, eps, rf, cb, extra, pc, pci, v);
    case 3: return minv_vector_gcr_var_precond(phi, b, size, max_iter, eps, cb, extra, pc, pci, v);
    case 4: return minv_vector_gcr_var_precond_restart(phi, b, size, max_iter, eps, rf, cb, extra, pc, pci, v);
    case 5: return minv_vector_bicgstab_precond(phi, b, size, max_iter, eps, cb, extra, pc, pci, v);
    case 6: return minv_vector_bicgstab_precond_restart(phi, b, size, max_iter, eps, rf, cb, extra, pc, pci, v);
  }
  return inversion_info();
}
}  // namespace

extern "C" {
int glbx_host_solve_precond(int solver, const glbx_opdesc* d, void* phi, const void* phi0, int max_iter, double eps,
                            int restart_freq, int precond, int n_step, double rel_res, int verbosity, glbx_result* out) {
  HostOp h;
  if (!build_host_op(d, &h)) return GLB_ERR_ARG;
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  inversion_info inf;
  if (h.cz)
    inf = run_host_precond<zc, gcr_precond_struct_complex>(solver, (zc*)phi, (zc*)phi0, h.size, max_iter, eps, restart_freq,
                                                          h.cz, h.extra, precond, n_step, rel_res, &v);
  else
    inf = run_host_precond<double, gcr_precond_struct_real>(solver, (double*)phi, (double*)phi0, h.size, max_iter, eps,
                                                           restart_freq, h.cd, h.extra, precond, n_step, rel_res, &v);
  flatten(inf, out);
  return GLB_OK;
}

// ---- multigrid-preconditioned solves (mg_complex.h): the hierarchy is handed over as device objects
typedef struct glbx_mg {
  mg_operator_struct_complex_dev mg;
  mg_precond_struct_complex_dev pc;
  std::vector<glb_operator*> ops;
  std::vector<glb_mg_transfer*> trs;
  std::vector<int> n_pre, n_post;
  std::vector<double> rel_res;
  // set-up on the device (glbx_mg_setup): this handle then owns the coarse operators, the transfers and the
  // device null vectors; the level-0 operator stays the caller's
  bool owns_hierarchy;
  glb_context* ctx;  // of the set-up (the level-0 operator may be destroyed by its owner before this handle)
  std::vector<int> bx, by, nvec;
  std::vector<std::vector<zc*> > null_dev;
  std::vector<zc**> null_tab;
  double setup_seconds[4];  // null vectors, block orthonormalisation, transfers + Galerkin products, total
  // mg(), pc(): every pointer of the two structs starts null
  glbx_mg() : mg(), pc(), owns_hierarchy(false), ctx(0) { setup_seconds[0] = setup_seconds[1] = setup_seconds[2] = setup_seconds[3] = 0.0; }
} glbx_mg;

static void mg_defaults(glbx_mg* h, int n_refine) {
  h->mg.n_refine = n_refine;
  h->mg.stencils = h->ops.data();
  h->mg.transfers = h->trs.data();
  h->mg.curr_level = 0;
  h->mg.dslash_count = new dslash_tracker(n_refine);
  // defaults of multigrid/aa_mg/input_params.cpp:751-800
  h->n_pre.assign(n_refine, 6);
  h->n_post.assign(n_refine, 6);
  h->rel_res.assign(n_refine, 1e-2);
  h->pc.in_smooth_type = MINV_GCR;
  h->pc.omega_smooth = 0.67;
  h->pc.n_pre_smooth = h->n_pre.data();
  h->pc.n_post_smooth = h->n_post.data();
  h->pc.normal_eqn_mg = false;
  h->pc.normal_eqn_smooth = false;
  h->pc.mlevel_type = MLEVEL_SMOOTH;
  h->pc.in_solve_type = GCR;
  h->pc.n_max = 1024;
  h->pc.n_restart = 64;
  h->pc.rel_res = h->rel_res.data();
  h->pc.mgstruct = &h->mg;
  h->pc.quiet = true;
}

glbx_mg* glbx_mg_create(int n_refine, glb_operator** level_ops, glb_mg_transfer** transfers) {
  if (n_refine < 1 || !level_ops || !transfers) return 0;
  glbx_mg* h = new glbx_mg();
  h->ops.assign(level_ops, level_ops + n_refine + 1);
  h->trs.assign(transfers, transfers + n_refine);
  mg_defaults(h, n_refine);
  return h;
}

void glbx_mg_destroy(glbx_mg* h) {
  if (!h) return;
  if (h->owns_hierarchy) {
    glb_context* ctx = h->ctx;
    for (size_t i = 1; i < h->ops.size(); i++)
      if (h->ops[i]) glb_op_destroy(h->ops[i]);
    for (size_t i = 0; i < h->trs.size(); i++)
      if (h->trs[i]) glb_mg_transfer_destroy(h->trs[i]);
    if (h->mg.symmshift_x) glb_op_destroy(h->mg.symmshift_x);
    if (h->mg.symmshift_y) glb_op_destroy(h->mg.symmshift_y);
    for (size_t l = 0; l < h->null_dev.size(); l++)
      for (size_t v = 0; v < h->null_dev[l].size(); v++)
        if (h->null_dev[l][v] && ctx) glb_vec_free(ctx, h->null_dev[l][v]);
  }
  delete h->mg.dslash_count;
  delete h;
}

// The set-up sequence of the reference's driver for --operator staggered --null-operator staggered
// (multigrid/aa_mg/aa_mg_square_staggered_u1.cpp:716-1143) on the device.  `fine` is the level-0 stencil2d operator
// (get_square_staggered_u1_stencil with the mass in the shift, :986-996); it stays the caller's.  Per refinement:
// null_generate_random_smooth_dev with the shift set to null_mass, block_orthonormalize_dev,
// generate_coarse_from_fine_stencil_dev(ignore_shifts = true) and the shift copied down; at the end every level's
// shift is the true mass again (the Galerkin products do not depend on it, so the reference's second "final" build
// :1066-1093 gives the same matrices).
//   nvec[l]   total null vectors of refinement l (after the partition);  bstrat: 0 none, 1 even/odd
//   null_gen  minv_inverter;  tol[l], max_iter[l] per refinement;  seed of the std::mt19937 behind the sources
//   do_free   free-field null vectors (null_generate_free_dev) instead of smoothed random ones
//   links     BLOCK_TOPO (bstrat 3) only: the host gauge field (reference layout) the symmetric shifts are built from
//   null_prec null_precond_strategy: 0 plain solve, 1 even/odd (top/bottom below the top level), 2 normal equations
// On y-slabs X, Y are the GLOBAL extents and `fine` the slab operator; every rank calls with the same arguments.
glbx_mg* glbx_mg_setup(glb_operator* fine, int X, int Y, int n_refine, const int* block, const int* nvec, int bstrat,
                       double null_mass, int null_gen, const double* tol, const int* max_iter, int restart_freq,
                       int bicgstab_l, int do_ortho_eo, int do_global_ortho_conj, unsigned seed, int verbosity,
                       int null_prec, int do_free, const void* links) {
  if (!fine || n_refine < 1 || !block || !nvec || !tol || !max_iter) return 0;
  glbx_mg* h = new glbx_mg();
  double mass_shift[2] = {0.0, 0.0};
  bool shifted = false;
  try {
    glb_context* ctx = glb_op_context(fine);
    h->ctx = ctx;
    h->owns_hierarchy = true;
    h->ops.assign(n_refine + 1, (glb_operator*)0);
    h->ops[0] = fine;
    h->trs.assign(n_refine, (glb_mg_transfer*)0);
    mg_defaults(h, n_refine);
    h->bx.assign(block, block + n_refine);
    h->by.assign(block, block + n_refine);
    h->nvec.assign(nvec, nvec + n_refine);
    h->mg.x_fine = X;
    h->mg.y_fine = Y;
    h->mg.blocksize_x = h->bx.data();
    h->mg.blocksize_y = h->by.data();
    h->mg.n_vectors = h->nvec.data();
    h->null_dev.resize(n_refine);
    h->null_tab.resize(n_refine);
    for (int l = 0; l < n_refine; l++) {  // aa_mg_square_staggered_u1.cpp:607-616: allocated and zeroed
      int lx, ly, ld, ly0, lyloc;
      mg_level_dims(&h->mg, l, &lx, &ly, &ld);
      GLBX(glb_slab_bounds(ctx, ly, &ly0, &lyloc));
      const size_t sz = (size_t)lx * lyloc * ld;  // this rank's rows
      h->null_dev[l].assign(nvec[l], (zc*)0);
      for (int v = 0; v < nvec[l]; v++) {
        void* p = 0;
        GLBX(glb_vec_alloc(ctx, GLB_COMPLEX, sz, &p));
        h->null_dev[l][v] = (zc*)p;
        GLBX(glb_vec_zero(ctx, GLB_COMPLEX, sz, p));
      }
      h->null_tab[l] = h->null_dev[l].data();
    }
    h->mg.null_vectors = h->null_tab.data();
    if (bstrat == BLOCK_TOPO) {  // null_gen.cpp:36-71 needs the symmetric shifts of the gauge field
      if (!links) throw glbx::Error("BLOCK_TOPO needs the gauge links");
      staggered_u1_op st;
      st.lattice = (zc*)links;
      st.mass = 0.0;
      st.x_fine = X;
      st.y_fine = Y;
      st.Nc = 1;
      st.wilson_coeff = 0.0;
      void (*sx)(zc*, zc*, void*) = &staggered_symmshift_x;
      void (*sy)(zc*, zc*, void*) = &staggered_symmshift_y;
      h->mg.symmshift_x = glb200_operator_from_callback(sx, (void*)&st);
      h->mg.symmshift_y = glb200_operator_from_callback(sy, (void*)&st);
      if (!h->mg.symmshift_x || !h->mg.symmshift_y) throw glbx::Error("BLOCK_TOPO: cannot build the symmetric shifts");
    }

    null_vector_params nv;
    nv.null_gen = (minv_inverter)null_gen;
    nv.null_prec = (null_precond_strategy)null_prec;
    nv.null_restart = restart_freq > 0;
    nv.null_restart_freq = restart_freq;
    nv.null_bicgstab_l = bicgstab_l;
    nv.null_mass = null_mass;
    nv.bstrat = (blocking_strategy)bstrat;
    nv.null_partitions = (bstrat == BLOCK_EO || bstrat == BLOCK_TOPO) ? 2 : (bstrat == BLOCK_CORNER) ? 4 : 1;  // :412-427
    nv.do_ortho_eo = do_ortho_eo != 0;
    nv.do_global_ortho_conj = do_global_ortho_conj != 0;
    nv.quiet = verbosity == 0;
    for (int l = 0; l < n_refine; l++) {
      nv.n_null_vectors.push_back(nvec[l] / nv.null_partitions);
      nv.null_precisions.push_back(tol[l]);
      nv.null_max_iters.push_back(max_iter[l]);
    }
    std::mt19937 generator(seed);
    inversion_verbose_struct verb;
    make_verb(verbosity, &verb);

    double null_shift[2] = {null_mass, 0.0};
    GLBX(glb_op_get_shifts(fine, mass_shift, 0, 0));
    GLBX(glb_op_set_shifts(fine, null_shift, 0, 0));  // :757
    shifted = true;
    typedef std::chrono::steady_clock clk;
    const clk::time_point t_all = clk::now();
    for (int n = 0; n < n_refine; n++) {
      verb.verb_prefix = "[L" + std::to_string(h->mg.curr_level + 1) + "_NULLVEC]: ";
      clk::time_point t0 = clk::now();
      if (do_free)  // :792-795
        null_generate_free_dev(&h->mg, &nv, false, 0);
      else
        null_generate_random_smooth_dev(&h->mg, &nv, &verb, &generator);
      GLBX(glb_synchronize(ctx));
      clk::time_point t1 = clk::now();
      block_orthonormalize_dev(&h->mg);
      GLBX(glb_synchronize(ctx));
      clk::time_point t2 = clk::now();
      generate_coarse_from_fine_stencil_dev(&h->mg, true);
      GLBX(glb_op_set_shifts(h->ops[n + 1], null_shift, 0, 0));  // :933
      GLBX(glb_synchronize(ctx));
      clk::time_point t3 = clk::now();
      h->setup_seconds[0] += std::chrono::duration<double>(t1 - t0).count();
      h->setup_seconds[1] += std::chrono::duration<double>(t2 - t1).count();
      h->setup_seconds[2] += std::chrono::duration<double>(t3 - t2).count();
      if (n != n_refine - 1) level_down(&h->mg);
    }
    h->mg.curr_level = 0;
    for (int n = 0; n <= n_refine; n++) GLBX(glb_op_set_shifts(h->ops[n], mass_shift, 0, 0));  // :996, :1086
    h->setup_seconds[3] = std::chrono::duration<double>(clk::now() - t_all).count();
  } catch (const std::exception& e) {
    std::cerr << "glbx_mg_setup: " << e.what() << "\n";
    if (shifted) glb_op_set_shifts(fine, mass_shift, 0, 0);
    glbx_mg_destroy(h);
    return 0;
  }
  return h;
}

// what the set-up produced: the operator of level l (l >= 1: owned by the handle), the transfer of refinement l,
// device null vector v of refinement l, wall-clock seconds {null vectors, orthonormalisation, Galerkin, total}
glb_operator* glbx_mg_level_op(glbx_mg* h, int level) { return (h && level >= 0 && level < (int)h->ops.size()) ? h->ops[level] : 0; }
glb_mg_transfer* glbx_mg_level_transfer(glbx_mg* h, int level) { return (h && level >= 0 && level < (int)h->trs.size()) ? h->trs[level] : 0; }
void* glbx_mg_null_vector(glbx_mg* h, int level, int v) {
  if (!h || level < 0 || level >= (int)h->null_dev.size() || v < 0 || v >= (int)h->null_dev[level].size()) return 0;
  return h->null_dev[level][v];
}
void glbx_mg_setup_seconds(glbx_mg* h, double out[4]) {
  for (int i = 0; i < 4; i++) out[i] = h->setup_seconds[i];
}

void glbx_mg_set(glbx_mg* h, int in_smooth_type, int n_pre, int n_post, int in_solve_type, int n_max, int n_restart,
                 double rel_res, int mlevel_type, int quiet) {
  h->pc.in_smooth_type = (minv_inverter)in_smooth_type;
  h->pc.in_solve_type = (inner_solver)in_solve_type;
  h->pc.n_max = n_max;
  h->pc.n_restart = n_restart;
  h->pc.mlevel_type = (mg_multilevel_type)mlevel_type;
  h->pc.quiet = (quiet != 0);
  for (int i = 0; i < h->mg.n_refine; i++) {
    h->n_pre[i] = n_pre;
    h->n_post[i] = n_post;
    h->rel_res[i] = rel_res;
  }
}

// one cycle on the top level: d_lhs = M^-1 d_rhs
int glbx_mg_vcycle(glbx_mg* h, void* d_lhs, void* d_rhs) {
  h->mg.curr_level = 0;
  const int size = (int)glb_op_local_size(h->ops[0]);
  if (glb_vec_zero(glb_op_context(h->ops[0]), GLB_COMPLEX, size, d_lhs) != GLB_OK) return GLB_ERR_CUDA;
  mg_preconditioner_dev((zc*)d_lhs, (zc*)d_rhs, size, (void*)&h->pc, 0);
  return GLB_OK;
}

// minv_vector_gcr_var_precond(_restart) on the level-0 operator with mg_preconditioner_dev
int glbx_mg_vpgcr(glbx_mg* h, void* d_phi, void* d_phi0, int max_iter, double res, int restart_freq, int verbosity,
                  glbx_result* out) {
  h->mg.curr_level = 0;
  inversion_verbose_struct v;
  make_verb(verbosity, &v);
  const int size = (int)glb_op_local_size(h->ops[0]);
  void (*cb)(zc*, zc*, void*) = &glb200_apply_dev;
  void (*pcb)(zc*, zc*, int, void*, inversion_verbose_struct*) = &mg_preconditioner_dev;
  inversion_info inf;
  if (restart_freq > 0)
    inf = minv_vector_gcr_var_precond_restart_dev((zc*)d_phi, (zc*)d_phi0, size, max_iter, res, restart_freq, cb,
                                                  (void*)h->ops[0], pcb, (void*)&h->pc, &v);
  else
    inf = minv_vector_gcr_var_precond_dev((zc*)d_phi, (zc*)d_phi0, size, max_iter, res, cb, (void*)h->ops[0], pcb,
                                          (void*)&h->pc, &v);
  flatten(inf, out);
  return GLB_OK;
}

// dslash counters (mg_complex.h:104-136): out[5*(n_refine+1)] = krylov, presmooth, postsmooth, residual, nullvectors
void glbx_mg_counts(glbx_mg* h, int* out) {
  const int n = h->mg.n_refine + 1;
  const dslash_tracker* d = h->mg.dslash_count;
  for (int i = 0; i < n; i++) {
    out[i] = d->krylov[i];
    out[n + i] = d->presmooth[i];
    out[2 * n + i] = d->postsmooth[i];
    out[3 * n + i] = d->residual[i];
    out[4 * n + i] = d->nullvectors[i];
  }
}

}  // extern "C"
