// ref_shim.cpp -- C-ABI shim over the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).  This file contains no
// algorithm of its own: every numerical routine it calls is compiled from
// /root/reference by oracle/Makefile (outputs only in oracle/_ref/).  The three
// operators the reference only defines inside its example/test main files
// (square_laplace.cpp:182, imag_laplace.cpp:126, tests/multishift/multishift.cpp:634,677)
// are cut out of those files at build time by the Makefile into
// oracle/_ref/extracted_*.inc and compiled here verbatim; their compile-time
// `N` / `MASS` macros are mapped onto variables so the lattice size can vary.
#define ORC_PREFIX ref_
#include "oracle_api.h"

#include <complex>
#include <cstring>
#include <random>
#include <string>
#include <vector>

// Reference headers (include paths supplied by the Makefile).
#include "generic_inverters.h"
#include "generic_cg_m.h"
#include "generic_cr_m.h"
#include "generic_bicgstab_m.h"
#include "generic_inverters_precond.h"
#include "generic_vector.h"
#include "u1_utils.h"
#include "operators.h"
#include "lattice.h"
#include "coarse_stencil.h"
#include "operators_stencil.h"
#include "mg_complex.h"

using std::complex;
typedef complex<double> cplx;

// ---- operators that live in the reference's main files, compiled verbatim ----
namespace extracted_real_nc {  // tests/multishift/multishift.cpp:634-726
#include "extracted_multishift_real_ops.inc"
}
static int g_ref_N = 0;
static double g_ref_MASS = 0.0;
#define N g_ref_N
#define MASS g_ref_MASS
namespace extracted_square_laplace {  // square_laplace.cpp:182-222
#include "extracted_square_laplacian_real.inc"
}
namespace extracted_imag_laplace {  // imag_laplace.cpp:126-166
#include "extracted_square_laplacian_imag.inc"
}
#undef N
#undef MASS

namespace {

struct RefOp {
  orc_op_desc d;
  staggered_u1_op stagif;
  Lattice* lat = nullptr;
  stencil_2d* stenc = nullptr;
  bool is_complex = true;
  int size = 0;
  ~RefOp() {
    delete stenc;
    delete lat;
  }
};

void apply_c(RefOp* op, cplx* lhs, cplx* rhs) {
  void* e = (void*)&op->stagif;
  switch (op->d.kind) {
    case ORC_OP_LAPLACE_IMAG:
      g_ref_N = op->d.X;
      g_ref_MASS = op->d.mass;
      extracted_imag_laplace::square_laplacian(lhs, rhs, nullptr);
      break;
    case ORC_OP_LAPLACE_NC: square_laplace(lhs, rhs, e); break;
    case ORC_OP_LAPLACE_U1: square_laplace_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_FREE: square_staggered(lhs, rhs, e); break;
    case ORC_OP_STAG_U1: square_staggered_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_GAMMA5_U1: square_staggered_gamma5_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_GAMMA5_FREE: square_staggered_gamma5(lhs, rhs, e); break;
    case ORC_OP_STAG_DAGGER_U1: square_staggered_dagger_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_NORMAL_U1: square_staggered_normal_u1(lhs, rhs, e); break;
    case ORC_OP_GAMMA5: gamma_5(lhs, rhs, e); break;
    case ORC_OP_STAG_DEO_U1: square_staggered_deo_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_DOE_U1: square_staggered_doe_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_M2MDEODOE_U1: square_staggered_m2mdeodoe_u1(lhs, rhs, e); break;
    case ORC_OP_SYMMSHIFT_X: staggered_symmshift_x(lhs, rhs, e); break;
    case ORC_OP_SYMMSHIFT_Y: staggered_symmshift_y(lhs, rhs, e); break;
    case ORC_OP_STAG_2LINK_U1: square_staggered_2linklaplace_u1(lhs, rhs, e); break;
    case ORC_OP_STAG_INDEX: staggered_index_operator(lhs, rhs, e); break;
    case ORC_OP_STENCIL:
    case ORC_OP_STENCIL_FROM_STAG:
      switch (op->d.view) {
        case 1: apply_square_staggered_m2mdeodoe_stencil(lhs, rhs, (void*)op->stenc); break;
        case 2: apply_square_staggered_m2mdtbdbt_stencil(lhs, rhs, (void*)op->stenc); break;
        case 3: apply_square_staggered_normal_eo_stencil(lhs, rhs, (void*)op->stenc); break;
        case 4: apply_square_staggered_normal_tb_stencil(lhs, rhs, (void*)op->stenc); break;
        case 5: apply_square_staggered_dagger_eo_stencil(lhs, rhs, (void*)op->stenc); break;
        case 6: apply_square_staggered_dagger_tb_stencil(lhs, rhs, (void*)op->stenc); break;
        default: apply_stencil_2d(lhs, rhs, (void*)op->stenc); break;
      }
      break;
    default: break;
  }
}

void apply_r(RefOp* op, double* lhs, double* rhs) {
  void* e = (void*)&op->stagif;
  switch (op->d.kind) {
    case ORC_OP_LAPLACE_REAL:
      g_ref_N = op->d.X;
      g_ref_MASS = op->d.mass;
      extracted_square_laplace::square_laplacian(lhs, rhs, nullptr);
      break;
    case ORC_OP_LAPLACE_REAL_NC: extracted_real_nc::square_laplace(lhs, rhs, e); break;
    case ORC_OP_STAG_FREE_REAL: extracted_real_nc::square_staggered(lhs, rhs, e); break;
    default: break;
  }
}

// The solvers take a plain function pointer + void*; route through these.
void cb_c(cplx* lhs, cplx* rhs, void* extra) { apply_c((RefOp*)extra, lhs, rhs); }
void cb_r(double* lhs, double* rhs, void* extra) { apply_r((RefOp*)extra, lhs, rhs); }

void fill_result(const inversion_info& info, orc_result* out) {
  std::memset(out, 0, sizeof(*out));
  out->resSq = info.resSq;
  out->iter = info.iter;
  out->success = info.success ? 1 : 0;
  out->ops_count = info.ops_count;
  out->n_rhs = info.n_rhs;
  if (info.resSqmrhs && info.n_rhs > 0)
    for (int i = 0; i < info.n_rhs && i < 32; i++) out->resSqmrhs[i] = info.resSqmrhs[i];
  std::strncpy(out->name, info.name.c_str(), sizeof(out->name) - 1);
}

void make_verb(int verbosity, inversion_verbose_struct* v) {
  v->verbosity = (inversion_verbose_level)verbosity;
  v->verb_prefix = "[ref] ";
  v->precond_verbosity = VERB_NONE;
  v->precond_verb_prefix = "";
}

template <typename T>
inversion_info run_solver(int solver, T* phi, T* phi0, int size, int max_iter, double eps, int rf, int l,
                          void (*cb)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {
  switch (solver) {
    case ORC_CG: return minv_vector_cg(phi, phi0, size, max_iter, eps, cb, extra, verb);
    case ORC_CG_RESTART: return minv_vector_cg_restart(phi, phi0, size, max_iter, eps, rf, cb, extra, verb);
    case ORC_CR: return minv_vector_cr(phi, phi0, size, max_iter, eps, cb, extra, verb);
    case ORC_CR_RESTART: return minv_vector_cr_restart(phi, phi0, size, max_iter, eps, rf, cb, extra, verb);
    case ORC_GCR: return minv_vector_gcr(phi, phi0, size, max_iter, eps, cb, extra, verb);
    case ORC_GCR_RESTART: return minv_vector_gcr_restart(phi, phi0, size, max_iter, eps, rf, cb, extra, verb);
    case ORC_BICGSTAB: return minv_vector_bicgstab(phi, phi0, size, max_iter, eps, cb, extra, verb);
    case ORC_BICGSTAB_RESTART:
      return minv_vector_bicgstab_restart(phi, phi0, size, max_iter, eps, rf, cb, extra, verb);
    case ORC_BICGSTAB_L: return minv_vector_bicgstab_l(phi, phi0, size, max_iter, eps, l, cb, extra, verb);
    case ORC_BICGSTAB_L_RESTART:
      return minv_vector_bicgstab_l_restart(phi, phi0, size, max_iter, eps, rf, l, cb, extra, verb);
    case ORC_GMRES: return minv_vector_gmres(phi, phi0, size, max_iter, eps, cb, extra, verb);
    case ORC_GMRES_RESTART: return minv_vector_gmres_restart(phi, phi0, size, max_iter, eps, rf, cb, extra, verb);
    default: return inversion_info();
  }
}

}  // namespace

extern "C" {

const char* ref_kind(void) { return "reference"; }

void* ref_rng_new(unsigned seed) { return new std::mt19937(seed); }
void ref_rng_free(void* rng) { delete (std::mt19937*)rng; }
void ref_gauss_gauge_u1(void* rng, double* links, int X, int Y, double beta) {
  gauss_gauge_u1((cplx*)links, X, Y, *(std::mt19937*)rng, beta);
}
void ref_unit_gauge_u1(double* links, int X, int Y) { unit_gauge_u1((cplx*)links, X, Y); }
void ref_gaussian_real(void* rng, double* v, int n) { gaussian<double>(v, n, *(std::mt19937*)rng); }
void ref_gaussian_complex(void* rng, double* v, int n) { gaussian<double>((cplx*)v, n, *(std::mt19937*)rng); }
int ref_read_gauge_u1(double* links, int X, int Y, const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) return 1;
  fclose(f);
  read_gauge_u1((cplx*)links, X, Y, std::string(path));
  return 0;
}
void ref_plaquette_u1(const double* links, int X, int Y, double out[2]) {
  cplx p = get_plaquette_u1((cplx*)links, X, Y);
  out[0] = p.real();
  out[1] = p.imag();
}

void ref_dot(int is_complex, const double* a, const double* b, int n, double out[2]) {
  if (is_complex) {
    cplx r = dot<double>((cplx*)a, (cplx*)b, n);
    out[0] = r.real();
    out[1] = r.imag();
  } else {
    out[0] = dot<double>((double*)a, (double*)b, n);
    out[1] = 0.0;
  }
}
double ref_norm2sq(int is_complex, const double* a, int n) {
  return is_complex ? norm2sq<double>((cplx*)a, n) : norm2sq<double>((double*)a, n);
}
double ref_diffnorm2sq(int is_complex, const double* a, const double* b, int n) {
  return is_complex ? diffnorm2sq<double>((cplx*)a, (cplx*)b, n) : diffnorm2sq<double>((double*)a, (double*)b, n);
}

void* ref_op_prepare(const orc_op_desc* d) {
  RefOp* op = new RefOp();
  op->d = *d;
  op->stagif.lattice = (cplx*)d->links;
  op->stagif.mass = d->mass;
  op->stagif.x_fine = d->X;
  op->stagif.y_fine = d->Y;
  op->stagif.Nc = d->Nc > 0 ? d->Nc : 1;
  op->stagif.wilson_coeff = d->wilson_coeff;
  const int V = d->X * d->Y;
  op->size = V;
  op->is_complex = !(d->kind == ORC_OP_LAPLACE_REAL || d->kind == ORC_OP_LAPLACE_REAL_NC ||
                     d->kind == ORC_OP_STAG_FREE_REAL);
  if (d->kind == ORC_OP_LAPLACE_NC || d->kind == ORC_OP_LAPLACE_REAL_NC) op->size = V * op->stagif.Nc;
  if (d->kind == ORC_OP_STENCIL || d->kind == ORC_OP_STENCIL_FROM_STAG) {
    int dims[2] = {d->X, d->Y};
    const int nc = (d->kind == ORC_OP_STENCIL_FROM_STAG) ? 1 : op->stagif.Nc;
    op->lat = new Lattice(2, dims, nc);
    op->size = V * nc;
    if (d->kind == ORC_OP_STENCIL_FROM_STAG) {
      op->stenc = new stencil_2d(op->lat, 1);
      get_square_staggered_u1_stencil(op->stenc, &op->stagif);
    } else {
      op->stenc = new stencil_2d(op->lat, d->has_two ? 2 : 1, cplx(d->shift[0], d->shift[1]),
                                 cplx(d->eo_shift[0], d->eo_shift[1]), cplx(d->dof_shift[0], d->dof_shift[1]));
      const size_t m = (size_t)V * nc * nc;
      std::memcpy((void*)op->stenc->clover, d->clover, m * sizeof(cplx));
      std::memcpy((void*)op->stenc->hopping, d->hopping, 4 * m * sizeof(cplx));
      if (d->has_two) std::memcpy((void*)op->stenc->two_link, d->two_link, 8 * m * sizeof(cplx));
      op->stenc->generated = true;
    }
  }
  return op;
}
void ref_op_free(void* op) { delete (RefOp*)op; }
int ref_op_is_complex(void* op) { return ((RefOp*)op)->is_complex ? 1 : 0; }
int ref_op_size(void* op) { return ((RefOp*)op)->size; }
void ref_op_apply(void* opv, double* lhs, const double* rhs) {
  RefOp* op = (RefOp*)opv;
  if (op->is_complex)
    apply_c(op, (cplx*)lhs, (cplx*)rhs);
  else
    apply_r(op, lhs, (double*)rhs);
}

// apply_square_staggered_{eo,tb}prec_{prepare,reconstruct}_stencil (operators_stencil.cpp:179,217; mg_complex.cpp:1211,1252)
// prepare: out = f(a); reconstruct: out = f(a = lhs_part, b = rhs_other)
void ref_stencil_prec(void* opv, int top_bottom, int reconstruct, double* out, const double* a, const double* b) {
  RefOp* op = (RefOp*)opv;
  if (!reconstruct) {
    if (top_bottom)
      apply_square_staggered_tbprec_prepare_stencil((cplx*)out, (cplx*)a, op->stenc);
    else
      apply_square_staggered_eoprec_prepare_stencil((cplx*)out, (cplx*)a, op->stenc);
  } else {
    if (top_bottom)
      apply_square_staggered_tbprec_reconstruct_stencil((cplx*)out, (cplx*)a, (cplx*)b, op->stenc);
    else
      apply_square_staggered_eoprec_reconstruct_stencil((cplx*)out, (cplx*)a, (cplx*)b, op->stenc);
  }
}

void ref_stencil_apply_part(void* opv, int part, double* lhs, const double* rhs) {
  RefOp* op = (RefOp*)opv;
  if (part == 1) apply_stencil_2d_eo((cplx*)lhs, (cplx*)rhs, (void*)op->stenc);
  if (part == 2) apply_stencil_2d_oe((cplx*)lhs, (cplx*)rhs, (void*)op->stenc);
  if (part == 3) apply_stencil_2d_tb((cplx*)lhs, (cplx*)rhs, (void*)op->stenc);
  if (part == 4) apply_stencil_2d_bt((cplx*)lhs, (cplx*)rhs, (void*)op->stenc);
}
void ref_eoprec_prepare(void* opv, double* rhs_e, const double* rhs_orig) {
  RefOp* op = (RefOp*)opv;
  square_staggered_eoprec_prepare((cplx*)rhs_e, (cplx*)rhs_orig, (void*)&op->stagif);
}
void ref_eoprec_reconstruct(void* opv, double* lhs_full, const double* lhs_e, const double* rhs_o) {
  RefOp* op = (RefOp*)opv;
  square_staggered_eoprec_reconstruct((cplx*)lhs_full, (cplx*)lhs_e, (cplx*)rhs_o, (void*)&op->stagif);
}

int ref_solve(int solver, void* opv, double* phi, const double* phi0, int max_iter, double eps, int restart_freq,
              int l, int verbosity, orc_result* out) {
  RefOp* op = (RefOp*)opv;
  inversion_verbose_struct verb;
  make_verb(verbosity, &verb);
  inversion_info info;
  if (op->is_complex)
    info = run_solver<cplx>(solver, (cplx*)phi, (cplx*)phi0, op->size, max_iter, eps, restart_freq, l, cb_c, opv,
                            &verb);
  else
    info = run_solver<double>(solver, phi, (double*)phi0, op->size, max_iter, eps, restart_freq, l, cb_r, opv,
                              &verb);
  fill_result(info, out);
  return 0;
}

// minv_vector_sor (which = 0, generic_sor.cpp:24,122) / minv_vector_minres (which = 1, generic_minres.cpp:22,128)
int ref_solve_relax(int which, void* opv, double* phi, const double* phi0, int max_iter, double eps, double omega,
                    int verbosity, orc_result* out) {
  RefOp* op = (RefOp*)opv;
  inversion_verbose_struct verb;
  make_verb(verbosity, &verb);
  inversion_info info;
  if (op->is_complex)
    info = which == 0 ? minv_vector_sor((cplx*)phi, (cplx*)phi0, op->size, max_iter, eps, omega, cb_c, opv, &verb)
                      : minv_vector_minres((cplx*)phi, (cplx*)phi0, op->size, max_iter, eps, omega, cb_c, opv, &verb);
  else
    info = which == 0 ? minv_vector_sor(phi, (double*)phi0, op->size, max_iter, eps, omega, cb_r, opv, &verb)
                      : minv_vector_minres(phi, (double*)phi0, op->size, max_iter, eps, omega, cb_r, opv, &verb);
  fill_result(info, out);
  return 0;
}

int ref_solve_cg_m(void* opv, double** phi, const double* phi0, int n_shift, int resid_freq_check, int max_iter,
                   double eps, double* shifts, int worst_first, int verbosity, orc_result* out) {
  RefOp* op = (RefOp*)opv;
  inversion_verbose_struct verb;
  make_verb(verbosity, &verb);
  if (op->is_complex) {
    inversion_info info = minv_vector_cg_m((cplx**)phi, (cplx*)phi0, n_shift, op->size, resid_freq_check, max_iter,
                                           eps, shifts, cb_c, opv, worst_first != 0, &verb);
    fill_result(info, out);
  } else {
    inversion_info info = minv_vector_cg_m(phi, (double*)phi0, n_shift, op->size, resid_freq_check, max_iter, eps,
                                           shifts, cb_r, opv, worst_first != 0, &verb);
    fill_result(info, out);
  }
  return 0;
}


// ---- SURVEY 8f-4 (reference library only, no port): multishift CR / BiCGStab and the preconditioned family
// which: 0 minv_vector_cg_m, 1 minv_vector_cr_m (generic_cr_m.cpp:24,323), 2 minv_vector_bicgstab_m (generic_bicgstab_m.cpp:26,402)
int ref_solve_multi(int which, void* opv, double** phi, const double* phi0, int n_shift, int resid_freq_check,
                    int max_iter, double eps, double* shifts, int worst_first, int verbosity, orc_result* out) {
  RefOp* op = (RefOp*)opv;
  inversion_verbose_struct verb;
  make_verb(verbosity, &verb);
  inversion_info info(n_shift);
  if (op->is_complex) {
    cplx** p = (cplx**)phi;
    cplx* b = (cplx*)phi0;
    if (which == 0) info = minv_vector_cg_m(p, b, n_shift, op->size, resid_freq_check, max_iter, eps, shifts, cb_c, opv, worst_first != 0, &verb);
    if (which == 1) info = minv_vector_cr_m(p, b, n_shift, op->size, resid_freq_check, max_iter, eps, shifts, cb_c, opv, worst_first != 0, &verb);
    if (which == 2) info = minv_vector_bicgstab_m(p, b, n_shift, op->size, resid_freq_check, max_iter, eps, shifts, cb_c, opv, worst_first != 0, &verb);
  } else {
    double* b = (double*)phi0;
    if (which == 0) info = minv_vector_cg_m(phi, b, n_shift, op->size, resid_freq_check, max_iter, eps, shifts, cb_r, opv, worst_first != 0, &verb);
    if (which == 1) info = minv_vector_cr_m(phi, b, n_shift, op->size, resid_freq_check, max_iter, eps, shifts, cb_r, opv, worst_first != 0, &verb);
    if (which == 2) info = minv_vector_bicgstab_m(phi, b, n_shift, op->size, resid_freq_check, max_iter, eps, shifts, cb_r, opv, worst_first != 0, &verb);
  }
  fill_result(info, out);
  return 0;
}

}  // extern "C"

// solver: 0 PCG, 1 FPCG, 2 FPCG restart, 3 VPGCR, 4 VPGCR restart, 5 PBiCGStab, 6 PBiCGStab restart
// precond: 0 identity_preconditioner, 1 gcr_preconditioner (n_step iterations to rel_res on the same operator)
template <typename T, typename G>
static inversion_info run_precond(int solver, T* phi, T* b, int size, int max_iter, double eps, int rf,
                                  void (*cb)(T*, T*, void*), void* opv, int precond, int n_step, double rel_res,
                                  inversion_verbose_struct* verb) {
  G g;
  g.n_step = n_step;
  g.rel_res = rel_res;
  g.matrix_vector = cb;
  g.matrix_extra_data = opv;
  typedef void (*pfn)(T*, T*, int, void*, inversion_verbose_struct*);
  // precond 2: minres_preconditioner; its struct has the layout of the GCR one (generic_precond.h:27-70)
  pfn pc_gcr = &gcr_preconditioner, pc_id = &identity_preconditioner, pc_mr = &minres_preconditioner;
  pfn pc = (precond == 1) ? pc_gcr : (precond == 2) ? pc_mr : pc_id;
  void* pci = (precond != 0) ? (void*)&g : 0;
  switch (solver) {
    case 0: return minv_vector_cg_precond(phi, b, size, max_iter, eps, cb, opv, pc, pci, verb);
    case 1: return minv_vector_cg_flex_precond(phi, b, size, max_iter, eps, cb, opv, pc, pci, verb);
    case 2: return minv_vector_cg_flex_precond_restart(phi, b, size, max_iter, eps, rf, cb, opv, pc, pci, verb);
    case 3: return minv_vector_gcr_var_precond(phi, b, size, max_iter, eps, cb, opv, pc, pci, verb);
    case 4: return minv_vector_gcr_var_precond_restart(phi, b, size, max_iter, eps, rf, cb, opv, pc, pci, verb);
    case 5: return minv_vector_bicgstab_precond(phi, b, size, max_iter, eps, cb, opv, pc, pci, verb);
    case 6: return minv_vector_bicgstab_precond_restart(phi, b, size, max_iter, eps, rf, cb, opv, pc, pci, verb);
  }
  return inversion_info();
}
extern "C" {
int ref_solve_precond(int solver, void* opv, double* phi, const double* phi0, int max_iter, double eps, int restart_freq,
                      int precond, int n_step, double rel_res, int verbosity, orc_result* out) {
  RefOp* op = (RefOp*)opv;
  inversion_verbose_struct verb;
  make_verb(verbosity, &verb);
  inversion_info info;
  if (op->is_complex)
    info = run_precond<cplx, gcr_precond_struct_complex>(solver, (cplx*)phi, (cplx*)phi0, op->size, max_iter, eps,
                                                        restart_freq, cb_c, opv, precond, n_step, rel_res, &verb);
  else
    info = run_precond<double, gcr_precond_struct_real>(solver, phi, (double*)phi0, op->size, max_iter, eps, restart_freq,
                                                       cb_r, opv, precond, n_step, rel_res, &verb);
  fill_result(info, out);
  return 0;
}

}  // extern "C"
