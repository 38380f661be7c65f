// generic_inverters.h -- public solver API, drop-in for the reference's generic_inverters.h and
// the per-solver headers it pulls in (generic_cg.h:18-23, generic_cr.h, generic_gcr.h,
// generic_bicgstab.h:17-21, generic_bicgstab_l.h:17-21, generic_gmres.h:17-25, generic_cg_m.h:16-17).
//
// Same names, argument order, defaults and return type as the reference.  Every solver exists in
// two forms selected by where the vectors live:
//   * HOST vectors + the reference's operator callbacks (operators.h, coarse_stencil.h): the
//     call is the reference's call, unchanged.  The shell recognises the callback, builds the
//     matching device operator, uploads phi/phi0, solves on the GPU and downloads phi.
//   * DEVICE vectors + the device variant of the callback contract (glb200_device.h):
//     `void (*)(T* d_lhs, T* d_rhs, void* extra_info)` with the same semantics on device
//     pointers.  These carry the suffix _dev.
// There is no CPU solver in this library: an unknown host callback is an error unless the
// explicit parity shim is enabled (glb200_device.h: glb200_allow_host_callback_shim).
#ifndef GLB200_GENERIC_INVERTERS_H
#define GLB200_GENERIC_INVERTERS_H

#include <complex>
#include <string>
using std::complex;

#include "generic_traits.h"
#include "inverter_struct.h"
#include "verbosity.h"

// Both scalar types share every prototype.
#define GLB200_FOR_BOTH(M) M(double) M(complex<double>)

#define GLB200_DECL_BASIC(NAME)                                                                                   \
  inversion_info NAME(double* phi, double* phi0, int size, int max_iter, double res,                              \
                      void (*matrix_vector)(double*, double*, void*), void* extra_info,                           \
                      inversion_verbose_struct* verbosity = 0);                                                   \
  inversion_info NAME(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double res,            \
                      void (*matrix_vector)(complex<double>*, complex<double>*, void*), void* extra_info,         \
                      inversion_verbose_struct* verbosity = 0);

#define GLB200_DECL_RESTART(NAME)                                                                                 \
  inversion_info NAME(double* phi, double* phi0, int size, int max_iter, double res, int restart_freq,            \
                      void (*matrix_vector)(double*, double*, void*), void* extra_info,                           \
                      inversion_verbose_struct* verbosity = 0);                                                   \
  inversion_info NAME(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double res,            \
                      int restart_freq, void (*matrix_vector)(complex<double>*, complex<double>*, void*),         \
                      void* extra_info, inversion_verbose_struct* verbosity = 0);

// Conjugate gradient (Hermitian positive definite A)          generic_cg.h
GLB200_DECL_BASIC(minv_vector_cg)
GLB200_DECL_RESTART(minv_vector_cg_restart)
// Conjugate residual (Hermitian A)                            generic_cr.h
GLB200_DECL_BASIC(minv_vector_cr)
GLB200_DECL_RESTART(minv_vector_cr_restart)
// Generalised conjugate residual (any A)                      generic_gcr.h
GLB200_DECL_BASIC(minv_vector_gcr)
GLB200_DECL_RESTART(minv_vector_gcr_restart)
// BiCGStab (any A)                                            generic_bicgstab.h
GLB200_DECL_BASIC(minv_vector_bicgstab)
GLB200_DECL_RESTART(minv_vector_bicgstab_restart)
// GMRES (any A)                                               generic_gmres.h
GLB200_DECL_BASIC(minv_vector_gmres)
GLB200_DECL_RESTART(minv_vector_gmres_restart)

// BiCGStab-l (Sleijpen-Fokkema)                               generic_bicgstab_l.h
inversion_info minv_vector_bicgstab_l(double* phi, double* phi0, int size, int max_iter, double res, int l,
                                      void (*matrix_vector)(double*, double*, void*), void* extra_info,
                                      inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_bicgstab_l(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double res,
                                      int l, void (*matrix_vector)(complex<double>*, complex<double>*, void*),
                                      void* extra_info, inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_bicgstab_l_restart(double* phi, double* phi0, int size, int max_iter, double res,
                                              int restart_freq, int l, void (*matrix_vector)(double*, double*, void*),
                                              void* extra_info, inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_bicgstab_l_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter,
                                              double res, int restart_freq, int l,
                                              void (*matrix_vector)(complex<double>*, complex<double>*, void*),
                                              void* extra_info, inversion_verbose_struct* verbosity = 0);

// Multishift CG: solves (A + shifts[n]) phi[n] = phi0 for all n with one operator apply per iteration.
//                                                              generic_cg_m.h
inversion_info minv_vector_cg_m(double** phi, double* phi0, int n_shift, int size, int resid_freq_check, int max_iter,
                                double eps, double* shifts, void (*matrix_vector)(double*, double*, void*),
                                void* extra_info, bool worst_first = false, inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_cg_m(complex<double>** phi, complex<double>* phi0, int n_shift, int size,
                                int resid_freq_check, int max_iter, double eps, double* shifts,
                                void (*matrix_vector)(complex<double>*, complex<double>*, void*), void* extra_info,
                                bool worst_first = false, inversion_verbose_struct* verbosity = 0);

// Successive over-relaxation x <- x + omega (b - A x)          generic_sor.h:14-18
inversion_info minv_vector_sor(double* phi, double* phi0, int size, int max_iter, double eps, double omega,
                               void (*matrix_vector)(double*, double*, void*), void* extra_info,
                               inversion_verbose_struct* verb = 0);
inversion_info minv_vector_sor(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                               double omega, void (*matrix_vector)(complex<double>*, complex<double>*, void*),
                               void* extra_info, inversion_verbose_struct* verb = 0);
// Minimum residual (Saad 5.3.2), with and without relaxation    generic_minres.h:16-23
inversion_info minv_vector_minres(double* phi, double* phi0, int size, int max_iter, double eps,
                                  void (*matrix_vector)(double*, double*, void*), void* extra_info,
                                  inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_minres(double* phi, double* phi0, int size, int max_iter, double eps, double omega,
                                  void (*matrix_vector)(double*, double*, void*), void* extra_info,
                                  inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_minres(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                  void (*matrix_vector)(complex<double>*, complex<double>*, void*), void* extra_info,
                                  inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_minres(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                  double omega, void (*matrix_vector)(complex<double>*, complex<double>*, void*),
                                  void* extra_info, inversion_verbose_struct* verbosity = 0);

// Gauss-Jordan elimination used by GMRES; stays on the host (generic_gelim.h)
int gaussian_elimination(double* x, double* b, double** matrix, int size);
int gaussian_elimination(complex<double>* x, complex<double>* b, complex<double>** matrix, int size);
// several right-hand sides (x[k], b[k] of length size) and the matrix inverse (minv may alias matrix); generic_gelim.h:19-24
int gaussian_elimination_multi_rhs(double** x, double** b, double** matrix, int n_rhs, int size);
int gaussian_elimination_multi_rhs(complex<double>** x, complex<double>** b, complex<double>** matrix, int n_rhs, int size);
int gaussian_elimination_matrix_inverse(double** minv, double** matrix, int size);
int gaussian_elimination_matrix_inverse(complex<double>** minv, complex<double>** matrix, int size);

// Solver selection by enum (generic_inverters.h:70-111)
enum minv_inverter {
  MINV_CG = 0,
  MINV_CR = 1,
  MINV_GCR = 2,
  MINV_BICGSTAB = 3,
  MINV_BICGSTAB_L = 4,
  MINV_GMRES = 5,
  MINV_SOR = 6,
  MINV_MINRES = 7,
  MINV_INVALID = -1,
};

struct minv_inverter_params {
  double tol;
  int max_iters;
  bool restart;
  int restart_freq;
  double sor_omega;
  double minres_omega;
  int bicgstabl_l;
};

inversion_info minv_unpreconditioned(double* lhs, double* rhs, int size, minv_inverter type,
                                     minv_inverter_params& params, void (*matrix_vector)(double*, double*, void*),
                                     void* extra_info, inversion_verbose_struct* verbosity = 0);
inversion_info minv_unpreconditioned(complex<double>* lhs, complex<double>* rhs, int size, minv_inverter type,
                                     minv_inverter_params& params,
                                     void (*matrix_vector)(complex<double>*, complex<double>*, void*),
                                     void* extra_info, inversion_verbose_struct* verbosity = 0);

#endif
