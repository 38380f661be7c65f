// generic_inverters_precond.h -- the preconditioned solver family, drop-in for the reference's
// generic_inverters_precond.h:28-58 and the headers it pulls in (generic_cg_precond.h:16-18,
// generic_cg_flex_precond.h, generic_gcr_var_precond.h:16-23, generic_bicgstab_precond.h, generic_precond.h),
// plus the multishift generic_cr_m.h:27-28 / generic_bicgstab_m.h:16-17.
//
// Same names, argument order and defaults as the reference.  HOST-vector forms accept the reference's
// operator callbacks (operators.h, coarse_stencil.h) and its stock preconditioners of generic_precond.h
// (identity_preconditioner, gcr_preconditioner); the DEVICE-vector forms (suffix _dev, glb200_device.h) accept
// any device preconditioner callback, e.g. mg_preconditioner_dev.
#ifndef GLB200_GENERIC_INVERTERS_PRECOND_H
#define GLB200_GENERIC_INVERTERS_PRECOND_H

#include "generic_inverters.h"

#define GLB200_PRECOND_ARGS(T)                                                                                      \
  void (*matrix_vector)(T*, T*, void*), void* extra_info,                                                           \
      void (*precond_matrix_vector)(T*, T*, int, void*, inversion_verbose_struct*), void* precond_info,             \
      inversion_verbose_struct* verbosity = 0

#define GLB200_DECL_PRECOND(T)                                                                                      \
  inversion_info minv_vector_cg_precond(T* phi, T* phi0, int size, int max_iter, double eps, GLB200_PRECOND_ARGS(T)); \
  inversion_info minv_vector_cg_flex_precond(T* phi, T* phi0, int size, int max_iter, double eps,                   \
                                             GLB200_PRECOND_ARGS(T));                                               \
  inversion_info minv_vector_cg_flex_precond_restart(T* phi, T* phi0, int size, int max_iter, double res,           \
                                                     int restart_freq, GLB200_PRECOND_ARGS(T));                     \
  inversion_info minv_vector_gcr_var_precond(T* phi, T* phi0, int size, int max_iter, double eps,                   \
                                             GLB200_PRECOND_ARGS(T));                                               \
  inversion_info minv_vector_gcr_var_precond_restart(T* phi, T* phi0, int size, int max_iter, double res,           \
                                                     int restart_freq, GLB200_PRECOND_ARGS(T));                     \
  inversion_info minv_vector_bicgstab_precond(T* phi, T* phi0, int size, int max_iter, double eps,                  \
                                              GLB200_PRECOND_ARGS(T));                                              \
  inversion_info minv_vector_bicgstab_precond_restart(T* phi, T* phi0, int size, int max_iter, double res,          \
                                                      int restart_freq, GLB200_PRECOND_ARGS(T));                    \
  inversion_info minv_vector_cr_m(T** phi, T* phi0, int n_shift, int size, int resid_freq_check, int max_iter,      \
                                  double eps, double* shifts, void (*matrix_vector)(T*, T*, void*), void* extra_info, \
                                  bool worst_first = false, inversion_verbose_struct* verbosity = 0);               \
  inversion_info minv_vector_bicgstab_m(T** phi, T* phi0, int n_shift, int size, int resid_freq_check, int max_iter, \
                                        double eps, double* shifts, void (*matrix_vector)(T*, T*, void*),           \
                                        void* extra_info, bool worst_first = false,                                 \
                                        inversion_verbose_struct* verbosity = 0);                                   \
  void identity_preconditioner(T* lhs, T* rhs, int size, void* extra_data, inversion_verbose_struct* verb = 0);     \
  void gcr_preconditioner(T* lhs, T* rhs, int size, void* extra_data, inversion_verbose_struct* verb = 0);     \
  void minres_preconditioner(T* lhs, T* rhs, int size, void* extra_data, inversion_verbose_struct* verb = 0);
GLB200_DECL_PRECOND(double)
GLB200_DECL_PRECOND(complex<double>)

// generic_precond.h:54-71 : `extra_data` of gcr_preconditioner (n_step GCR iterations on matrix_vector)
struct gcr_precond_struct_real {
  int n_step;
  double rel_res;
  void (*matrix_vector)(double*, double*, void*);
  void* matrix_extra_data;
};
struct gcr_precond_struct_complex {
  int n_step;
  double rel_res;
  void (*matrix_vector)(complex<double>*, complex<double>*, void*);
  void* matrix_extra_data;
};

// generic_precond.h:27-43 : `extra_data` of minres_preconditioner (n_step MinRes iterations on matrix_vector)
struct minres_precond_struct_real {
  int n_step;
  double rel_res;
  void (*matrix_vector)(double*, double*, void*);
  void* matrix_extra_data;
};
struct minres_precond_struct_complex {
  int n_step;
  double rel_res;
  void (*matrix_vector)(complex<double>*, complex<double>*, void*);
  void* matrix_extra_data;
};

// generic_inverters_precond.h:30-58
enum minv_inverter_precond {
  MINV_PRE_CG = 0,
  MINV_PRE_FPCG = 1,
  MINV_PRE_VPGCR = 2,
  MINV_PRE_BICGSTAB = 3,
  MINV_PRE_INVALID = -1,
};
struct minv_inverter_precond_params {
  double tol;
  int max_iters;
  bool restart;
  int restart_freq;
};
inversion_info minv_preconditioned(double* lhs, double* rhs, int size, minv_inverter_precond type,
                                   minv_inverter_precond_params& params, GLB200_PRECOND_ARGS(double));
inversion_info minv_preconditioned(complex<double>* lhs, complex<double>* rhs, int size, minv_inverter_precond type,
                                   minv_inverter_precond_params& params, GLB200_PRECOND_ARGS(complex<double>));

#endif
