// glb200_device.h -- the device side of the drop-in: device-resident vectors and the device
// variant of the reference's operator-callback contract.
//
// Reference contract (e.g. generic_cg.h:18-19):   void (*matrix_vector)(T* lhs, T* rhs, void* extra_info)
//   computes lhs = A rhs, lhs fully overwritten, lhs must not alias rhs, no return value.
// Device variant (same signature, same semantics): lhs and rhs are DEVICE pointers and the
//   callback must enqueue its work on glb_stream(glb200_default_context()) without synchronising.
// The standard device callback is glb200_apply_dev, whose extra_info is a glb_operator* from
// include/glb200.h; solvers that see it use the fused kernels (apply+dot, device-resident CG).
#ifndef GLB200_DEVICE_H
#define GLB200_DEVICE_H

#include <complex>

#include "generic_inverters.h"
#include "generic_inverters_precond.h"
#include "glb200.h"

// Process-wide context used by the host-pointer entry points and by direct operator calls.
// Device = $GLB200_DEVICE, else $LOCAL_RANK, else 0.  Throws std::runtime_error without a GPU.
glb_context* glb200_default_context();
void glb200_set_default_context(glb_context* ctx);  // e.g. one created with a slab communicator

// The standard device callback: extra_info is a glb_operator*.
void glb200_apply_dev(double* d_lhs, double* d_rhs, void* glb_operator_handle);
void glb200_apply_dev(std::complex<double>* d_lhs, std::complex<double>* d_rhs, void* glb_operator_handle);

// Parity-test shim: allow solvers called with HOST vectors and an UNKNOWN host callback to run by
// download -> callback -> upload around every apply.  Off by default (no CPU path in the product).
void glb200_allow_host_callback_shim(bool allow);
// Force the host-scalar CG shell even when the device-resident CG applies (cross-checks).
void glb200_force_host_scalars(bool force);

// Map a host callback + its extra_info to a device operator (what the host-pointer solvers do
// internally).  Returns 0 if the callback is not one of operators.h / coarse_stencil.h.
glb_operator* glb200_operator_from_callback(void (*matrix_vector)(double*, double*, void*), void* extra_info);
glb_operator* glb200_operator_from_callback(void (*matrix_vector)(std::complex<double>*, std::complex<double>*, void*),
                                            void* extra_info);

#define GLB200_DECL_DEV(NAME)                                                                                        \
  inversion_info NAME(double* d_phi, double* d_phi0, int size, int max_iter, double res,                             \
                      void (*matrix_vector_dev)(double*, double*, void*), void* extra_info,                          \
                      inversion_verbose_struct* verbosity = 0);                                                      \
  inversion_info NAME(std::complex<double>* d_phi, std::complex<double>* d_phi0, int size, int max_iter, double res, \
                      void (*matrix_vector_dev)(std::complex<double>*, std::complex<double>*, void*),               \
                      void* extra_info, inversion_verbose_struct* verbosity = 0);
#define GLB200_DECL_DEV_RESTART(NAME)                                                                                \
  inversion_info NAME(double* d_phi, double* d_phi0, int size, int max_iter, double res, int restart_freq,           \
                      void (*matrix_vector_dev)(double*, double*, void*), void* extra_info,                          \
                      inversion_verbose_struct* verbosity = 0);                                                      \
  inversion_info NAME(std::complex<double>* d_phi, std::complex<double>* d_phi0, int size, int max_iter, double res, \
                      int restart_freq,                                                                              \
                      void (*matrix_vector_dev)(std::complex<double>*, std::complex<double>*, void*),               \
                      void* extra_info, inversion_verbose_struct* verbosity = 0);

GLB200_DECL_DEV(minv_vector_cg_dev)
GLB200_DECL_DEV_RESTART(minv_vector_cg_restart_dev)
GLB200_DECL_DEV(minv_vector_cr_dev)
GLB200_DECL_DEV_RESTART(minv_vector_cr_restart_dev)
GLB200_DECL_DEV(minv_vector_gcr_dev)
GLB200_DECL_DEV_RESTART(minv_vector_gcr_restart_dev)
GLB200_DECL_DEV(minv_vector_bicgstab_dev)
GLB200_DECL_DEV_RESTART(minv_vector_bicgstab_restart_dev)
GLB200_DECL_DEV(minv_vector_gmres_dev)
GLB200_DECL_DEV_RESTART(minv_vector_gmres_restart_dev)

// generic_sor.h:14-18, generic_minres.h:16-23 (relaxation parameter omega; MinRes also without it)
#define GLB200_DECL_DEV_RELAX(T)                                                                                   \
  inversion_info minv_vector_sor_dev(T* d_phi, T* d_phi0, int size, int max_iter, double eps, double omega,        \
                                     void (*matrix_vector_dev)(T*, T*, void*), void* extra_info,                   \
                                     inversion_verbose_struct* verbosity = 0);                                     \
  inversion_info minv_vector_minres_dev(T* d_phi, T* d_phi0, int size, int max_iter, double eps, double omega,     \
                                        void (*matrix_vector_dev)(T*, T*, void*), void* extra_info,                \
                                        inversion_verbose_struct* verbosity = 0);                                  \
  inversion_info minv_vector_minres_dev(T* d_phi, T* d_phi0, int size, int max_iter, double eps,                   \
                                        void (*matrix_vector_dev)(T*, T*, void*), void* extra_info,                \
                                        inversion_verbose_struct* verbosity = 0);
GLB200_DECL_DEV_RELAX(double)
GLB200_DECL_DEV_RELAX(std::complex<double>)

inversion_info minv_vector_bicgstab_l_dev(double* d_phi, double* d_phi0, int size, int max_iter, double res, int l,
                                          void (*matrix_vector_dev)(double*, double*, void*), void* extra_info,
                                          inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_bicgstab_l_dev(std::complex<double>* d_phi, std::complex<double>* d_phi0, int size,
                                          int max_iter, double res, int l,
                                          void (*matrix_vector_dev)(std::complex<double>*, std::complex<double>*, void*),
                                          void* extra_info, inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_bicgstab_l_restart_dev(double* d_phi, double* d_phi0, int size, int max_iter, double res,
                                                  int restart_freq, int l,
                                                  void (*matrix_vector_dev)(double*, double*, void*), void* extra_info,
                                                  inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_bicgstab_l_restart_dev(
    std::complex<double>* d_phi, std::complex<double>* d_phi0, int size, int max_iter, double res, int restart_freq,
    int l, void (*matrix_vector_dev)(std::complex<double>*, std::complex<double>*, void*), void* extra_info,
    inversion_verbose_struct* verbosity = 0);

// phi: HOST array of n_shift DEVICE pointers (permuted and restored like the reference does)
inversion_info minv_vector_cg_m_dev(double** d_phi, double* d_phi0, int n_shift, int size, int resid_freq_check,
                                    int max_iter, double eps, double* shifts,
                                    void (*matrix_vector_dev)(double*, double*, void*), void* extra_info,
                                    bool worst_first = false, inversion_verbose_struct* verbosity = 0);
inversion_info minv_vector_cg_m_dev(std::complex<double>** d_phi, std::complex<double>* d_phi0, int n_shift, int size,
                                    int resid_freq_check, int max_iter, double eps, double* shifts,
                                    void (*matrix_vector_dev)(std::complex<double>*, std::complex<double>*, void*),
                                    void* extra_info, bool worst_first = false,
                                    inversion_verbose_struct* verbosity = 0);

// Variably preconditioned GCR (generic_gcr_var_precond.h:16-23) on device vectors.  The preconditioner is
// the device variant of the reference's precond contract: d_lhs = M^-1 d_rhs on the context's stream.
#define GLB200_DECL_VPGCR(T)                                                                                       \
  inversion_info minv_vector_gcr_var_precond_dev(                                                                  \
      T* d_phi, T* d_phi0, int size, int max_iter, double res, void (*matrix_vector_dev)(T*, T*, void*),           \
      void* extra_info, void (*precond_matrix_vector_dev)(T*, T*, int, void*, inversion_verbose_struct*),          \
      void* precond_info, inversion_verbose_struct* verbosity = 0);                                                \
  inversion_info minv_vector_gcr_var_precond_restart_dev(                                                          \
      T* d_phi, T* d_phi0, int size, int max_iter, double res, int restart_freq,                                   \
      void (*matrix_vector_dev)(T*, T*, void*), void* extra_info,                                                  \
      void (*precond_matrix_vector_dev)(T*, T*, int, void*, inversion_verbose_struct*), void* precond_info,        \
      inversion_verbose_struct* verbosity = 0);
GLB200_DECL_VPGCR(double)
GLB200_DECL_VPGCR(std::complex<double>)

// The rest of the preconditioned family and the multishift CR / BiCGStab on device vectors
// (generic_cg_precond.h, generic_cg_flex_precond.h, generic_bicgstab_precond.h, generic_cr_m.h, generic_bicgstab_m.h)
#define GLB200_DEV_PRECOND_ARGS(T)                                                                                  \
  void (*matrix_vector_dev)(T*, T*, void*), void* extra_info,                                                       \
      void (*precond_matrix_vector_dev)(T*, T*, int, void*, inversion_verbose_struct*), void* precond_info,         \
      inversion_verbose_struct* verbosity = 0
#define GLB200_DECL_DEV_PRECOND(T)                                                                                  \
  inversion_info minv_vector_cg_precond_dev(T* d_phi, T* d_phi0, int size, int max_iter, double eps,                \
                                            GLB200_DEV_PRECOND_ARGS(T));                                            \
  inversion_info minv_vector_cg_flex_precond_dev(T* d_phi, T* d_phi0, int size, int max_iter, double eps,           \
                                                 GLB200_DEV_PRECOND_ARGS(T));                                       \
  inversion_info minv_vector_cg_flex_precond_restart_dev(T* d_phi, T* d_phi0, int size, int max_iter, double res,   \
                                                         int restart_freq, GLB200_DEV_PRECOND_ARGS(T));             \
  inversion_info minv_vector_bicgstab_precond_dev(T* d_phi, T* d_phi0, int size, int max_iter, double eps,          \
                                                  GLB200_DEV_PRECOND_ARGS(T));                                      \
  inversion_info minv_vector_bicgstab_precond_restart_dev(T* d_phi, T* d_phi0, int size, int max_iter, double res,  \
                                                          int restart_freq, GLB200_DEV_PRECOND_ARGS(T));            \
  inversion_info minv_vector_cr_m_dev(T** d_phi, T* d_phi0, int n_shift, int size, int resid_freq_check, int max_iter, \
                                      double eps, double* shifts, void (*matrix_vector_dev)(T*, T*, void*),         \
                                      void* extra_info, bool worst_first = false,                                   \
                                      inversion_verbose_struct* verbosity = 0);                                     \
  inversion_info minv_vector_bicgstab_m_dev(T** d_phi, T* d_phi0, int n_shift, int size, int resid_freq_check,      \
                                            int max_iter, double eps, double* shifts,                               \
                                            void (*matrix_vector_dev)(T*, T*, void*), void* extra_info,             \
                                            bool worst_first = false, inversion_verbose_struct* verbosity = 0);     \
  /* generic_precond.cpp:23-77 on device vectors: lhs = rhs; n_step GCR iterations on gps->matrix_vector (a device   \
     callback with its extra_info) from the lhs handed in */                                                        \
  void identity_preconditioner_dev(T* d_lhs, T* d_rhs, int size, void* extra_data, inversion_verbose_struct* verb = 0); \
  void gcr_preconditioner_dev(T* d_lhs, T* d_rhs, int size, void* extra_data, inversion_verbose_struct* verb = 0); \
  void minres_preconditioner_dev(T* d_lhs, T* d_rhs, int size, void* extra_data, inversion_verbose_struct* verb = 0);
GLB200_DECL_DEV_PRECOND(double)
GLB200_DECL_DEV_PRECOND(std::complex<double>)

inversion_info minv_preconditioned_dev(double* d_lhs, double* d_rhs, int size, minv_inverter_precond type,
                                       minv_inverter_precond_params& params, GLB200_DEV_PRECOND_ARGS(double));
inversion_info minv_preconditioned_dev(std::complex<double>* d_lhs, std::complex<double>* d_rhs, int size,
                                       minv_inverter_precond type, minv_inverter_precond_params& params,
                                       GLB200_DEV_PRECOND_ARGS(std::complex<double>));

// minv_unpreconditioned (generic_inverters.h:109-111) on device vectors
inversion_info minv_unpreconditioned_dev(double* d_lhs, double* d_rhs, int size, minv_inverter type,
                                         minv_inverter_params& params, void (*matrix_vector_dev)(double*, double*, void*),
                                         void* extra_info, inversion_verbose_struct* verbosity = 0);
inversion_info minv_unpreconditioned_dev(std::complex<double>* d_lhs, std::complex<double>* d_rhs, int size,
                                         minv_inverter type, minv_inverter_params& params,
                                         void (*matrix_vector_dev)(std::complex<double>*, std::complex<double>*, void*),
                                         void* extra_info, inversion_verbose_struct* verbosity = 0);

#endif
