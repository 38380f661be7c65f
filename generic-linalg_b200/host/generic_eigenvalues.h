// generic_eigenvalues.h -- largest-eigenvalue estimate by power iteration on the accelerated path: drop-in for the
// reference's generic_eigenvalues.h / generic_poweriter.cpp.  One operator apply, one norm and one scaling per step,
// vectors on the device; `eigenvalue_info` is returned by value as in the reference.
#ifndef GLB200_GENERIC_EIGENVALUES_H
#define GLB200_GENERIC_EIGENVALUES_H

#include <complex>
#include <string>
using std::complex;

// generic_eigenvalues.h:17-23
struct eigenvalue_info {
  double relative_diff;  // |beta_k - beta_(k-1)| of the last step (despite the name an absolute difference, as in the reference)
  int iter;              // steps taken
  bool success;          // the difference fell below relres before max_iter
  std::string name;      // "Power Iteration"
};

// generic_poweriter.cpp:23-88: q = phi0/|phi0|; repeat x = A q, beta = |x|, stop when |beta - beta_prev| < relres,
// q = x/beta.  *eig receives the last beta.  HOST vectors; the operator is one of operators.h (or any host callback when
// the shim is enabled, see glb200_device.h).
eigenvalue_info eig_vector_poweriter(double* eig, double* phi0, int size, int max_iter, double relres,
                                     void (*matrix_vector)(double*, double*, void*), void* extra_info);
// the same on DEVICE vectors with a device callback (glb200_apply_dev + operator handle)
eigenvalue_info eig_vector_poweriter_dev(double* eig, double* d_phi0, int size, int max_iter, double relres,
                                         void (*matrix_vector_dev)(double*, double*, void*), void* extra_info);

#endif
