// gelim.cpp -- dense Gauss-Jordan elimination with partial pivoting for the small (restart x restart)
// GMRES least-squares system.  Host code by design ("gelim stays on host"); follows
// generic_gelim.cpp:19-119 (double) and :122-225 (complex): pivot on the largest |entry| of the
// column at or below the diagonal, scale the pivot row, eliminate it from every other row.
// Returns 0 when a pivot column is entirely zero (GMRES treats that as "already converged").
#include <cmath>
#include <complex>
#include <vector>

#include "generic_inverters.h"

namespace {
template <typename T>
int gauss_jordan(T* x, T* b, T** matrix, int n) {
  std::vector<std::vector<T> > aug(n, std::vector<T>(n + 1));
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) aug[i][j] = matrix[i][j];
    aug[i][n] = b[i];
  }
  for (int col = 0; col < n; col++) {
    int pivot = -1;
    double biggest = 0.0;
    for (int row = col; row < n; row++) {
      const double mag = std::abs(aug[row][col]);
      if (mag > biggest) {
        biggest = mag;
        pivot = row;
      }
    }
    if (pivot < 0) return 0;
    if (pivot != col)
      for (int j = col; j <= n; j++) std::swap(aug[col][j], aug[pivot][j]);
    for (int j = col + 1; j <= n; j++) aug[col][j] = aug[col][j] / aug[col][col];
    aug[col][col] = 1.0;
    for (int row = 0; row < n; row++) {
      if (row == col) continue;
      for (int j = col + 1; j <= n; j++) aug[row][j] = aug[row][j] - aug[row][col] * aug[col][j];
      aug[row][col] = 0.0;
    }
  }
  for (int i = 0; i < n; i++) x[i] = aug[i][n];
  return 1;
}

// generic_gelim.cpp:228-388 (double), :390-550 (complex): several right-hand sides at once.  x[k], b[k] are the k-th
// solution / right-hand side (length n).  Same operation order as the reference -- forward elimination with partial
// pivoting (row_j += row_i * (-a_ji/a_ii), columns i..), then back substitution from the last row up with the pivot
// row normalised afterwards -- without its debugging output (the reference prints the augmented matrix after every step).
template <typename T>
int gauss_multi(T** x, T** b, T** matrix, int n_rhs, int n) {
  const int w = n + n_rhs;
  std::vector<std::vector<T> > a(n, std::vector<T>(w));
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) a[i][j] = matrix[i][j];
    for (int k = 0; k < n_rhs; k++) a[i][n + k] = b[k][i];
  }
  for (int i = 0; i < n; i++) {
    int pivot = -1;
    double biggest = 0.0;
    for (int j = i; j < n; j++) {
      if (std::abs(a[j][i]) > biggest) {
        pivot = j;
        biggest = std::abs(a[j][i]);
      }
    }
    if (pivot < 0) return 0;
    if (pivot != i) a[i].swap(a[pivot]);
    for (int j = i + 1; j < n; j++) {
      const T factor = -a[j][i] / a[i][i];
      for (int k = i; k < w; k++) a[j][k] = a[j][k] + a[i][k] * factor;
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    for (int j = i - 1; j >= 0; j--) {
      const T factor = -a[j][i] / a[i][i];
      for (int k = i; k < w; k++) a[j][k] = a[j][k] + a[i][k] * factor;
    }
    for (int k = i + 1; k < w; k++) a[i][k] /= a[i][i];
    a[i][i] = 1.0;
  }
  for (int k = 0; k < n_rhs; k++)
    for (int i = 0; i < n; i++) x[k][i] = a[i][n + k];
  return 1;
}

// generic_gelim.cpp:553-596 / :598-641: the inverse through n unit right-hand sides; minv may be `matrix` itself
template <typename T>
int gauss_inverse(T** minv, T** matrix, int n) {
  std::vector<std::vector<T> > unit(n, std::vector<T>(n, T(0.0)));
  std::vector<T*> rows(n);
  for (int i = 0; i < n; i++) {
    unit[i][i] = 1.0;
    rows[i] = unit[i].data();
  }
  const int rc = gauss_multi<T>(minv, rows.data(), matrix, n, n);
  for (int i = 0; i < n - 1; i++)  // solutions come out as rows: transpose
    for (int j = i + 1; j < n; j++) std::swap(minv[i][j], minv[j][i]);
  return rc;
}
}  // namespace

int gaussian_elimination_multi_rhs(double** x, double** b, double** matrix, int n_rhs, int size) {
  return gauss_multi<double>(x, b, matrix, n_rhs, size);
}
int gaussian_elimination_multi_rhs(complex<double>** x, complex<double>** b, complex<double>** matrix, int n_rhs, int size) {
  return gauss_multi<complex<double> >(x, b, matrix, n_rhs, size);
}
int gaussian_elimination_matrix_inverse(double** minv, double** matrix, int size) {
  return gauss_inverse<double>(minv, matrix, size);
}
int gaussian_elimination_matrix_inverse(complex<double>** minv, complex<double>** matrix, int size) {
  return gauss_inverse<complex<double> >(minv, matrix, size);
}

int gaussian_elimination(double* x, double* b, double** matrix, int size) {
  return gauss_jordan<double>(x, b, matrix, size);
}
int gaussian_elimination(complex<double>* x, complex<double>* b, complex<double>** matrix, int size) {
  return gauss_jordan<complex<double> >(x, b, matrix, size);
}
