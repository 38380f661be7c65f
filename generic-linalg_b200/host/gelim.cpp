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
}  // namespace

int gaussian_elimination(double* x, double* b, double** matrix, int size) {
  return gauss_jordan<double>(x, b, matrix, size);
}
int gaussian_elimination(complex<double>* x, complex<double>* b, complex<double>** matrix, int size) {
  return gauss_jordan<complex<double> >(x, b, matrix, size);
}
