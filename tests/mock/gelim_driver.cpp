// TEST INFRASTRUCTURE: prints the results of the dense elimination routines of generic_gelim.h on fixed random inputs
// (17 digits, on stderr); tests/test_reference_programs_cpu.py builds it against the reference and against host/ and compares.
#include <complex>
#include <cstdio>
#include <random>
#include <vector>
#include <iostream>
using namespace std;
int gaussian_elimination_multi_rhs(complex<double>** x, complex<double>** b, complex<double>** matrix, int n_rhs, int size);
int gaussian_elimination_matrix_inverse(double** minv, double** matrix, int size);
int main() {
  const int n = 7, nr = 3;
  mt19937 g(5); normal_distribution<> d;
  vector<vector<complex<double> > > M(n, vector<complex<double> >(n)), B(nr, vector<complex<double> >(n)), X(nr, vector<complex<double> >(n));
  vector<complex<double>*> Mp(n), Bp(nr), Xp(nr);
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) M[i][j] = complex<double>(d(g), d(g)); Mp[i] = M[i].data(); }
  for (int k = 0; k < nr; k++) { for (int i = 0; i < n; i++) B[k][i] = complex<double>(d(g), d(g)); Bp[k] = B[k].data(); Xp[k] = X[k].data(); }
  int rc = gaussian_elimination_multi_rhs(Xp.data(), Bp.data(), Mp.data(), nr, n);
  vector<vector<double> > R(n, vector<double>(n)), Ri(n, vector<double>(n)); vector<double*> Rp(n), Rip(n);
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) R[i][j] = d(g); Rp[i] = R[i].data(); Rip[i] = Ri[i].data(); }
  int rc2 = gaussian_elimination_matrix_inverse(Rip.data(), Rp.data(), n);
  fprintf(stderr, "%d %d\n", rc, rc2);
  for (int k = 0; k < nr; k++) for (int i = 0; i < n; i++) fprintf(stderr, "%.17g %.17g\n", X[k][i].real(), X[k][i].imag());
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) fprintf(stderr, "%.17g\n", Ri[i][j]);
  return 0;
}
