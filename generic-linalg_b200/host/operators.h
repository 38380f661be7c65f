// operators.h -- the reference's operator callbacks (operator_utils/operators.h:7-100 and the
// in-file Laplacians of its examples), same names and signatures, executed on the GPU.
//
// Passed to a solver of generic_inverters.h they select the matching device kernel; called
// directly with host pointers they upload rhs, apply on the device and download lhs.
#ifndef GLB200_OPERATORS_H
#define GLB200_OPERATORS_H

#include <complex>
using namespace std;

// operator_utils/operators.h:7-15
struct staggered_u1_op {
  complex<double>* lattice;  // U(1) links, lattice[y*x_fine*2 + x*2 + mu], mu = 0:x, 1:y
  double mass;
  int x_fine;
  int y_fine;
  int Nc;               // only relevant for square_laplace
  double wilson_coeff;  // two-link Laplace coefficient of square_staggered_2linklaplace_u1
};

// operator_utils/operators.h:17-25
enum op_type {
  STAGGERED = 0,
  LAPLACE = 1,
  LAPLACE_NC2 = 2,
  G5_STAGGERED = 3,
  STAGGERED_NORMAL = 4,
  STAGGERED_INDEX = 5,
};
int get_stencil_size(op_type opt);  // operators.cpp:9-25

// operators.cpp:28   free Laplacian with Nc colours, diag 4+mass          extra: staggered_u1_op*
void square_laplace(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// tests/multishift/multishift.cpp:634   real twin of the above            extra: staggered_u1_op*
void square_laplace(double* lhs, double* rhs, void* extra_data);
// operators.cpp:73   gauged Laplacian
void square_laplace_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:127 / :184   staggered D, free and gauged
void square_staggered(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// tests/multishift/multishift.cpp:677   real free staggered                  extra: staggered_u1_op*
void square_staggered(double* lhs, double* rhs, void* extra_data);
void square_staggered_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:242   gamma_5 = (-1)^(x+y)
void gamma_5(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:262 / :316   gamma_5 D
void square_staggered_gamma5(complex<double>* lhs, complex<double>* rhs, void* extra_data);
void square_staggered_gamma5_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:372   D^dagger
void square_staggered_dagger_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:444   D^dagger D
void square_staggered_normal_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);

// operators.cpp:456 / :494   even/odd pieces: the hopping term on even (odd) sites, the other parity zeroed
void square_staggered_deo_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
void square_staggered_doe_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:549   m^2 - D_eo D_oe on even sites (odd sites zeroed): the e/o preconditioned, Hermitian system
void square_staggered_m2mdeodoe_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:528 / :574   rhs_e = m rhs - D_eo rhs (even) ; lhs_full = lhs_e (even), (rhs_o - D_oe lhs_e)/m (odd)
void square_staggered_eoprec_prepare(complex<double>* rhs_e, complex<double>* rhs_orig, void* extra_data);
void square_staggered_eoprec_reconstruct(complex<double>* lhs_full, complex<double>* lhs_e, complex<double>* rhs_o,
                                         void* extra_data);

// The examples' in-file Laplacians (square_laplace.cpp:182, unit_test.cpp:573, imag_laplace.cpp:126)
// take their size from `#define N` / `#define MASS`; here the same numbers travel in extra_data.
struct laplace_op {
  int N;          // N x N lattice
  double mass_sq; // MASS of the examples (diag = 4 + MASS, or 4 + MASS + i for the complex one)
};
void square_laplacian(double* lhs, double* rhs, void* extra_data);                    // extra: laplace_op*
void square_laplacian(complex<double>* lhs, complex<double>* rhs, void* extra_data);  // extra: laplace_op*

// operators.cpp:625   staggered D plus a two-link Laplace term scaled by wilson_coeff (a 13-point stencil)
void square_staggered_2linklaplace_u1(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:688 / :728   symmetric shifts (1/2)(U psi(x+mu) + U^* psi(x-mu)), eta_1 in the y direction
void staggered_symmshift_x(complex<double>* lhs, complex<double>* rhs, void* extra_data);
void staggered_symmshift_y(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// operators.cpp:782   staggered index operator i D_st - m Gamma_5, Gamma_5 = (i/2)(S_x S_y - S_y S_x) of the symmetric shifts
void staggered_index_operator(complex<double>* lhs, complex<double>* rhs, void* extra_data);

#endif
