// u1_utils.h -- 2-D U(1) gauge fields on the host, with the names of the reference's u1_utils/u1_utils.h so that its
// driver programs compile against this directory alone.
//
// A gauge field here is what a driver hands to staggered_u1_op::lattice: 2*x_len*y_len unit-modulus complex numbers,
// link mu (0 = x, 1 = y) of site (x, y) at index y*x_len*2 + x*2 + mu.  Everything in this header runs once at
// start-up on the host (the operators upload the field when they are created); nothing of it is on the accelerated
// path.  Arithmetic and random-number consumption follow u1_utils.cpp statement by statement, so a seeded generator
// produces the reference's field bit for bit (tests/test_reference_programs_cpu.py).
#ifndef GLB200_U1_UTILS_H
#define GLB200_U1_UTILS_H

#include <complex>
#include <random>
#include <string>
using std::complex;
using std::string;

// ---- observables -----------------------------------------------------------------------------------------------
// Mean plaquette, (1/V) sum over sites of U_x(x,y) U_y(x+1,y) conj(U_x(x,y+1)) conj(U_y(x,y))   (u1_utils.cpp:198)
complex<double> get_plaquette_u1(complex<double>* gauge_field, int x_len, int y_len);
// Geometric topological charge: the plaquette angles summed and divided by 2 pi                   (u1_utils.cpp:214)
double get_topo_u1(complex<double>* gauge_field, int x_len, int y_len);

// ---- field generators ------------------------------------------------------------------------------------------
// how a driver obtains its field (u1_utils.h:12-17)
enum gauge_create_type { GAUGE_LOAD = 0, GAUGE_RANDOM = 1, GAUGE_UNIT = 2 };
// cold start: every link 1
void unit_gauge_u1(complex<double>* gauge_field, int x_len, int y_len);
// hot start: one phase per link, uniform in (-pi, pi); std::mt19937 generator(seed)
void rand_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, std::mt19937& generator);
// non-compact gaussian field: phases normal with variance 1/|beta| (beta -> infinity is the cold start).  At beta = 0
// a uniform field is drawn first and the gaussian loop still runs, exactly as the reference does.
void gauss_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, std::mt19937& generator, double beta);

// ---- transformations -------------------------------------------------------------------------------------------
// a random gauge transformation g(x): one uniform phase per site
void rand_trans_u1(complex<double>* gauge_trans, int x_len, int y_len, std::mt19937& generator);
// U_mu(x) <- g(x) U_mu(x) conj(g(x + mu))
void apply_gauge_trans_u1(complex<double>* gauge_field, complex<double>* gauge_trans, int x_len, int y_len);
// APE smearing: n_iter sweeps of link <- link + alpha*(upper staple + lower staple), projected back onto U(1) after
// every sweep; smeared_field receives the result (it may not alias gauge_field)
void apply_ape_smear_u1(complex<double>* smeared_field, complex<double>* gauge_field, int x_len, int y_len, double alpha,
                        int n_iter);

// ---- files -----------------------------------------------------------------------------------------------------
// Text format of the reference's configurations: one phase per line, x slowest, then y, then mu -- the transpose of
// the in-memory order (u1_utils.cpp:14-33, :38-60).  Reading builds exp(i phase); writing prints arg(link) with 20
// fixed digits.
void read_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, string input_file);
void write_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, string output_file);

#endif
