// TEST INFRASTRUCTURE: exercises every routine of u1_utils.h, generic_vector.h and lattice_functions.h on seeded inputs and prints the results
// with 17 digits on stderr; tests/test_reference_programs_cpu.py builds it against the reference's headers / sources and
// against generic-linalg_b200/host and compares the two outputs line by line.
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

using namespace std;  // the reference's lattice_functions.h says complex<double> unqualified, as its programs do

#include "generic_vector.h"
#include "lattice_functions.h"
#include "u1_utils.h"

static void show(const char* what, const std::complex<double>* v, int n) {
  std::complex<double> s = 0.0;
  for (int i = 0; i < n; i++) s += v[i] * (double)(1 + i % 7);
  fprintf(stderr, "%s %.17g %.17g | %.17g %.17g | %.17g %.17g\n", what, s.real(), s.imag(), v[0].real(), v[0].imag(),
          v[n - 1].real(), v[n - 1].imag());
}

int main(int argc, char** argv) {
  const int X = 12, Y = 10, V = X * Y;
  std::mt19937 gen(4242u);
  std::vector<std::complex<double> > U(2 * V), W(2 * V), g(V), a(V), b(V);
  unit_gauge_u1(U.data(), X, Y);
  show("unit", U.data(), 2 * V);
  rand_gauge_u1(U.data(), X, Y, gen);
  show("rand", U.data(), 2 * V);
  gauss_gauge_u1(U.data(), X, Y, gen, 6.0);
  show("gauss", U.data(), 2 * V);
  gauss_gauge_u1(W.data(), X, Y, gen, -2.5);
  show("gauss_negbeta", W.data(), 2 * V);
  std::complex<double> p = get_plaquette_u1(U.data(), X, Y);
  fprintf(stderr, "plaq %.17g %.17g topo %.17g\n", p.real(), p.imag(), get_topo_u1(U.data(), X, Y));
  rand_trans_u1(g.data(), X, Y, gen);
  show("trans", g.data(), V);
  apply_gauge_trans_u1(U.data(), g.data(), X, Y);
  show("transformed", U.data(), 2 * V);
  p = get_plaquette_u1(U.data(), X, Y);
  fprintf(stderr, "plaq_after_trans %.17g %.17g topo %.17g\n", p.real(), p.imag(), get_topo_u1(U.data(), X, Y));
  apply_ape_smear_u1(W.data(), U.data(), X, Y, 0.5, 3);
  show("ape", W.data(), 2 * V);
  if (argc > 1) {  // file round trip through the text format
    write_gauge_u1(W.data(), X, Y, argv[1]);
    std::vector<std::complex<double> > R(2 * V);
    read_gauge_u1(R.data(), X, Y, argv[1]);
    show("reread", R.data(), 2 * V);
  }
  // generic_vector.h
  gaussian<double>(a.data(), V, gen);
  gaussian<double>(b.data(), V, gen);
  show("gaussian", a.data(), V);
  std::complex<double> d = dot<double>(a.data(), b.data(), V);
  fprintf(stderr, "dot %.17g %.17g norm %.17g diff %.17g\n", d.real(), d.imag(), norm2sq<double>(a.data(), V),
          diffnorm2sq<double>(a.data(), b.data(), V));
  orthogonal<double>(a.data(), b.data(), V);
  show("orthogonal", a.data(), V);
  normalize<double>(a.data(), V);
  show("normalize", a.data(), V);
  conj<double>(a.data(), V);
  show("conj", a.data(), V);
  copy<double>(b.data(), a.data(), V);
  zero<double>(a.data(), V);
  show("copy", b.data(), V);
  show("zero", a.data(), V);
  {  // lattice_functions.h: epsilon on a one-colour and sigma3 on a 4- and a 3-colour lattice, in place and out of place
    int dims[2] = {X, Y};
    Lattice l1(2, dims, 1), l4(2, dims, 4), l3(2, dims, 3);
    std::vector<std::complex<double> > w(4 * V), z(4 * V);
    gaussian<double>(w.data(), 4 * V, gen);
    lattice_epsilon(z.data(), w.data(), &l1);
    show("epsilon", z.data(), V);
    lattice_epsilon(z.data(), w.data(), &l4);
    show("epsilon4", z.data(), 4 * V);
    lattice_sigma3(z.data(), w.data(), &l4);
    show("sigma3", z.data(), 4 * V);
    lattice_sigma3(w.data(), w.data(), &l4);
    lattice_sigma3(w.data(), w.data(), &l3);
    lattice_epsilon(w.data(), w.data(), &l3);
    show("inplace", w.data(), 3 * V);
  }
  std::vector<double> r(V), q(V);
  gaussian<double>(r.data(), V, gen);
  gaussian<double>(q.data(), V, gen);
  orthogonal<double>(r.data(), q.data(), V);
  normalize<double>(r.data(), V);
  fprintf(stderr, "real %.17g %.17g %.17g %.17g\n", dot<double>(r.data(), q.data(), V), norm2sq<double>(r.data(), V),
          diffnorm2sq<double>(r.data(), q.data(), V), r[V - 1]);
  return 0;
}
