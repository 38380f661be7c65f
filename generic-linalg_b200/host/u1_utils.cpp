// u1_utils.cpp -- host-side U(1) gauge-field utilities with the reference's names, file format and arithmetic
// (u1_utils/u1_utils.cpp:14-228; link layout lattice[y*X*2 + x*2 + mu]).  Start-up code around the hot path.
#include "u1_utils.h"

#include <cmath>
#include <fstream>
#include <iostream>
#include <vector>

namespace {

const double kPi = 3.14159265358979323846;  // u1_utils.cpp:11

// periodic 2-D link field addressed by (x, y, mu)
struct Links {
  complex<double>* u;
  int X, Y;
  complex<double>& at(int x, int y, int mu) const { return u[((size_t)y * X + x) * 2 + mu]; }
  int right(int x) const { return (x + 1) % X; }
  int left(int x) const { return (x - 1 + X) % X; }
  int up(int y) const { return (y + 1) % Y; }
  int down(int y) const { return (y - 1 + Y) % Y; }
  int count() const { return 2 * X * Y; }
};

// the oriented plaquette with corner (x, y): U_x(x,y) U_y(x+1,y) conj(U_x(x,y+1)) conj(U_y(x,y))
complex<double> plaquette_at(const Links& L, int x, int y) {
  return L.at(x, y, 0) * L.at(L.right(x), y, 1) * std::conj(L.at(x, L.up(y), 0)) * std::conj(L.at(x, y, 1));
}

}  // namespace

void read_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, string input_file) {
  const Links L = {gauge_field, x_len, y_len};
  std::ifstream in(input_file.c_str());
  if (!in) std::cerr << "[glb200] read_gauge_u1: cannot open " << input_file << " (the field is left undefined, as in the reference)\n";
  double phase = 0.0;
  for (int x = 0; x < x_len; x++)        // the file is x-major (its writer had y as the fast coordinate)
    for (int y = 0; y < y_len; y++)
      for (int mu = 0; mu < 2; mu++) {
        in >> phase;
        L.at(x, y, mu) = std::polar(1.0, phase);
      }
}

void write_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, string output_file) {
  const Links L = {gauge_field, x_len, y_len};
  std::ofstream out(output_file.c_str(), std::ios::out | std::ios::trunc);
  out.setf(std::ios_base::fixed, std::ios_base::floatfield);
  out.precision(20);
  for (int x = 0; x < x_len; x++)
    for (int y = 0; y < y_len; y++)
      for (int mu = 0; mu < 2; mu++) out << std::arg(L.at(x, y, mu)) << "\n";
}

void unit_gauge_u1(complex<double>* gauge_field, int x_len, int y_len) {
  for (int i = 0; i < 2 * y_len * x_len; i++) gauge_field[i] = 1.0;
}

void rand_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, std::mt19937& generator) {
  std::uniform_real_distribution<> phase(-kPi, kPi);
  for (int i = 0; i < 2 * y_len * x_len; i++) gauge_field[i] = std::polar(1.0, phase(generator));
}

void gauss_gauge_u1(complex<double>* gauge_field, int x_len, int y_len, std::mt19937& generator, double beta) {
  if (beta < 0) beta = -beta;
  // u1_utils.cpp:99-102: at beta = 0 the reference draws a uniform field and then carries on (the gaussian draw
  // below then has infinite width); the generator state a caller sees afterwards is reproduced
  if (beta == 0) rand_gauge_u1(gauge_field, x_len, y_len, generator);
  std::normal_distribution<> phase(0.0, 1.0 / sqrt(beta));
  for (int i = 0; i < 2 * y_len * x_len; i++) gauge_field[i] = std::polar(1.0, phase(generator));
}

void rand_trans_u1(complex<double>* gauge_trans, int x_len, int y_len, std::mt19937& generator) {
  std::uniform_real_distribution<> phase(-kPi, kPi);
  for (int i = 0; i < y_len * x_len; i++) gauge_trans[i] = std::polar(1.0, phase(generator));
}

void apply_gauge_trans_u1(complex<double>* gauge_field, complex<double>* gauge_trans, int x_len, int y_len) {
  const Links L = {gauge_field, x_len, y_len};
  for (int y = 0; y < y_len; y++)
    for (int x = 0; x < x_len; x++) {
      const complex<double> g = gauge_trans[(size_t)x_len * y + x];
      L.at(x, y, 0) = g * L.at(x, y, 0) * std::conj(gauge_trans[(size_t)x_len * y + L.right(x)]);
      L.at(x, y, 1) = g * L.at(x, y, 1) * std::conj(gauge_trans[(size_t)x_len * L.up(y) + x]);
    }
}

void apply_ape_smear_u1(complex<double>* smeared_field, complex<double>* gauge_field, int x_len, int y_len, double alpha,
                        int n_iter) {
  std::vector<complex<double> > cur(gauge_field, gauge_field + 2 * (size_t)x_len * y_len);
  const Links S = {cur.data(), x_len, y_len};
  const Links O = {smeared_field, x_len, y_len};
  for (int it = 0; it < n_iter; it++) {
    for (int y = 0; y < y_len; y++)
      for (int x = 0; x < x_len; x++) {
        const int xr = S.right(x), xl = S.left(x), yu = S.up(y), yd = S.down(y);
        // x link: the staple over the plaquette above, then the one below
        O.at(x, y, 0) = S.at(x, y, 0);
        O.at(x, y, 0) += alpha * S.at(x, y, 1) * S.at(x, yu, 0) * std::conj(S.at(xr, y, 1));
        O.at(x, y, 0) += alpha * std::conj(S.at(x, yd, 1)) * S.at(x, yd, 0) * S.at(xr, yd, 1);
        // y link: the staple over the plaquette to the left, then the one to the right
        O.at(x, y, 1) = S.at(x, y, 1);
        O.at(x, y, 1) += alpha * std::conj(S.at(xl, y, 0)) * S.at(xl, y, 1) * S.at(xl, yu, 0);
        O.at(x, y, 1) += alpha * S.at(x, y, 0) * S.at(xr, y, 1) * std::conj(S.at(x, yu, 0));
      }
    for (int i = 0; i < S.count(); i++) cur[i] = std::polar(1.0, std::arg(smeared_field[i]));  // back to U(1)
  }
  for (int i = 0; i < S.count(); i++) smeared_field[i] = cur[i];
}

complex<double> get_plaquette_u1(complex<double>* gauge_field, int x_len, int y_len) {
  const Links L = {gauge_field, x_len, y_len};
  complex<double> sum = 0.0;
  for (int y = 0; y < y_len; y++)
    for (int x = 0; x < x_len; x++) sum += plaquette_at(L, x, y);
  return sum / ((double)(x_len * y_len));
}

double get_topo_u1(complex<double>* gauge_field, int x_len, int y_len) {
  const Links L = {gauge_field, x_len, y_len};
  double angle_sum = 0.0;
  for (int y = 0; y < y_len; y++)
    for (int x = 0; x < x_len; x++) angle_sum += std::arg(plaquette_at(L, x, y));
  return 0.5 * angle_sum / kPi;
}
