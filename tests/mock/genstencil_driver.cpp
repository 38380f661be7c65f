// TEST INFRASTRUCTURE: generate_stencil_2d (coarse_stencil.h) on several operators and lattices; prints every entry of
// the generated stencil with 17 digits on stderr.  Built against the reference and against host/ by
// tests/test_reference_programs_cpu.py.
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "coarse_stencil.h"
#include "generic_vector.h"
#include "lattice.h"
#include "operators.h"
#include "u1_utils.h"

static void dump(const char* tag, stencil_2d& st) {
  const int m = st.lat->get_volume() * st.lat->get_nc() * st.lat->get_nc();
  fprintf(stderr, "%s generated %d\n", tag, (int)st.generated);
  for (int i = 0; i < m; i++) fprintf(stderr, "c %.17g %.17g\n", st.clover[i].real(), st.clover[i].imag());
  for (int i = 0; i < 4 * m; i++) fprintf(stderr, "h %.17g %.17g\n", st.hopping[i].real(), st.hopping[i].imag());
  if (st.has_two)
    for (int i = 0; i < 8 * m; i++) fprintf(stderr, "t %.17g %.17g\n", st.two_link[i].real(), st.two_link[i].imag());
}

int main() {
  const int sizes[4][2] = {{8, 6}, {10, 10}, {7, 5}, {4, 2}};
  for (int k = 0; k < 4; k++) {
    const int X = sizes[k][0], Y = sizes[k][1];
    std::mt19937 gen(100u + k);
    std::vector<std::complex<double> > U(2 * X * Y);
    gauss_gauge_u1(U.data(), X, Y, gen, 4.0);
    staggered_u1_op stagif;
    stagif.lattice = U.data();
    stagif.mass = 0.1;
    stagif.x_fine = X;
    stagif.y_fine = Y;
    stagif.Nc = 1;
    stagif.wilson_coeff = 0.25;
    int dims[2] = {X, Y};
    Lattice lat(2, dims, 1);
    stencil_2d one(&lat, 1);
    generate_stencil_2d(&one, square_staggered_u1, (void*)&stagif);
    dump("staggered", one);
    stencil_2d lap(&lat, 1);
    generate_stencil_2d(&lap, square_laplace_u1, (void*)&stagif);
    dump("laplace_u1", lap);
    if (X >= 5 && Y >= 5) {  // on a narrower lattice several of the 13 terms land on one site and the two-link operator's
      stencil_2d twol(&lat, 2);  // own summation order (1 ulp, see DESIGN) would show: not what is tested here
      generate_stencil_2d(&twol, square_staggered_2linklaplace_u1, (void*)&stagif);
      dump("two_link", twol);
    }
    stencil_2d nrm(&lat, 2);
    generate_stencil_2d(&nrm, square_staggered_normal_u1, (void*)&stagif);
    dump("normal", nrm);
  }
  return 0;
}
