// Host-side sign flips on a lattice vector -- the two helpers of the reference's lattice/lattice_functions.h:9-53,
// offered under the same names so that its programs (tests/staggered_pieces) compile here unchanged.  Host utilities,
// not on the accelerated path: the device applies these signs inside the stencil kernels (coarse_sign_kernel, the
// gamma5 / sigma3 post-operations of the partial applies).  Both allow out == in.
#ifndef GLB200_HOST_LATTICE_FUNCTIONS_H
#define GLB200_HOST_LATTICE_FUNCTIONS_H

#include <complex>

#include "lattice.h"

// epsilon(x) = (-1)^(sum of the coordinates): odd sites change sign (lattice_functions.h:11-25)
inline void lattice_epsilon(std::complex<double>* out, std::complex<double>* in, Lattice* latt) {
  const int n = latt->get_lattice_size();
  for (int i = 0; i < n; i++) {
    int idx = i;
    const std::complex<double> v = in[i];
    out[i] = latt->index_is_even(idx) ? v : -v;
  }
}

// sigma_3 on the internal index: the upper half of the colours changes sign; a copy when the colour count is odd
// (lattice_functions.h:29-52)
inline void lattice_sigma3(std::complex<double>* out, std::complex<double>* in, Lattice* latt) {
  const int nc = latt->get_nc(), n = latt->get_volume() * nc;
  const int half = (nc % 2 == 0) ? nc / 2 : nc;  // odd: nothing is flipped
  for (int i = 0; i < n; i++) {
    const std::complex<double> v = in[i];
    out[i] = (i % nc < half) ? v : -v;
  }
}

#endif
