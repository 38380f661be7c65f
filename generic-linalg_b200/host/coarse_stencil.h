// coarse_stencil.h -- the data-driven (up to two-link) stencil operator: drop-in for the parts of
// stencil_2d/coarse_stencil.h:14-154 on the solver hot path.
//   clover   [c + nc*i]                      nc x nc per site (row-major in the dof index i)
//   hopping  [c + nc*i + dir*nc*L]           dir = +x, +y, -x, -y ; L = lattice_size
//   two_link [c + nc*i + dir*nc*L]           dir = +2x, +x+y, +2y, -x+y, -2x, -x-y, -2y, +x-y
// apply_stencil_2d computes lhs = clover + hopping (+ two_link) + shift + eo_shift + dof_shift terms
// (stencil_2d/coarse_stencil.cpp:29-172).  sdir != DIR_ALL applies one direction only (:173-393; the reference's
// set-up probes with it, no solver uses it): served by the same kernels on a copy that keeps that one plane.
#ifndef GLB200_COARSE_STENCIL_H
#define GLB200_COARSE_STENCIL_H

#include <complex>

#include "lattice.h"
using namespace std;

enum stencil_dir {
  DIR_ALL = 0, DIR_0 = 1, DIR_XP1 = 2, DIR_YP1 = 3, DIR_XM1 = 4, DIR_YM1 = 5, DIR_XP2 = 6, DIR_XP1YP1 = 7,
  DIR_YP2 = 8, DIR_XM1YP1 = 9, DIR_XM2 = 10, DIR_XM1YM1 = 11, DIR_YM2 = 12, DIR_XP1YM1 = 13,
};

struct stencil_2d {
  Lattice* lat;
  stencil_dir sdir;
  complex<double>* clover;
  complex<double>* hopping;
  bool has_two;
  int stencil_size;
  complex<double>* two_link;
  bool generated;
  complex<double> shift;      // lhs += shift * rhs
  complex<double> eo_shift;   // lhs += (+-1 by site parity) * eo_shift * rhs
  complex<double> dof_shift;  // lhs += (+1 top half of the colours, -1 bottom half) * dof_shift * rhs

  stencil_2d(Lattice* in_lat, int in_stencil_size, complex<double> in_shift = 0.0, complex<double> in_eo_shift = 0.0,
             complex<double> in_dof_shift = 0.0)
      : lat(in_lat), sdir(DIR_ALL), has_two(in_stencil_size == 2), stencil_size(in_stencil_size), two_link(0),
        generated(false), shift(in_shift), eo_shift(in_eo_shift), dof_shift(in_dof_shift) {
    const size_t m = (size_t)lat->get_volume() * lat->get_nc() * lat->get_nc();
    clover = new complex<double>[m]();
    hopping = new complex<double>[4 * m]();
    if (has_two) two_link = new complex<double>[8 * m]();
  }
  stencil_2d(const stencil_2d& o)
      : lat(o.lat), sdir(o.sdir), has_two(o.has_two), stencil_size(o.stencil_size), two_link(0),
        generated(o.generated), shift(o.shift), eo_shift(o.eo_shift), dof_shift(o.dof_shift) {
    const size_t m = (size_t)lat->get_volume() * lat->get_nc() * lat->get_nc();
    clover = new complex<double>[m];
    hopping = new complex<double>[4 * m];
    for (size_t i = 0; i < m; i++) clover[i] = o.clover[i];
    for (size_t i = 0; i < 4 * m; i++) hopping[i] = o.hopping[i];
    if (has_two) {
      two_link = new complex<double>[8 * m];
      for (size_t i = 0; i < 8 * m; i++) two_link[i] = o.two_link[i];
    }
  }
  ~stencil_2d() {
    delete[] clover;
    delete[] hopping;
    if (has_two) delete[] two_link;
  }
  void clear_stencils() {
    const size_t m = (size_t)lat->get_volume() * lat->get_nc() * lat->get_nc();
    for (size_t i = 0; i < m; i++) clover[i] = 0.0;
    for (size_t i = 0; i < 4 * m; i++) hopping[i] = 0.0;
    if (has_two)
      for (size_t i = 0; i < 8 * m; i++) two_link[i] = 0.0;
    generated = false;
  }

 private:
  stencil_2d& operator=(const stencil_2d&);
};

// stencil_2d/coarse_stencil.cpp:12   extra_data: stencil_2d*
void apply_stencil_2d(complex<double>* lhs, complex<double>* rhs, void* extra_data);
// coarse_stencil.cpp:395 / :560 / :725 / :1120 (sdir == DIR_ALL): hopping term between the site parities, and the
// clover + hopping (+ two-link) blocks between the halves of the colour index; extra_data: stencil_2d*
void apply_stencil_2d_eo(complex<double>* lhs, complex<double>* rhs, void* extra_data);
void apply_stencil_2d_oe(complex<double>* lhs, complex<double>* rhs, void* extra_data);
void apply_stencil_2d_tb(complex<double>* lhs, complex<double>* rhs, void* extra_data);
void apply_stencil_2d_bt(complex<double>* lhs, complex<double>* rhs, void* extra_data);

// coarse_stencil.cpp:1515-1640: fill an allocated, not yet generated stencil (stencil_size 1 or 2) from ANY operator
// callback by probing it.  The reference applies the operator once per lattice dof; here unit sources that are further
// apart than the stencil reaches share one apply (a comb with period >= 2*stencil_size+1 in each direction, when such a
// period divides the lattice -- otherwise one source per row / column as in the reference), which gives the same
// entries bit for bit with nc * period_x * period_y applies.  Host vectors; the callback is one of this library's
// (each apply then runs on the device) or the caller's own.
void generate_stencil_2d(stencil_2d* stenc, void (*matrix_vector)(complex<double>*, complex<double>*, void*),
                         void* extra_data);

#endif
