// lattice.h -- index <-> coordinate conventions (drop-in subset of lattice/lattice.h:9-200).
// dof index i = site*nc + colour, site index = sum_mu coord[mu] * prod_{nu<mu} dims[nu]
// (x fastest), so for nd = 2: i = (y*X + x)*nc + c.
#ifndef GLB200_LATTICE_H
#define GLB200_LATTICE_H

class Lattice {
 public:
  Lattice(int my_nd, int* my_lattice, int my_nc) : nd(my_nd), nc(my_nc), volume(1) {
    dims = new int[nd];
    for (int mu = 0; mu < nd; mu++) {
      dims[mu] = my_lattice[mu];
      volume *= my_lattice[mu];
    }
    lattice_size = volume * nc;
  }
  Lattice(const Lattice& o) : nd(o.nd), nc(o.nc), volume(o.volume), lattice_size(o.lattice_size) {
    dims = new int[nd];
    for (int mu = 0; mu < nd; mu++) dims[mu] = o.dims[mu];
  }
  ~Lattice() { delete[] dims; }

  inline int coord_to_index(int* coord, int color) {
    int i = 0;
    for (int mu = nd - 1; mu >= 0; mu--) i = i * dims[mu] + coord[mu];
    return i * nc + color;
  }
  inline void coord_to_index(int& i, int* coord, int color) { i = coord_to_index(coord, color); }
  inline void index_to_coord(int i, int* coord, int& color) {
    color = i % nc;
    i = (i - color) / nc;
    for (int mu = 0; mu < nd; mu++) {
      coord[mu] = i % dims[mu];
      i = (i - coord[mu]) / dims[mu];
    }
  }
  inline int index_to_color(int i) { return i % nc; }
  inline bool index_is_even(int& i) {  // true when the coordinate sum is even
    int rest = (i - i % nc) / nc, sum = 0;
    for (int mu = 0; mu < nd; mu++) {
      sum += rest % dims[mu];
      rest /= dims[mu];
    }
    return (sum + 1) % 2;
  }
  inline bool coord_is_even(int* coord) {  // NB: as in the reference this returns sum % 2
    int sum = 0;
    for (int mu = 0; mu < nd; mu++) sum += coord[mu];
    return sum % 2;
  }
  inline void get_lattice(int* out) {
    for (int mu = 0; mu < nd; mu++) out[mu] = dims[mu];
  }
  inline int get_lattice_dimension(int mu) { return (mu >= 0 && mu < nd) ? dims[mu] : -1; }
  inline int get_nd() { return nd; }
  inline int get_nc() { return nc; }
  inline int get_volume() { return volume; }
  inline int get_lattice_size() { return lattice_size; }

 private:
  int nd;
  int* dims;
  int nc;
  int volume;
  int lattice_size;
};

#endif
