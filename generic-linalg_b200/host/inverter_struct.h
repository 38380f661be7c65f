// inverter_struct.h -- result record returned BY VALUE by every solver.
// Drop-in for the reference's inverter_struct.h:14-89: same member names, same order, same
// ownership rules (resSqmrhs is owned; deep copy on copy-construction; assignment takes the
// by-value argument's buffer).  Written from the interface description in SURVEY.md section 8b.
#ifndef GLB200_INVERTER_STRUCT_H
#define GLB200_INVERTER_STRUCT_H

#include <iostream>
#include <string>
using namespace std;  // the reference's headers leak this and user code relies on it

struct inversion_info {
  double resSq;       // true |b - A x|^2 from the final extra operator application
  int iter;           // iterations performed
  bool success;       // tolerance reached before max_iter
  std::string name;   // "CG", "BiCGStab-4", "GMRES(20)", ...
  int ops_count;      // number of operator (callback) applications, initial and final included
  double* resSqmrhs;  // per-system residuals of multishift solves, else 0
  int n_rhs;          // length of resSqmrhs, -1 when unused

  inversion_info() : resSq(0.0), iter(0), success(false), name(""), ops_count(0), resSqmrhs(0), n_rhs(-1) {}
  explicit inversion_info(int in_n_rhs)
      : resSq(0.0), iter(0), success(false), name(""), ops_count(0), resSqmrhs(0), n_rhs(in_n_rhs) {
    resSqmrhs = new double[in_n_rhs];
  }
  inversion_info(const inversion_info& o)
      : resSq(o.resSq), iter(o.iter), success(o.success), name(o.name), ops_count(o.ops_count), resSqmrhs(0),
        n_rhs(o.n_rhs) {
    if (o.resSqmrhs != 0 && n_rhs > 0) {
      resSqmrhs = new double[n_rhs];
      for (int i = 0; i < n_rhs; i++) resSqmrhs[i] = o.resSqmrhs[i];
    }
  }
  inversion_info& operator=(inversion_info o) {  // by value: steal the copy's buffer
    resSq = o.resSq;
    iter = o.iter;
    success = o.success;
    name = o.name;
    ops_count = o.ops_count;
    n_rhs = o.n_rhs;
    if (o.resSqmrhs != 0 && n_rhs > 0) {
      resSqmrhs = o.resSqmrhs;
      o.resSqmrhs = 0;
    }
    return *this;
  }
  ~inversion_info() {
    if (resSqmrhs != 0 && n_rhs > 0) {
      delete[] resSqmrhs;
      resSqmrhs = 0;
    }
  }
};

#endif
