// port_oracle.cpp -- CPU restatement ("port") of the reference's solver hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h): the product never links,
// loads or calls this file.  It restates, in our own words, the arithmetic of the
// reference operators, BLAS-1 and Krylov solvers, keeping the exact evaluation
// order (serial left-to-right sums, std::complex operator semantics, no FMA under
// plain -O2 on x86-64) so that it agrees BIT-FOR-BIT with oracle/_ref/libref_oracle.so
// (the unmodified reference sources).  That agreement, plus the golden vectors in
// tests/golden/, is what pins this oracle (tests/test_oracle_cpu.py).
//
// Every routine cites the reference file:line it follows (paths relative to
// the reference root).
#define ORC_PREFIX port_
#include "oracle_api.h"

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

typedef std::complex<double> cplx;

namespace {

// ------------------------------------------------------------------------------------------
// BLAS-1  (generic_vector.h:87-169) -- serial accumulation in index order.
// ------------------------------------------------------------------------------------------
inline double conj_of(double a) { return a; }
inline cplx conj_of(const cplx& a) { return std::conj(a); }
inline double real_of(double a) { return a; }
inline double real_of(const cplx& a) { return a.real(); }
inline double abs_of(double a) { return a; }  // generic_cg_m.cpp:166 (real overload uses zeta itself)
inline double abs_of(const cplx& a) { return std::abs(a); }

template <typename T>
T v_dot(const T* a, const T* b, int n) {  // generic_vector.h:87,102 : sum conj(a_i) b_i
  T s = T(0.0);
  for (int i = 0; i < n; i++) s = s + conj_of(a[i]) * b[i];
  return s;
}
template <typename T>
double v_norm2(const T* a, int n) {  // generic_vector.h:117,132
  double s = 0.0;
  for (int i = 0; i < n; i++) s = s + real_of(conj_of(a[i]) * a[i]);
  return s;
}
template <typename T>
double v_diffnorm2(const T* a, const T* b, int n) {  // generic_vector.h:148,160
  double s = 0.0;
  for (int i = 0; i < n; i++) s = s + real_of(conj_of(a[i] - b[i]) * (a[i] - b[i]));
  return s;
}
template <typename T>
void v_zero(T* a, int n) {
  for (int i = 0; i < n; i++) a[i] = 0.0;
}
template <typename T>
void v_copy(T* dst, const T* src, int n) {
  for (int i = 0; i < n; i++) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// Operators
// ------------------------------------------------------------------------------------------
struct PortOp {
  orc_op_desc d;
  bool is_complex;
  int size;
  int nc;
  // stencil storage (owned copies, reference layout coarse_stencil.h:33-60)
  std::vector<cplx> clover, hopping, two_link;
  cplx shift, eo_shift, dof_shift;
  bool has_two;
};

inline int wrap_up(int a, int L) { return (a + 1) % L; }
inline int wrap_dn(int a, int L) { return (a + L - 1) % L; }

// square_laplace.cpp:182-222 (T=double, diag = 4+m2 with an int 4), imag_laplace.cpp:126-166
// (T=cplx, diag = 4.0+m2+i), operators.cpp:28-66 and tests/multishift/multishift.cpp:634-672
// (Nc colours, diag = 4+mass).  Order: -(x+1) -(x-1) -(y+1) -(y-1) + diag*self.
template <typename T, typename D>
void op_laplace(T* out, const T* in, int X, int Y, int Nc, D diag) {
  const int n = X * Y * Nc;
  for (int i = 0; i < n; i++) {
    const int c = i % Nc;
    const int s = (i - c) / Nc;
    const int x = s % X, y = s / X;
    T acc = 0.0;
    acc = acc - in[y * X * Nc + wrap_up(x, X) * Nc + c];
    acc = acc - in[y * X * Nc + wrap_dn(x, X) * Nc + c];
    acc = acc - in[wrap_up(y, Y) * X * Nc + x * Nc + c];
    acc = acc - in[wrap_dn(y, Y) * X * Nc + x * Nc + c];
    acc = acc + diag * in[i];
    out[i] = acc;
  }
}

// operators.cpp:73-120 square_laplace_u1
void op_laplace_u1(cplx* out, const cplx* in, const cplx* U, int X, int Y, double mass) {
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    const int xp = wrap_up(x, X), xm = wrap_dn(x, X), yp = wrap_up(y, Y), ym = wrap_dn(y, Y);
    cplx acc = 0.0;
    acc = acc - U[y * X * 2 + x * 2] * in[y * X + xp];
    acc = acc - std::conj(U[y * X * 2 + xm * 2]) * in[y * X + xm];
    acc = acc - U[y * X * 2 + x * 2 + 1] * in[yp * X + x];
    acc = acc - std::conj(U[ym * X * 2 + x * 2 + 1]) * in[ym * X + x];
    acc = acc + (4 + mass) * in[i];
    out[i] = acc;
  }
}

// Staggered family.  hop = -1 for D (operators.cpp:184-239), +1 for D^dagger (operators.cpp:372-441);
// g5 multiplies every term by eps(x,y)=(-1)^(x+y) in the positions operators.cpp:316-368 does.
// U == nullptr is the free operator (operators.cpp:127-176, 262-312).
void op_staggered(cplx* out, const cplx* in, const cplx* U, int X, int Y, double mass, bool dagger, bool g5) {
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    const int xp = wrap_up(x, X), xm = wrap_dn(x, X), yp = wrap_up(y, Y), ym = wrap_dn(y, Y);
    const double eta1 = 1 - 2 * (x % 2);
    cplx acc = 0.0;
    if (!g5) {
      if (U) {
        if (!dagger) {
          acc = acc - U[y * X * 2 + x * 2] * in[y * X + xp];
          acc = acc + std::conj(U[y * X * 2 + xm * 2]) * in[y * X + xm];
          acc = acc - eta1 * U[y * X * 2 + x * 2 + 1] * in[yp * X + x];
          acc = acc + eta1 * std::conj(U[ym * X * 2 + x * 2 + 1]) * in[ym * X + x];
        } else {
          acc = acc + U[y * X * 2 + x * 2] * in[y * X + xp];
          acc = acc - std::conj(U[y * X * 2 + xm * 2]) * in[y * X + xm];
          acc = acc + eta1 * U[y * X * 2 + x * 2 + 1] * in[yp * X + x];
          acc = acc - eta1 * std::conj(U[ym * X * 2 + x * 2 + 1]) * in[ym * X + x];
        }
      } else {
        acc = acc - in[y * X + xp];
        acc = acc + in[y * X + xm];
        acc = acc - eta1 * in[yp * X + x];
        acc = acc + eta1 * in[ym * X + x];
      }
      acc = 0.5 * acc;
      acc = acc + mass * in[i];
    } else {
      const double eo = ((x + y) % 2 == 0) ? 1.0 : -1.0;
      if (U) {
        acc = acc - eo * U[y * X * 2 + x * 2] * in[y * X + xp];
        acc = acc + eo * std::conj(U[y * X * 2 + xm * 2]) * in[y * X + xm];
        acc = acc - eo * eta1 * U[y * X * 2 + x * 2 + 1] * in[yp * X + x];
        acc = acc + eo * eta1 * std::conj(U[ym * X * 2 + x * 2 + 1]) * in[ym * X + x];
      } else {
        acc = acc - eo * in[y * X + xp];
        acc = acc + eo * in[y * X + xm];
        acc = acc - eo * eta1 * in[yp * X + x];
        acc = acc + eo * eta1 * in[ym * X + x];
      }
      acc = 0.5 * acc;
      acc = acc + eo * mass * in[i];
    }
    out[i] = acc;
  }
}

// tests/multishift/multishift.cpp:677-726 (real free staggered)
void op_staggered_free_real(double* out, const double* in, int X, int Y, double mass) {
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    const double eta1 = 1 - 2 * (x % 2);
    double acc = 0.0;
    acc = acc - in[y * X + wrap_up(x, X)];
    acc = acc + in[y * X + wrap_dn(x, X)];
    acc = acc - eta1 * in[wrap_up(y, Y) * X + x];
    acc = acc + eta1 * in[wrap_dn(y, Y) * X + x];
    acc = 0.5 * acc;
    acc = acc + mass * in[i];
    out[i] = acc;
  }
}

// operators.cpp:456-525 square_staggered_deo_u1 / _doe_u1: the hopping term on one parity, the other zeroed
void op_staggered_eo(cplx* out, const cplx* in, const cplx* U, int X, int Y, int parity) {
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    out[i] = 0.0;
    const double eta1 = 1 - 2 * (x % 2);
    if ((x + y) % 2 != parity) continue;
    const int xp = wrap_up(x, X), xm = wrap_dn(x, X), yp = wrap_up(y, Y), ym = wrap_dn(y, Y);
    out[i] = out[i] - U[y * X * 2 + x * 2] * in[y * X + xp];
    out[i] = out[i] + std::conj(U[y * X * 2 + xm * 2]) * in[y * X + xm];
    out[i] = out[i] - eta1 * U[y * X * 2 + x * 2 + 1] * in[yp * X + x];
    out[i] = out[i] + eta1 * std::conj(U[ym * X * 2 + x * 2 + 1]) * in[ym * X + x];
    out[i] = 0.5 * out[i];
  }
}
// operators.cpp:549-571 square_staggered_m2mdeodoe_u1
void op_staggered_m2mdeodoe(cplx* out, const cplx* in, const cplx* U, int X, int Y, double mass) {
  std::vector<cplx> tmp((size_t)X * Y);
  op_staggered_eo(tmp.data(), in, U, X, Y, 1);
  op_staggered_eo(out, tmp.data(), U, X, Y, 0);
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    if ((x + y) % 2 == 0) out[i] = mass * mass * in[i] - out[i];
  }
}
// operators.cpp:528-545 / :574-598
void op_eoprec_prepare(cplx* rhs_e, const cplx* rhs_orig, const cplx* U, int X, int Y, double mass) {
  op_staggered_eo(rhs_e, rhs_orig, U, X, Y, 0);
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    if ((x + y) % 2 == 0) rhs_e[i] = mass * rhs_orig[i] - rhs_e[i];
  }
}
void op_eoprec_reconstruct(cplx* lhs_full, const cplx* lhs_e, const cplx* rhs_o, const cplx* U, int X, int Y, double mass) {
  const double inv_mass = 1.0 / mass;
  op_staggered_eo(lhs_full, lhs_e, U, X, Y, 1);
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    if ((x + y) % 2 == 1)
      lhs_full[i] = inv_mass * (rhs_o[i] - lhs_full[i]);
    else
      lhs_full[i] = lhs_e[i];
  }
}

// operators.cpp:242-259 gamma_5
void op_gamma5(cplx* out, const cplx* in, int X, int Y) {
  for (int i = 0; i < X * Y; i++) {
    const int x = i % X, y = i / X;
    out[i] = ((double)(1 - 2 * ((x + y) % 2))) * in[i];
  }
}

// stencil_2d/coarse_stencil.cpp:29-172  apply_stencil_2d, DIR_ALL.
// dof index i = site*nc + row, site = y*X + x (lattice.h:52-66).  Matrices: clover[c + nc*i],
// hopping[c + nc*i + dir*nc*L] with L = V*nc, dir order +x,+y,-x,-y; two_link dir order
// +2x, +x+y, +2y, -x+y, -2x, -x-y, -2y, +x-y.  Accumulation order exactly as the reference.
void op_stencil(const PortOp& op, cplx* out, const cplx* in) {
  const int X = op.d.X, Y = op.d.Y, nc = op.nc;
  const int L = X * Y * nc;
  static const int hop_dx[4] = {1, 0, -1, 0}, hop_dy[4] = {0, 1, 0, -1};
  static const int two_dx[8] = {2, 1, 0, -1, -2, -1, 0, 1}, two_dy[8] = {0, 1, 2, 1, 0, -1, -2, -1};
  for (int i = 0; i < L; i++) {
    const int row = i % nc;
    const int site = (i - row) / nc;
    const int x = site % X, y = (site - x) / X;
    cplx acc = 0.0;
    for (int c = 0; c < nc; c++) acc += op.clover[c + nc * i] * in[site * nc + c];
    for (int dir = 0; dir < 4; dir++) {
      const int xn = (x + hop_dx[dir] + X) % X, yn = (y + hop_dy[dir] + Y) % Y;
      const int nbr = (yn * X + xn) * nc;
      for (int c = 0; c < nc; c++) acc += op.hopping[c + nc * i + dir * nc * L] * in[nbr + c];
    }
    if (op.has_two) {
      for (int dir = 0; dir < 8; dir++) {
        const int xn = (x + two_dx[dir] + 2 * X) % X, yn = (y + two_dy[dir] + 2 * Y) % Y;
        const int nbr = (yn * X + xn) * nc;
        for (int c = 0; c < nc; c++) acc += op.two_link[c + nc * i + dir * nc * L] * in[nbr + c];
      }
    }
    if (std::abs(op.shift) != 0.0) acc += op.shift * in[i];
    if (std::abs(op.eo_shift) != 0.0) acc += (((x + y) % 2 == 0) ? 1.0 : -1.0) * op.eo_shift * in[i];
    if (std::abs(op.dof_shift) != 0.0) acc += (row < nc / 2 ? 1.0 : -1.0) * op.dof_shift * in[i];
    out[i] = acc;
  }
}

// operator_utils/operators_stencil.cpp:14-63  get_square_staggered_u1_stencil (nc = 1)
void build_stag_stencil(PortOp& op) {
  const int X = op.d.X, Y = op.d.Y, V = X * Y;
  const cplx* U = (const cplx*)op.d.links;
  op.nc = 1;
  op.clover.assign(V, cplx(0.0));
  op.hopping.assign(4 * (size_t)V, cplx(0.0));
  op.has_two = false;
  for (int i = 0; i < V; i++) {
    const int x = i % X, y = i / X;
    const int eta1 = 1 - 2 * (x % 2);
    op.hopping[i] = -0.5 * U[2 * i];
    op.hopping[i + V] = -0.5 * eta1 * U[2 * i + 1];
    op.hopping[i + 2 * V] = 0.5 * std::conj(U[2 * (y * X + wrap_dn(x, X))]);
    op.hopping[i + 3 * V] = 0.5 * eta1 * std::conj(U[2 * (wrap_dn(y, Y) * X + x) + 1]);
  }
  op.shift = op.d.mass;
  op.eo_shift = 0.0;
  op.dof_shift = 0.0;
}

void apply_c(PortOp* op, cplx* out, const cplx* in) {
  const orc_op_desc& d = op->d;
  const cplx* U = (const cplx*)d.links;
  switch (d.kind) {
    case ORC_OP_LAPLACE_IMAG: op_laplace<cplx, cplx>(out, in, d.X, d.Y, 1, 4.0 + d.mass + cplx(0.0, 1.0)); break;
    case ORC_OP_LAPLACE_NC: op_laplace<cplx, double>(out, in, d.X, d.Y, op->nc, 4 + d.mass); break;
    case ORC_OP_LAPLACE_U1: op_laplace_u1(out, in, U, d.X, d.Y, d.mass); break;
    case ORC_OP_STAG_FREE: op_staggered(out, in, nullptr, d.X, d.Y, d.mass, false, false); break;
    case ORC_OP_STAG_U1: op_staggered(out, in, U, d.X, d.Y, d.mass, false, false); break;
    case ORC_OP_STAG_GAMMA5_U1: op_staggered(out, in, U, d.X, d.Y, d.mass, false, true); break;
    case ORC_OP_STAG_GAMMA5_FREE: op_staggered(out, in, nullptr, d.X, d.Y, d.mass, false, true); break;
    case ORC_OP_STAG_DAGGER_U1: op_staggered(out, in, U, d.X, d.Y, d.mass, true, false); break;
    case ORC_OP_STAG_NORMAL_U1: {  // operators.cpp:444-453 : tmp = D in ; out = D^dag tmp
      std::vector<cplx> tmp((size_t)d.X * d.Y);
      op_staggered(tmp.data(), in, U, d.X, d.Y, d.mass, false, false);
      op_staggered(out, tmp.data(), U, d.X, d.Y, d.mass, true, false);
      break;
    }
    case ORC_OP_GAMMA5: op_gamma5(out, in, d.X, d.Y); break;
    case ORC_OP_STENCIL:
    case ORC_OP_STENCIL_FROM_STAG: op_stencil(*op, out, in); break;
    case ORC_OP_STAG_DEO_U1: op_staggered_eo(out, in, U, d.X, d.Y, 0); break;
    case ORC_OP_STAG_DOE_U1: op_staggered_eo(out, in, U, d.X, d.Y, 1); break;
    case ORC_OP_STAG_M2MDEODOE_U1: op_staggered_m2mdeodoe(out, in, U, d.X, d.Y, d.mass); break;
    default: break;
  }
}
void apply_r(PortOp* op, double* out, const double* in) {
  const orc_op_desc& d = op->d;
  switch (d.kind) {
    case ORC_OP_LAPLACE_REAL: op_laplace<double, double>(out, in, d.X, d.Y, 1, 4 + d.mass); break;
    case ORC_OP_LAPLACE_REAL_NC: op_laplace<double, double>(out, in, d.X, d.Y, op->nc, 4 + d.mass); break;
    case ORC_OP_STAG_FREE_REAL: op_staggered_free_real(out, in, d.X, d.Y, d.mass); break;
    default: break;
  }
}
inline void apply(PortOp* op, cplx* out, const cplx* in) { apply_c(op, out, in); }
inline void apply(PortOp* op, double* out, const double* in) { apply_r(op, out, in); }

// ------------------------------------------------------------------------------------------
// Reporting (verbosity.h:73-133), result struct (inverter_struct.h:14-33)
// ------------------------------------------------------------------------------------------
struct Info {
  double resSq = 0.0;
  int iter = 0;
  bool success = false;
  std::string name;
  int ops = 0;
  std::vector<double> res_multi;
};

struct Verb {
  int level;
  std::string prefix;
};
inline void say_resid(const Verb* v, const std::string& alg, int iter, int ops, double rel) {
  if (v && v->level == 3) std::cout << v->prefix << alg << " Iter " << iter << " Ops " << ops << " RelRes " << rel << "\n";
}
inline void say_summary(const Verb* v, const std::string& alg, bool ok, int iter, int ops, double rel) {
  if (v && v->level >= 1)
    std::cout << v->prefix << alg << " Success " << (ok ? "Y" : "N") << " Iter " << iter << " Ops " << ops
              << " RelRes " << rel << "\n";
}
inline void say_restart(const Verb* v, const std::string& alg, int iter, int ops, double rel) {
  if (v && v->level >= 2) std::cout << v->prefix << alg << " Iter " << iter << " Ops " << ops << " RelRes " << rel << "\n";
}
// verbosity.h:27-44 : the inner solver of a restarted run is silent unless DETAIL was asked for.
inline Verb inner_verb(const Verb* v) {
  Verb w;
  w.level = (!v || v->level == 1 || v->level == 2) ? 0 : v->level;
  w.prefix = v ? v->prefix : "";
  return w;
}

// ------------------------------------------------------------------------------------------
// Solvers
// ------------------------------------------------------------------------------------------

// generic_cg.cpp:132-231 (double), :278-377 (complex)
template <typename T>
Info solve_cg(PortOp* A, T* x, const T* b, int n, int max_iter, double eps, const Verb* verb) {
  Info inf;
  std::vector<T> r(n, T(0.0)), p(n, T(0.0)), Ap(n, T(0.0));
  const double bnorm = std::sqrt(v_norm2(b, n));
  apply(A, p.data(), x);
  inf.ops++;
  for (int i = 0; i < n; i++) r[i] = b[i] - p[i];
  v_copy(p.data(), r.data(), n);
  apply(A, Ap.data(), p.data());
  inf.ops++;
  double rsq = v_norm2(r.data(), n), rsq_new = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T alpha = rsq / v_dot(p.data(), Ap.data(), n);
    for (int i = 0; i < n; i++) {
      x[i] = x[i] + alpha * p[i];
      r[i] = r[i] - alpha * Ap[i];
    }
    rsq_new = v_norm2(r.data(), n);
    say_resid(verb, "CG", k + 1, inf.ops, std::sqrt(rsq_new) / bnorm);
    if (std::sqrt(rsq_new) < eps * bnorm || k == max_iter - 1) break;
    const T beta = rsq_new / rsq;
    rsq = rsq_new;
    for (int i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
    apply(A, Ap.data(), p.data());
    inf.ops++;
  }
  inf.success = !(k == max_iter - 1);
  k++;
  apply(A, Ap.data(), x);
  inf.ops++;
  const double truersq = v_diffnorm2(Ap.data(), b, n);
  say_summary(verb, "CG", inf.success, k, inf.ops, std::sqrt(truersq) / bnorm);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "CG";
  return inf;
}

// generic_cr.cpp:28-150 (double), :198-322 (complex).  The complex overload tests k==max_iter
// for failure and therefore never reports it (generic_cr.cpp:288 vs :117).
template <typename T>
Info solve_cr(PortOp* A, T* phi, const T* b, int n, int max_iter, double eps, const Verb* verb, bool is_complex) {
  Info inf;
  std::vector<T> x(n), r(n, T(0.0)), Ar(n, T(0.0)), p(n, T(0.0)), Ap(n, T(0.0));
  v_copy(x.data(), phi, n);
  const double bnorm = std::sqrt(v_norm2(b, n));
  apply(A, p.data(), x.data());
  inf.ops++;
  for (int i = 0; i < n; i++) r[i] = b[i] - p[i];
  v_copy(p.data(), r.data(), n);
  apply(A, Ap.data(), p.data());
  inf.ops++;
  v_copy(Ar.data(), Ap.data(), n);
  double Apsq = v_norm2(Ap.data(), n), rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T alpha = v_dot(Ap.data(), r.data(), n) / Apsq;
    for (int i = 0; i < n; i++) {
      x[i] = x[i] + alpha * p[i];
      r[i] = r[i] - alpha * Ap[i];
    }
    rsq = v_norm2(r.data(), n);
    say_resid(verb, "CR", k + 1, inf.ops, std::sqrt(rsq) / bnorm);
    if (std::sqrt(rsq) < eps * bnorm || k == max_iter - 1) break;
    v_zero(Ar.data(), n);
    apply(A, Ar.data(), r.data());
    inf.ops++;
    const T beta = -v_dot(Ap.data(), Ar.data(), n) / Apsq;
    for (int i = 0; i < n; i++) {
      p[i] = r[i] + beta * p[i];
      Ap[i] = Ar[i] + beta * Ap[i];
    }
    Apsq = v_norm2(Ap.data(), n);
  }
  inf.success = !(k == (is_complex ? max_iter : max_iter - 1));
  k++;
  v_zero(p.data(), n);
  apply(A, p.data(), x.data());
  inf.ops++;
  double truersq = 0.0;
  for (int i = 0; i < n; i++) truersq += real_of(conj_of(p[i] - b[i]) * (p[i] - b[i]));
  v_copy(phi, x.data(), n);
  say_summary(verb, "CR", inf.success, k, inf.ops, std::sqrt(truersq) / bnorm);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "CR";
  return inf;
}

// generic_gcr.cpp:27-156 (double), :202-333 (complex)
template <typename T>
Info solve_gcr(PortOp* A, T* phi, const T* b, int n, int max_iter, double eps, const Verb* verb) {
  Info inf;
  std::vector<T> x(n), r(n, T(0.0)), Ar(n, T(0.0));
  std::vector<std::vector<T> > ps, Aps;  // all search directions are kept (explicit re-orthogonalisation)
  ps.emplace_back(n, T(0.0));
  Aps.emplace_back(n, T(0.0));
  v_copy(x.data(), phi, n);
  const double bnorm = std::sqrt(v_norm2(b, n));
  apply(A, ps[0].data(), x.data());
  inf.ops++;
  for (int i = 0; i < n; i++) r[i] = b[i] - ps[0][i];
  v_copy(ps[0].data(), r.data(), n);
  apply(A, Aps[0].data(), ps[0].data());
  inf.ops++;
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T* p = ps[k].data();
    const T* Ap = Aps[k].data();
    const T alpha = v_dot(Ap, r.data(), n) / v_norm2(Ap, n);
    for (int i = 0; i < n; i++) {
      x[i] = x[i] + alpha * p[i];
      r[i] = r[i] - alpha * Ap[i];
    }
    rsq = v_norm2(r.data(), n);
    say_resid(verb, "GCR", k + 1, inf.ops, std::sqrt(rsq) / bnorm);
    if (std::sqrt(rsq) < eps * bnorm || k == max_iter - 1) break;
    v_zero(Ar.data(), n);
    apply(A, Ar.data(), r.data());
    inf.ops++;
    ps.emplace_back(r);    // p_{k+1} starts as r
    Aps.emplace_back(Ar);  // Ap_{k+1} starts as Ar
    T* pn = ps[k + 1].data();
    T* Apn = Aps[k + 1].data();
    for (int ii = 0; ii <= k; ii++) {
      const T beta = -v_dot(Aps[ii].data(), Ar.data(), n) / v_norm2(Aps[ii].data(), n);
      const T* pi = ps[ii].data();
      const T* Api = Aps[ii].data();
      for (int i = 0; i < n; i++) {
        pn[i] += beta * pi[i];
        Apn[i] += beta * Api[i];
      }
    }
  }
  inf.success = !(k == max_iter - 1);
  k++;
  std::vector<T> chk(n, T(0.0));
  apply(A, chk.data(), x.data());
  inf.ops++;
  double truersq = 0.0;
  for (int i = 0; i < n; i++) truersq += real_of(conj_of(chk[i] - b[i]) * (chk[i] - b[i]));
  v_copy(phi, x.data(), n);
  say_summary(verb, "GCR", inf.success, k, inf.ops, std::sqrt(truersq) / bnorm);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "GCR";
  return inf;
}

// generic_bicgstab.cpp:22-158 (double), :205-341 (complex)
template <typename T>
Info solve_bicgstab(PortOp* A, T* x, const T* b, int n, int max_iter, double eps, const Verb* verb) {
  Info inf;
  std::vector<T> r(n, T(0.0)), r0(n, T(0.0)), p(n, T(0.0)), Ap(n, T(0.0)), s(n, T(0.0)), As(n, T(0.0));
  const double bnorm = std::sqrt(v_norm2(b, n));
  apply(A, Ap.data(), x);
  inf.ops++;
  for (int i = 0; i < n; i++) r[i] = b[i] - Ap[i];
  v_copy(r0.data(), r.data(), n);
  v_copy(p.data(), r.data(), n);
  T rho = v_dot(r0.data(), r.data(), n);
  apply(A, Ap.data(), p.data());
  inf.ops++;
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T alpha = rho / v_dot(r0.data(), Ap.data(), n);
    for (int i = 0; i < n; i++) s[i] = r[i] - alpha * Ap[i];
    apply(A, As.data(), s.data());
    inf.ops++;
    const T omega = v_dot(As.data(), s.data(), n) / v_dot(As.data(), As.data(), n);
    for (int i = 0; i < n; i++) x[i] = x[i] + alpha * p[i] + omega * s[i];
    for (int i = 0; i < n; i++) r[i] = s[i] - omega * As[i];
    rsq = v_norm2(r.data(), n);
    say_resid(verb, "BiCGStab", k + 1, inf.ops, std::sqrt(rsq) / bnorm);
    if (std::sqrt(rsq) < eps * bnorm || k == max_iter - 1) break;
    const T rho_new = v_dot(r0.data(), r.data(), n);
    const T beta = rho_new / rho * (alpha / omega);
    rho = rho_new;
    for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega * Ap[i]);
    v_zero(Ap.data(), n);
    apply(A, Ap.data(), p.data());
    inf.ops++;
  }
  inf.success = !(k == max_iter - 1);
  k++;
  apply(A, Ap.data(), x);
  inf.ops++;
  const double truersq = v_diffnorm2(Ap.data(), b, n);
  say_summary(verb, "BiCGStab", inf.success, k, inf.ops, std::sqrt(truersq) / bnorm);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "BiCGStab";
  return inf;
}

// generic_bicgstab_l.cpp:25-268 (double), :317-560 (complex).  Sleijpen-Fokkema BiCGStab(l):
// l BiCG steps, then a modified Gram-Schmidt "MR part" whose by-products update x, r, u.
template <typename T>
Info solve_bicgstab_l(PortOp* A, T* x, const T* b, int n, int max_iter, double eps, int l, const Verb* verb) {
  Info inf;
  std::ostringstream nm;
  nm << "BiCGStab-" << l;
  std::vector<T> r0(n, T(0.0));
  std::vector<std::vector<T> > r(l + 1, std::vector<T>(n, T(0.0))), u(l + 1, std::vector<T>(n, T(0.0)));
  std::vector<double> sigma(l + 1, 0.0);
  std::vector<T> gam(l + 1, T(0.0)), gam_p(l + 1, T(0.0)), gam_pp(l + 1, T(0.0));
  std::vector<std::vector<T> > tau(l + 1, std::vector<T>(l + 1, T(0.0)));
  T rho0 = 1, rho1, alpha = 0, omega = 1, beta;
  const double bnorm = std::sqrt(v_norm2(b, n));
  apply(A, u[0].data(), x);
  inf.ops++;
  for (int i = 0; i < n; i++) r[0][i] = b[i] - u[0][i];
  sigma[0] = v_norm2(r[0].data(), n);
  v_copy(r0.data(), r[0].data(), n);
  v_zero(u[0].data(), n);
  int k;
  for (k = 0; k < max_iter; k += l) {
    rho0 *= -omega;
    for (int j = 0; j < l; j++) {  // BiCG part
      rho1 = v_dot(r0.data(), r[j].data(), n);
      beta = alpha * rho1 / rho0;
      rho0 = rho1;
      for (int i = 0; i <= j; i++)
        for (int h = 0; h < n; h++) u[i][h] = r[i][h] - beta * u[i][h];
      v_zero(u[j + 1].data(), n);
      apply(A, u[j + 1].data(), u[j].data());
      inf.ops++;
      alpha = rho0 / v_dot(r0.data(), u[j + 1].data(), n);
      for (int i = 0; i <= j; i++)
        for (int h = 0; h < n; h++) r[i][h] = r[i][h] - alpha * u[i + 1][h];
      apply(A, r[j + 1].data(), r[j].data());
      inf.ops++;
      for (int h = 0; h < n; h++) x[h] = x[h] + alpha * u[0][h];
    }
    for (int j = 1; j <= l; j++) {  // MR part (modified Gram-Schmidt)
      for (int i = 1; i < j; i++) {
        tau[i][j] = v_dot(r[i].data(), r[j].data(), n) / sigma[i];
        for (int h = 0; h < n; h++) r[j][h] = r[j][h] - tau[i][j] * r[i][h];
      }
      sigma[j] = v_norm2(r[j].data(), n);
      gam_p[j] = v_dot(r[j].data(), r[0].data(), n) / sigma[j];
    }
    gam[l] = gam_p[l];
    omega = gam[l];
    for (int j = l - 1; j > 0; j--) {  // gamma = T^{-1} gamma'
      gam[j] = gam_p[j];
      for (int i = j + 1; i <= l; i++) gam[j] = gam[j] - tau[j][i] * gam[i];
    }
    for (int j = 1; j < l; j++) {  // gamma'' = T S gamma
      gam_pp[j] = gam[j + 1];
      for (int i = j + 1; i < l; i++) gam_pp[j] = gam_pp[j] + tau[j][i] * gam[i + 1];
    }
    for (int h = 0; h < n; h++) {
      x[h] = x[h] + gam[1] * r[0][h];
      u[0][h] = u[0][h] - gam[l] * u[l][h];
      r[0][h] = r[0][h] - gam_p[l] * r[l][h];
    }
    for (int j = 1; j < l; j++)
      for (int h = 0; h < n; h++) {
        u[0][h] = u[0][h] - gam[j] * u[j][h];
        x[h] = x[h] + gam_pp[j] * r[j][h];
        r[0][h] = r[0][h] - gam_p[j] * r[j][h];
      }
    sigma[0] = v_norm2(r[0].data(), n);
    say_resid(verb, nm.str(), k + l, inf.ops, std::sqrt(sigma[0]) / bnorm);
    if (std::sqrt(sigma[0]) < eps * bnorm) break;
  }
  inf.success = !(k >= max_iter - 1);
  k++;
  apply(A, u[0].data(), x);
  inf.ops++;
  const double truersq = v_diffnorm2(u[0].data(), b, n);
  say_summary(verb, nm.str(), inf.success, k, inf.ops, std::sqrt(truersq) / bnorm);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = nm.str();
  return inf;
}

// generic_gelim.cpp:19-119 (double), :122-225 (complex): Gauss-Jordan with partial pivoting on
// the augmented matrix; returns false on an all-zero pivot column.
template <typename T>
bool gauss_jordan(T* x, const T* rhs, const std::vector<std::vector<T> >& M, int m) {
  std::vector<std::vector<T> > G(m, std::vector<T>(m + 1));
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < m; j++) G[i][j] = M[i][j];
    G[i][m] = rhs[i];
  }
  for (int i = 0; i < m; i++) {
    double best = 0.0;
    int piv = -1;
    for (int j = i; j < m; j++)
      if (std::abs(G[j][i]) > best) {
        piv = j;
        best = std::abs(G[j][i]);
      }
    if (piv == -1) return false;
    if (piv != i)
      for (int j = i; j < m + 1; j++) std::swap(G[i][j], G[piv][j]);
    for (int j = i + 1; j < m + 1; j++) G[i][j] = G[i][j] / G[i][i];
    G[i][i] = 1.0;
    for (int j = 0; j < m; j++) {
      if (j == i) continue;
      for (int k = i + 1; k < m + 1; k++) G[j][k] = G[j][k] - G[j][i] * G[i][k];
      G[j][i] = 0.0;
    }
  }
  for (int i = 0; i < m; i++) x[i] = G[i][m];
  return true;
}

// generic_gmres.cpp:40-383 (double), :429-774 (complex).  Arnoldi (modified Gram-Schmidt), then
// the least-squares problem through the UN-conjugated normal equations H^T H y = beta H^T e1
// (generic_gmres.cpp:590-602), and an explicit residual with a second operator apply per step.
template <typename T>
Info solve_gmres(PortOp* A, T* phi, const T* b, int n, int max_iter, double eps, const Verb* verb,
                 bool is_complex) {
  Info inf;
  if (n < max_iter) max_iter = n;
  max_iter++;
  std::vector<std::vector<T> > q, h;  // h[c] is column c of the Hessenberg matrix (c+2 entries used)
  std::vector<T> res(n, T(0.0)), tmp(n, T(0.0)), tmp2(n, T(0.0)), y(max_iter, T(0.0)), bhTy(max_iter, T(0.0));
  apply(A, tmp.data(), phi);
  inf.ops++;
  for (int i = 0; i < n; i++) res[i] = b[i] - tmp[i];
  q.emplace_back(n);
  const double beta = std::sqrt(v_norm2(res.data(), n));
  const double bnorm = std::sqrt(v_norm2(b, n));
  for (int i = 0; i < n; i++) q[0][i] = res[i] / beta;
  double localres = 0.0;
  int iter;
  for (iter = 1; iter < max_iter; iter++) {
    q.emplace_back(n, T(0.0));
    h.emplace_back(max_iter + 1, T(0.0));
    apply(A, q[iter].data(), q[iter - 1].data());
    inf.ops++;
    for (int j = 0; j < iter; j++) {
      h[iter - 1][j] = v_dot(q[j].data(), q[iter].data(), n);
      for (int i = 0; i < n; i++) q[iter][i] = q[iter][i] - h[iter - 1][j] * q[j][i];
    }
    h[iter - 1][iter] = std::sqrt(v_norm2(q[iter].data(), n));
    for (int i = 0; i < n; i++) q[iter][i] = q[iter][i] / h[iter - 1][iter];
    for (int i = 0; i < iter; i++) bhTy[i] = beta * h[i][0];
    std::vector<std::vector<T> > hTh(iter, std::vector<T>(iter));
    for (int i = 0; i < iter; i++)
      for (int j = 0; j < iter; j++) {
        T s = 0.0;
        for (int k = 0; k < iter + 1; k++) s = s + h[i][k] * h[j][k];
        hTh[i][j] = s;
      }
    if (!gauss_jordan(y.data(), bhTy.data(), hTh, iter)) break;
    for (int i = 0; i < n; i++) tmp[i] = 0.0;
    for (int j = 0; j < iter; j++)
      for (int i = 0; i < n; i++) tmp[i] = tmp[i] + q[j][i] * y[j];
    for (int i = 0; i < n; i++) tmp2[i] = phi[i] + tmp[i];
    apply(A, res.data(), tmp2.data());
    inf.ops++;
    localres = 0.0;
    for (int i = 0; i < n; i++) localres = localres + real_of(conj_of(b[i] - res[i]) * (b[i] - res[i]));
    localres = std::sqrt(localres);
    say_resid(verb, "GMRES", iter, inf.ops, localres / bnorm);
    if (localres < eps * bnorm) break;
  }
  for (int i = 0; i < n; i++) phi[i] = tmp2[i];
  if (iter == max_iter) {
    inf.success = false;
    iter--;
  } else {
    inf.success = true;
  }
  say_summary(verb, "GMRES", inf.success, iter, inf.ops, localres / bnorm);
  inf.resSq = localres * localres;
  inf.iter = iter;
  if (is_complex && !inf.success) inf.iter--;  // generic_gmres.cpp:768 (complex overload only)
  inf.name = "GMRES";
  return inf;
}

// generic_cg_m.cpp:23-309 (double), :312-599 (complex).  Multishift CG: solves (A + shift_n) x_n = b.
template <typename T>
Info solve_cg_m(PortOp* A, T** phi, const T* b, int n_shift, int n, int check_every, int max_iter, double eps,
                double* shifts, bool worst_first, const Verb* verb) {
  Info inf;
  inf.res_multi.assign(n_shift, 0.0);
  std::vector<T> alpha_s(n_shift, T(0.0)), beta_s(n_shift, T(1.0)), zeta_s(n_shift, T(1.0)), zeta_prev(n_shift, T(1.0));
  std::vector<std::vector<T> > store(n_shift, std::vector<T>(n));
  std::vector<T*> p_s(n_shift);
  for (int s = 0; s < n_shift; s++) p_s[s] = store[s].data();
  std::vector<T> r(n, T(0.0)), p(n, T(0.0)), Ap(n, T(0.0));
  std::vector<int> mapping(n_shift);
  for (int s = 0; s < n_shift; s++) mapping[s] = s;
  int live = n_shift;
  T beta = 1.0, alpha = 0.0, beta_prev;
  const double bnorm = std::sqrt(v_norm2(b, n));
  for (int s = 0; s < n_shift; s++) {
    v_copy(p_s[s], b, n);
    v_zero(phi[s], n);
  }
  v_copy(p.data(), b, n);
  v_copy(r.data(), b, n);
  apply(A, Ap.data(), p.data());
  inf.ops++;
  double rsq = v_norm2(r.data(), n), rsq_new = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    beta_prev = beta;
    beta = -rsq / v_dot(p.data(), Ap.data(), n);
    for (int s = 0; s < live; s++) {
      const T z_old = zeta_s[s];
      zeta_s[s] = (zeta_s[s] * zeta_prev[s] * beta_prev) /
                  (beta * alpha * (zeta_prev[s] - zeta_s[s]) + zeta_prev[s] * beta_prev * (1.0 - shifts[s] * beta));
      zeta_prev[s] = z_old;
      beta_s[s] = beta * zeta_s[s] / zeta_prev[s];
      for (int i = 0; i < n; i++) phi[s][i] = phi[s][i] - beta_s[s] * p_s[s][i];
    }
    for (int i = 0; i < n; i++) r[i] = r[i] + beta * Ap[i];
    rsq_new = v_norm2(r.data(), n);
    say_resid(verb, "CG-M", k + 1, inf.ops, std::sqrt(rsq_new) / bnorm);
    if (k % check_every == 0) {
      for (int s = 0; s < live; s++) {
        if (abs_of(zeta_s[s]) * std::sqrt(rsq_new) < eps * bnorm) {
          live--;
          if (live != s) {  // retire shift s by swapping it with the last live one
            std::swap(mapping[live], mapping[s]);
            std::swap(phi[live], phi[s]);
            std::swap(p_s[live], p_s[s]);
            std::swap(alpha_s[live], alpha_s[s]);
            std::swap(beta_s[live], beta_s[s]);
            std::swap(zeta_s[live], zeta_s[s]);
            std::swap(zeta_prev[live], zeta_prev[s]);
            std::swap(shifts[live], shifts[s]);
            s--;
          }
        }
      }
    }
    if ((worst_first && std::abs(zeta_s[0]) * std::sqrt(rsq_new) < eps * bnorm) || live == 0 || k == max_iter - 1)
      break;
    alpha = rsq_new / rsq;
    rsq = rsq_new;
    for (int s = 0; s < live; s++) {
      alpha_s[s] = alpha * zeta_s[s] * beta_s[s] / (zeta_prev[s] * beta);
      for (int i = 0; i < n; i++) p_s[s][i] = zeta_s[s] * r[i] + alpha_s[s] * p_s[s][i];
    }
    for (int i = 0; i < n; i++) p[i] = r[i] + alpha * p[i];
    apply(A, Ap.data(), p.data());
    inf.ops++;
  }
  inf.success = !(k == max_iter - 1);
  k++;
  for (int s = 0; s < n_shift; s++) {  // undo the permutation of phi[] and shifts[]
    if (mapping[s] != s) {
      for (int m = s + 1; m < n_shift; m++) {
        if (mapping[m] == s) {
          std::swap(phi[m], phi[s]);
          std::swap(shifts[m], shifts[s]);
          mapping[m] = mapping[s];
          mapping[s] = s;
          s--;
          break;
        }
      }
    }
  }
  std::vector<double> rel(n_shift);
  for (int s = 0; s < n_shift; s++) {
    v_zero(Ap.data(), n);
    apply(A, Ap.data(), phi[s]);
    inf.ops++;
    for (int i = 0; i < n; i++) Ap[i] = Ap[i] + (shifts[s] * phi[s][i]);
    inf.res_multi[s] = v_diffnorm2(Ap.data(), b, n);
    rel[s] = std::sqrt(inf.res_multi[s]) / bnorm;
  }
  if (verb && verb->level >= 1) {  // verbosity.h:101-117 (label is hard-coded "CG-M ")
    std::cout << verb->prefix << "CG-M " << " Success " << (inf.success ? "Y" : "N") << " Iter " << k << " Ops "
              << inf.ops << " RelRes ";
    for (int s = 0; s < n_shift; s++) std::cout << rel[s] << " ";
    std::cout << "\n";
  }
  inf.resSq = 0.0;  // generic_cg_m.cpp:596 returns truersq which is never assigned
  inf.iter = k;
  inf.name = "CG-M";
  return inf;
}

// The *_restart wrappers (e.g. generic_cg.cpp:235-275): repeat solver(restart_freq) until converged.
// GMRES compares an absolute residual in its final success test (generic_gmres.cpp:807).
template <typename T, typename F>
Info restarted(const std::string& label, T* phi, const T* b, int n, int max_iter, double eps, int rf,
               const Verb* verb, bool gmres_abs_quirk, F inner) {
  const double bnorm = std::sqrt(v_norm2(b, n));
  Verb vin = inner_verb(verb);
  Info inf;
  int iter = 0, ops = 0;
  do {
    inf = inner(rf, &vin);
    iter += inf.iter;
    ops += inf.ops;
    say_restart(verb, label, iter, ops, std::sqrt(inf.resSq) / bnorm);
  } while (iter < max_iter && inf.success == false && std::sqrt(inf.resSq) / bnorm > eps);
  inf.iter = iter;
  inf.ops = ops;
  say_summary(verb, label, inf.success, iter, ops, std::sqrt(inf.resSq) / bnorm);
  inf.name = label;
  if (gmres_abs_quirk)
    inf.success = !(std::sqrt(inf.resSq) > eps);
  else
    inf.success = !(std::sqrt(inf.resSq) / bnorm > eps);
  return inf;
}

template <typename T>
Info dispatch(int solver, PortOp* A, T* phi, const T* b, int max_iter, double eps, int rf, int l, const Verb* verb) {
  const int n = A->size;
  const bool isc = A->is_complex;
  std::ostringstream lab;
  switch (solver) {
    case ORC_CG: return solve_cg<T>(A, phi, b, n, max_iter, eps, verb);
    case ORC_CR: return solve_cr<T>(A, phi, b, n, max_iter, eps, verb, isc);
    case ORC_GCR: return solve_gcr<T>(A, phi, b, n, max_iter, eps, verb);
    case ORC_BICGSTAB: return solve_bicgstab<T>(A, phi, b, n, max_iter, eps, verb);
    case ORC_BICGSTAB_L: return solve_bicgstab_l<T>(A, phi, b, n, max_iter, eps, l, verb);
    case ORC_GMRES: return solve_gmres<T>(A, phi, b, n, max_iter, eps, verb, isc);
    case ORC_CG_RESTART:
      lab << "CG(" << rf << ")";
      return restarted<T>(lab.str(), phi, b, n, max_iter, eps, rf, verb, false,
                          [&](int m, const Verb* v) { return solve_cg<T>(A, phi, b, n, m, eps, v); });
    case ORC_CR_RESTART:
      lab << "CR(" << rf << ")";
      return restarted<T>(lab.str(), phi, b, n, max_iter, eps, rf, verb, false,
                          [&](int m, const Verb* v) { return solve_cr<T>(A, phi, b, n, m, eps, v, isc); });
    case ORC_GCR_RESTART:
      lab << "GCR(" << rf << ")";
      return restarted<T>(lab.str(), phi, b, n, max_iter, eps, rf, verb, false,
                          [&](int m, const Verb* v) { return solve_gcr<T>(A, phi, b, n, m, eps, v); });
    case ORC_BICGSTAB_RESTART:
      lab << "BiCGStab(" << rf << ")";
      return restarted<T>(lab.str(), phi, b, n, max_iter, eps, rf, verb, false,
                          [&](int m, const Verb* v) { return solve_bicgstab<T>(A, phi, b, n, m, eps, v); });
    case ORC_BICGSTAB_L_RESTART:
      lab << "BiCGStab-" << l << "(" << rf << ")";
      return restarted<T>(lab.str(), phi, b, n, max_iter, eps, rf, verb, false,
                          [&](int m, const Verb* v) { return solve_bicgstab_l<T>(A, phi, b, n, m, eps, l, v); });
    case ORC_GMRES_RESTART:
      lab << "GMRES(" << rf << ")";
      return restarted<T>(lab.str(), phi, b, n, max_iter, eps, rf, verb, true,
                          [&](int m, const Verb* v) { return solve_gmres<T>(A, phi, b, n, m, eps, v, isc); });
    default: return Info();
  }
}

void export_info(const Info& inf, orc_result* out) {
  std::memset(out, 0, sizeof(*out));
  out->resSq = inf.resSq;
  out->iter = inf.iter;
  out->success = inf.success ? 1 : 0;
  out->ops_count = inf.ops;
  out->n_rhs = inf.res_multi.empty() ? -1 : (int)inf.res_multi.size();
  for (size_t i = 0; i < inf.res_multi.size() && i < 32; i++) out->resSqmrhs[i] = inf.res_multi[i];
  std::strncpy(out->name, inf.name.c_str(), sizeof(out->name) - 1);
}

}  // namespace

extern "C" {

const char* port_kind(void) { return "port"; }

void* port_rng_new(unsigned seed) { return new std::mt19937(seed); }
void port_rng_free(void* rng) { delete (std::mt19937*)rng; }

// u1_utils/u1_utils.cpp:92-112 gauss_gauge_u1 : theta ~ N(0, 1/sqrt(beta)), link = polar(1, theta)
void port_gauss_gauge_u1(void* rng, double* links, int X, int Y, double beta) {
  std::mt19937& g = *(std::mt19937*)rng;
  cplx* U = (cplx*)links;
  if (beta < 0) beta = -beta;
  if (beta == 0) {  // falls back to a hot start first (u1_utils.cpp:79-89), then still runs the loop below
    std::uniform_real_distribution<> flat(-3.14159265358979323846, 3.14159265358979323846);
    for (int i = 0; i < 2 * X * Y; i++) U[i] = std::polar(1.0, flat(g));
  }
  std::normal_distribution<> dist(0.0, 1.0 / std::sqrt(beta));
  for (int i = 0; i < 2 * X * Y; i++) U[i] = std::polar(1.0, dist(g));
}
void port_unit_gauge_u1(double* links, int X, int Y) {  // u1_utils.cpp:69-76
  for (int i = 0; i < 2 * X * Y; i++) {
    links[2 * i] = 1.0;
    links[2 * i + 1] = 0.0;
  }
}
void port_gaussian_real(void* rng, double* v, int n) {  // generic_vector.h:35-45
  std::normal_distribution<> dist(0.0, 1.0);
  std::mt19937& g = *(std::mt19937*)rng;
  for (int i = 0; i < n; i++) v[i] = dist(g);
}
void port_gaussian_complex(void* rng, double* v, int n) {  // generic_vector.h:48-60
  std::normal_distribution<> dist(0.0, 1.0);
  std::mt19937& g = *(std::mt19937*)rng;
  cplx* c = (cplx*)v;
  for (int i = 0; i < n; i++) {
    // g++ evaluates the two constructor arguments right-to-left, so the reference draws the
    // IMAGINARY part first (pinned by tests/test_oracle_cpu.py against oracle/_ref).
    const double im = dist(g);
    const double re = dist(g);
    c[i] = cplx(re, im);
  }
}
// u1_utils.cpp:17-33 read_gauge_u1: file order is x outer, y, mu inner; stored transposed.
int port_read_gauge_u1(double* links, int X, int Y, const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) return 1;
  cplx* U = (cplx*)links;
  for (int x = 0; x < X; x++)
    for (int y = 0; y < Y; y++)
      for (int mu = 0; mu < 2; mu++) {
        double th = 0.0;
        if (fscanf(f, "%lf", &th) != 1) {
          fclose(f);
          return 2;
        }
        U[y * 2 * X + x * 2 + mu] = std::polar(1.0, th);
      }
  fclose(f);
  return 0;
}
// u1_utils.cpp:190-207 get_plaquette_u1
void port_plaquette_u1(const double* links, int X, int Y, double out[2]) {
  const cplx* U = (const cplx*)links;
  cplx plaq = 0.0;
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < X; x++) {
      cplx t = U[2 * X * y + 2 * x] * U[2 * X * y + 2 * ((x + 1) % X) + 1] * std::conj(U[2 * X * ((y + 1) % Y) + 2 * x]) *
               std::conj(U[2 * X * y + 2 * x + 1]);
      plaq += t;
    }
  plaq = plaq / ((double)(X * Y));
  out[0] = plaq.real();
  out[1] = plaq.imag();
}

void port_dot(int is_complex, const double* a, const double* b, int n, double out[2]) {
  if (is_complex) {
    cplx r = v_dot((const cplx*)a, (const cplx*)b, n);
    out[0] = r.real();
    out[1] = r.imag();
  } else {
    out[0] = v_dot(a, b, n);
    out[1] = 0.0;
  }
}
double port_norm2sq(int is_complex, const double* a, int n) {
  return is_complex ? v_norm2((const cplx*)a, n) : v_norm2(a, n);
}
double port_diffnorm2sq(int is_complex, const double* a, const double* b, int n) {
  return is_complex ? v_diffnorm2((const cplx*)a, (const cplx*)b, n) : v_diffnorm2(a, b, n);
}

void* port_op_prepare(const orc_op_desc* d) {
  if (d->view != 0 || d->kind > ORC_OP_STAG_M2MDEODOE_U1) return 0;  // composite stencil operators, symmetric shifts, 2-link
                                                                      // Laplace and the index operator: reference library only
  PortOp* op = new PortOp();
  op->d = *d;
  op->nc = d->Nc > 0 ? d->Nc : 1;
  const int V = d->X * d->Y;
  op->is_complex =
      !(d->kind == ORC_OP_LAPLACE_REAL || d->kind == ORC_OP_LAPLACE_REAL_NC || d->kind == ORC_OP_STAG_FREE_REAL);
  op->size = V;
  op->has_two = false;
  if (d->kind == ORC_OP_LAPLACE_NC || d->kind == ORC_OP_LAPLACE_REAL_NC) op->size = V * op->nc;
  if (d->kind == ORC_OP_STENCIL_FROM_STAG) {
    build_stag_stencil(*op);
    op->size = V;
  } else if (d->kind == ORC_OP_STENCIL) {
    const size_t m = (size_t)V * op->nc * op->nc;
    const cplx* c = (const cplx*)d->clover;
    const cplx* h = (const cplx*)d->hopping;
    op->clover.assign(c, c + m);
    op->hopping.assign(h, h + 4 * m);
    op->has_two = d->has_two != 0;
    if (op->has_two) {
      const cplx* t = (const cplx*)d->two_link;
      op->two_link.assign(t, t + 8 * m);
    }
    op->shift = cplx(d->shift[0], d->shift[1]);
    op->eo_shift = cplx(d->eo_shift[0], d->eo_shift[1]);
    op->dof_shift = cplx(d->dof_shift[0], d->dof_shift[1]);
    op->size = V * op->nc;
  }
  return op;
}
void port_op_free(void* op) { delete (PortOp*)op; }
int port_op_is_complex(void* op) { return ((PortOp*)op)->is_complex ? 1 : 0; }
int port_op_size(void* op) { return ((PortOp*)op)->size; }
void port_op_apply(void* opv, double* lhs, const double* rhs) {
  PortOp* op = (PortOp*)opv;
  if (op->is_complex)
    apply_c(op, (cplx*)lhs, (const cplx*)rhs);
  else
    apply_r(op, lhs, rhs);
}

// coarse_stencil.cpp:395-1512 (DIR_ALL): eo / oe = hopping on one parity; tb / bt = clover + hopping (+ two-link)
// between the halves of the colour index.  Same accumulation order as op_stencil, no shifts.
void port_stencil_apply_part(void* opv, int part, double* lhs_, const double* rhs_) {
  PortOp& op = *(PortOp*)opv;
  cplx* out = (cplx*)lhs_;
  const cplx* in = (const cplx*)rhs_;
  const int X = op.d.X, Y = op.d.Y, nc = op.nc;
  const int L = X * Y * nc;
  static const int hop_dx[4] = {1, 0, -1, 0}, hop_dy[4] = {0, 1, 0, -1};
  static const int two_dx[8] = {2, 1, 0, -1, -2, -1, 0, 1}, two_dy[8] = {0, 1, 2, 1, 0, -1, -2, -1};
  for (int i = 0; i < L; i++) {
    out[i] = 0.0;
    const int row = i % nc, site = i / nc, x = site % X, y = site / X;
    int c0 = 0, c1 = nc;
    bool live;
    const bool colour_split = (part == 3 || part == 4);
    if (!colour_split) {
      const bool even = ((x + y) % 2 == 0);
      live = (part == 1) ? even : !even;
    } else {
      const bool top = row < nc / 2;
      live = (part == 3) ? top : !top;
      c0 = (part == 3) ? nc / 2 : 0;
      c1 = (part == 3) ? nc : nc / 2;
    }
    if (!live) continue;
    if (colour_split)
      for (int c = c0; c < c1; c++) out[i] += op.clover[c + nc * i] * in[site * nc + c];
    for (int d = 0; d < 4; d++) {
      const int xn = (x + hop_dx[d] + X) % X, yn = (y + hop_dy[d] + Y) % Y;
      for (int c = c0; c < c1; c++) out[i] += op.hopping[c + nc * i + d * nc * L] * in[(yn * X + xn) * nc + c];
    }
    if (colour_split && op.has_two) {
      for (int d = 0; d < 8; d++) {
        const int xn = (x + two_dx[d] + 2 * X) % X, yn = (y + two_dy[d] + 2 * Y) % Y;
        for (int c = c0; c < c1; c++) out[i] += op.two_link[c + nc * i + d * nc * L] * in[(yn * X + xn) * nc + c];
      }
    }
  }
}

void port_eoprec_prepare(void* opv, double* rhs_e, const double* rhs_orig) {
  PortOp* op = (PortOp*)opv;
  op_eoprec_prepare((cplx*)rhs_e, (const cplx*)rhs_orig, (const cplx*)op->d.links, op->d.X, op->d.Y, op->d.mass);
}
void port_eoprec_reconstruct(void* opv, double* lhs_full, const double* lhs_e, const double* rhs_o) {
  PortOp* op = (PortOp*)opv;
  op_eoprec_reconstruct((cplx*)lhs_full, (const cplx*)lhs_e, (const cplx*)rhs_o, (const cplx*)op->d.links, op->d.X,
                        op->d.Y, op->d.mass);
}

int port_solve(int solver, void* opv, double* phi, const double* phi0, int max_iter, double eps, int restart_freq,
               int l, int verbosity, orc_result* out) {
  PortOp* op = (PortOp*)opv;
  Verb verb{verbosity, "[port] "};
  Info inf;
  if (op->is_complex)
    inf = dispatch<cplx>(solver, op, (cplx*)phi, (const cplx*)phi0, max_iter, eps, restart_freq, l, &verb);
  else
    inf = dispatch<double>(solver, op, phi, phi0, max_iter, eps, restart_freq, l, &verb);
  export_info(inf, out);
  return 0;
}

int port_solve_cg_m(void* opv, double** phi, const double* phi0, int n_shift, int resid_freq_check, int max_iter,
                    double eps, double* shifts, int worst_first, int verbosity, orc_result* out) {
  PortOp* op = (PortOp*)opv;
  Verb verb{verbosity, "[port] "};
  Info inf;
  if (op->is_complex)
    inf = solve_cg_m<cplx>(op, (cplx**)phi, (const cplx*)phi0, n_shift, op->size, resid_freq_check, max_iter, eps,
                           shifts, worst_first != 0, &verb);
  else
    inf = solve_cg_m<double>(op, phi, phi0, n_shift, op->size, resid_freq_check, max_iter, eps, shifts,
                             worst_first != 0, &verb);
  export_info(inf, out);
  return 0;
}

}  // extern "C"
