// dev_solvers.cpp -- the Krylov shells on DEVICE-resident vectors.
//
// These are the reference's solver loops (generic_cg.cpp, generic_cr.cpp, generic_gcr.cpp,
// generic_bicgstab.cpp, generic_bicgstab_l.cpp, generic_gmres.cpp, generic_cg_m.cpp) with the
// vectors living in HBM: every vector operation is one fused C-ABI call (include/glb200.h), the
// scalar algebra (alpha, beta, omega, the tau/gamma recurrences, the small GMRES least-squares
// system) stays on the host in std::complex<double> exactly as the reference evaluates it.
// Stopping tests, iteration/ops counting, success flags, names and printed lines follow the
// reference line by line, including its quirks (see the comment at each).
#include <cmath>
#include <sstream>
#include <utility>

#include "dev_internal.hpp"
#include "generic_eigenvalues.h"

using namespace glbx;

namespace {

bool g_force_host_scalars = false;

inline double real_part(double a) { return a; }
inline double real_part(const zcplx& a) { return a.real(); }
inline double conj_of(double a) { return a; }
inline zcplx conj_of(const zcplx& a) { return std::conj(a); }
inline double zeta_mag(double z) { return z; }  // generic_cg_m.cpp:166 : the real overload omits abs()
inline double zeta_mag(const zcplx& z) { return std::abs(z); }
template <typename T>
struct IsComplex {
  enum { value = 0 };
};
template <>
struct IsComplex<zcplx> {
  enum { value = 1 };
};

template <typename T>
DevOp<T> make_op(void (*fn)(T*, T*, void*), void* extra, int size) {
  DevOp<T> A;
  A.fn = fn;
  A.extra = extra;
  A.n = (size_t)size;
  A.ops = 0;
  A.native = 0;
  void (*std_cb)(T*, T*, void*) = &glb200_apply_dev;
  if (fn == std_cb) {
    A.native = (glb_operator*)extra;
    A.ctx = glb_op_context(A.native);
    if ((size_t)size != glb_op_local_size(A.native)) throw Error("vector size does not match the operator");
    if (glb_op_dtype(A.native) != (int)Traits<T>::dtype) throw Error("scalar type does not match the operator");
  } else {
    A.ctx = glb200_default_context();
  }
  return A;
}

// common tail: report a failed device call the only way the API allows (success=false)
inversion_info failed(const char* alg, const std::exception& e) {
  std::cerr << "[glb200] " << alg << " aborted: " << e.what() << std::endl;
  inversion_info inf;
  inf.success = false;
  inf.name = alg;
  return inf;
}

// ------------------------------------------------------------------------------------------ CG
// generic_cg.cpp:132-231 / :278-377
template <typename T>
inversion_info cg_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                      inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  double bsqrt, truersq;
  int k;
  T* Ap = W.get();

  if (A.native && !g_force_host_scalars && glb_cg_solve_supported(A.native) && max_iter >= 1) {
    // device-resident loop: alpha, beta and the stopping test never leave the GPU
    std::vector<double> hist;
    const bool detail = (verb != 0 && verb->verbosity == VERB_DETAIL);
    if (detail) hist.resize(max_iter);
    glb_cg_report rep;
    GLBX(glb_cg_solve(A.native, x, b, max_iter, eps, &rep, detail ? hist.data() : 0, detail ? max_iter : 0));
    bsqrt = rep.bnorm;
    if (detail)
      for (int i = 0; i < rep.iterations; i++) print_verbosity_resid(verb, "CG", i + 1, 2 + i, sqrt(hist[i]) / bsqrt);
    A.ops = rep.ops;
    inf.success = !rep.hit_max_iter;
    k = rep.iterations;
  } else {
    T* r = W.get();
    T* p = W.get();
    bsqrt = sqrt(B.norm2sq(b));
    A.apply(p, x);
    B.sub(b, p, r);
    B.copy(p, r);
    T pAp = A.apply_dot(Ap, p, p);
    double rsq = B.norm2sq(r), rsqNew = 0.0;
    for (k = 0; k < max_iter; k++) {
      const T alpha = rsq / pAp;
      rsqNew = B.update_xr_norm(alpha, p, x, -alpha, Ap, r);
      print_verbosity_resid(verb, "CG", k + 1, A.ops, sqrt(rsqNew) / bsqrt);
      if (sqrt(rsqNew) < eps * bsqrt || k == max_iter - 1) break;
      const T beta = rsqNew / rsq;
      rsq = rsqNew;
      B.xpay(r, beta, p);
      pAp = A.apply_dot(Ap, p, p);
    }
    inf.success = !(k == max_iter - 1);
    k++;
  }
  A.apply(Ap, x);
  truersq = B.diffnorm2sq(Ap, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "CG", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "CG";
  return inf;
}

// ------------------------------------------------------------------------------------------ CR
// generic_cr.cpp:28-150 / :198-322
template <typename T>
inversion_info cr_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                      inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  if (A.native && !g_force_host_scalars && max_iter >= 1 && glb_krylov_solve_supported(A.native, GLB_KRYLOV_CR)) {
    // device-resident loop (csrc/krylov.cu): alpha, beta and the stopping test never leave the GPU
    std::vector<double> hist;
    const bool detail = (verb != 0 && verb->verbosity == VERB_DETAIL);
    if (detail) hist.resize(max_iter);
    glb_cg_report rep;
    GLBX(glb_krylov_solve(A.native, GLB_KRYLOV_CR, x, b, max_iter, eps, &rep, detail ? hist.data() : 0,
                          detail ? max_iter : 0));
    if (detail)
      for (int i = 0; i < rep.iterations; i++) print_verbosity_resid(verb, "CR", i + 1, 2 + i, sqrt(hist[i]) / rep.bnorm);
    A.ops = rep.ops;
    // the complex overload tests k == max_iter after the loop and so never reports failure (generic_cr.cpp:288 vs :117)
    inf.success = IsComplex<T>::value ? true : !rep.hit_max_iter;
    T* t = W.get();
    A.apply(t, x);
    const double truersq = B.diffnorm2sq(t, b);
    inf.ops_count = A.ops;
    print_verbosity_summary(verb, "CR", inf.success, rep.iterations, inf.ops_count, sqrt(truersq) / rep.bnorm);
    inf.resSq = truersq;
    inf.iter = rep.iterations;
    inf.name = "CR";
    return inf;
  }
  T *r = W.get(), *Ar = W.get(), *p = W.get(), *Ap = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(p, x);
  B.sub(b, p, r);
  B.copy(p, r);
  A.apply(Ap, p);
  B.copy(Ar, Ap);
  double Apsq = B.norm2sq(Ap), rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T alpha = B.dot(Ap, r) / Apsq;
    rsq = B.update_xr_norm(alpha, p, x, -alpha, Ap, r);
    print_verbosity_resid(verb, "CR", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt || k == max_iter - 1) break;
    const T ApAr = A.apply_dot(Ar, r, Ap);  // Ar = A r and <Ap,Ar> in one pass
    const T beta = -ApAr / Apsq;
    double c[2];
    Traits<T>::pack(beta, c);
    GLBX(glb_update_p_ap_norm(A.ctx, Traits<T>::dtype, size, r, Ar, c, p, Ap, &Apsq));
  }
  // the complex overload tests k == max_iter and so never reports failure (generic_cr.cpp:288 vs :117)
  inf.success = !(k == (IsComplex<T>::value ? max_iter : max_iter - 1));
  k++;
  A.apply(p, x);
  const double truersq = B.diffnorm2sq(p, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "CR", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "CR";
  return inf;
}

// ------------------------------------------------------------------------------------------ GCR
// generic_gcr.cpp:27-156 / :202-333.  All search directions are kept; the re-orthogonalisation
// coefficients beta_ij = -<Ap_i,Ar>/|Ap_i|^2 all use the same Ar (generic_gcr.cpp:286), so the
// inner products are taken in one batched pass and p, Ap are rebuilt with one ordered
// accumulation each -- the arithmetic per element is the reference's.
template <typename T>
inversion_info gcr_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                       inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  T *r = W.get(), *Ar = W.get();
  std::vector<const void*> ps, Aps;
  std::vector<double> Apnorm;
  T* p = W.get();
  T* Ap = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(p, x);
  B.sub(b, p, r);
  B.copy(p, r);
  A.apply(Ap, p);
  double rsq = 0.0;
  int k;
  std::vector<double> dots, coef;
  for (k = 0; k < max_iter; k++) {
    ps.push_back(p);
    Aps.push_back(Ap);
    double dn[3];
    GLBX(glb_dot_norm(A.ctx, Traits<T>::dtype, size, Ap, r, dn));  // <Ap,r>, |Ap|^2
    Apnorm.push_back(dn[2]);
    const T alpha = Traits<T>::unpack(dn) / dn[2];
    rsq = B.update_xr_norm(alpha, p, x, -alpha, Ap, r);
    print_verbosity_resid(verb, "GCR", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt || k == max_iter - 1) break;
    A.apply(Ar, r);
    dots.resize(2 * (k + 1));
    coef.resize(2 * (k + 1));
    GLBX(glb_multi_dot(A.ctx, Traits<T>::dtype, size, k + 1, Aps.data(), Ar, dots.data()));
    for (int ii = 0; ii <= k; ii++) {
      const T beta = -Traits<T>::unpack(&dots[2 * ii]) / Apnorm[ii];
      Traits<T>::pack(beta, &coef[2 * ii]);
    }
    p = W.get();
    Ap = W.get();
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, k + 1, coef.data(), ps.data(), r, p));
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, k + 1, coef.data(), Aps.data(), Ar, Ap));
  }
  inf.success = !(k == max_iter - 1);
  k++;
  A.apply(Ar, x);
  const double truersq = B.diffnorm2sq(Ar, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "GCR", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "GCR";
  return inf;
}

// ------------------------------------------------------------------------------------------ VPGCR
// generic_gcr_var_precond.cpp:28-195 (real) / :220-365 (complex): GCR whose search direction is the
// variably preconditioned residual z = M^-1 r.  As in gcr_dev the coefficients
// beta_ij = -<Ap_i,Az>/|Ap_i|^2 (:310) all use the same Az, so they come from one batched pass.
// The preconditioner follows the device variant of the reference's precond contract
// (generic_gcr_var_precond.h:16-23): void(T* d_lhs, T* d_rhs, int size, void* extra, verbosity*), lhs
// zeroed by the caller before every call but the first (:245 / :296).
template <typename T>
inversion_info gcr_var_precond_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                                   void (*precond)(T*, T*, int, void*, inversion_verbose_struct*), void* precond_info,
                                   inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  inversion_verbose_struct verb_prec;
  shuffle_verbosity_precond(&verb_prec, verb);
  T *r = W.get(), *z = W.get(), *Az = W.get();
  std::vector<const void*> ps, Aps;
  std::vector<double> Apnorm;
  T* p = W.get();
  T* Ap = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(p, x);
  B.sub(b, p, r);
  B.zero(z);
  precond(z, r, size, precond_info, &verb_prec);
  B.copy(p, z);
  A.apply(Ap, p);
  double rsq = 0.0;
  int k;
  std::vector<double> dots, coef;
  for (k = 0; k < max_iter; k++) {
    ps.push_back(p);
    Aps.push_back(Ap);
    double dn[3];
    GLBX(glb_dot_norm(A.ctx, Traits<T>::dtype, size, Ap, r, dn));  // <Ap,r>, |Ap|^2
    Apnorm.push_back(dn[2]);
    const T alpha = Traits<T>::unpack(dn) / dn[2];
    rsq = B.update_xr_norm(alpha, p, x, -alpha, Ap, r);
    print_verbosity_resid(verb, "VPGCR", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt || k == max_iter - 1) break;
    B.zero(z);
    precond(z, r, size, precond_info, &verb_prec);
    A.apply(Az, z);
    dots.resize(2 * (k + 1));
    coef.resize(2 * (k + 1));
    GLBX(glb_multi_dot(A.ctx, Traits<T>::dtype, size, k + 1, Aps.data(), Az, dots.data()));
    for (int ii = 0; ii <= k; ii++) {
      const T beta = -Traits<T>::unpack(&dots[2 * ii]) / Apnorm[ii];
      Traits<T>::pack(beta, &coef[2 * ii]);
    }
    p = W.get();
    Ap = W.get();
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, k + 1, coef.data(), ps.data(), z, p));
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, k + 1, coef.data(), Aps.data(), Az, Ap));
  }
  inf.success = !(k == max_iter - 1);
  k++;
  A.apply(Az, x);
  const double truersq = B.diffnorm2sq(Az, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "VPGCR", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "Variably Preconditioned GCR";
  return inf;
}

// ------------------------------------------------------------------------------------------ BiCGStab
// generic_bicgstab.cpp:22-158 / :205-341
template <typename T>
inversion_info bicgstab_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                            inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  if (A.native && !g_force_host_scalars && max_iter >= 1 && glb_krylov_solve_supported(A.native, GLB_KRYLOV_BICGSTAB)) {
    // device-resident loop (csrc/krylov.cu): alpha, omega, beta and the stopping test never leave the GPU
    std::vector<double> hist;
    const bool detail = (verb != 0 && verb->verbosity == VERB_DETAIL);
    if (detail) hist.resize(max_iter);
    glb_cg_report rep;
    GLBX(glb_krylov_solve(A.native, GLB_KRYLOV_BICGSTAB, x, b, max_iter, eps, &rep, detail ? hist.data() : 0,
                          detail ? max_iter : 0));
    if (detail)  // ops at the time of the print: 2 set-up applies, As of every iteration, Ap of the earlier ones
      for (int i = 0; i < rep.iterations; i++)
        print_verbosity_resid(verb, "BiCGStab", i + 1, 2 * i + 3, sqrt(hist[i]) / rep.bnorm);
    A.ops = rep.ops;
    inf.success = !rep.hit_max_iter;
    T* t = W.get();
    A.apply(t, x);
    const double truersq = B.diffnorm2sq(t, b);
    inf.ops_count = A.ops;
    print_verbosity_summary(verb, "BiCGStab", inf.success, rep.iterations, inf.ops_count, sqrt(truersq) / rep.bnorm);
    inf.resSq = truersq;
    inf.iter = rep.iterations;
    inf.name = "BiCGStab";
    return inf;
  }
  T *r = W.get(), *r0 = W.get(), *p = W.get(), *Ap = W.get(), *s = W.get(), *As = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(Ap, x);
  B.sub(b, Ap, r);
  B.copy(r0, r);
  B.copy(p, r);
  T rho = B.dot(r0, r);
  T r0Ap = A.apply_dot(Ap, p, r0);
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T alpha = rho / r0Ap;
    B.axpyz(-alpha, Ap, r, s);  // s = r - alpha Ap
    double AsAs = 0.0;
    const T sAs = A.apply_dot_norm(As, s, s, &AsAs);  // As = A s, <s,As>, |As|^2 in one pass
    // generic_bicgstab.cpp:271 : omega = dot(As,s)/dot(As,As), both complex
    const T omega = conj_of(sAs) / T(AsAs);
    double ca[2], co[2], out[3];
    Traits<T>::pack(alpha, ca);
    Traits<T>::pack(omega, co);
    GLBX(glb_bicgstab_update(A.ctx, Traits<T>::dtype, size, ca, p, co, s, As, r0, x, r, out));
    rsq = out[0];
    print_verbosity_resid(verb, "BiCGStab", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt || k == max_iter - 1) break;
    const T rhoNew = Traits<T>::unpack(out + 1);
    const T beta = rhoNew / rho * (alpha / omega);
    rho = rhoNew;
    double cb[2];
    Traits<T>::pack(beta, cb);
    GLBX(glb_bicgstab_pupdate(A.ctx, Traits<T>::dtype, size, r, cb, co, Ap, p));
    r0Ap = A.apply_dot(Ap, p, r0);
  }
  inf.success = !(k == max_iter - 1);
  k++;
  A.apply(Ap, x);
  const double truersq = B.diffnorm2sq(Ap, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "BiCGStab", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "BiCGStab";
  return inf;
}

// ------------------------------------------------------------------------------------------ BiCGStab-l
// generic_bicgstab_l.cpp:25-268 / :317-560
template <typename T>
inversion_info bicgstab_l_dev(T* x, T* b, int size, int max_iter, double eps, int l, void (*fn)(T*, T*, void*),
                              void* extra, inversion_verbose_struct* verb) {
  inversion_info inf;
  std::ostringstream nm;
  nm << "BiCGStab-" << l;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  T* r0 = W.get();
  std::vector<T*> r(l + 1), u(l + 1);
  for (int i = 0; i <= l; i++) {
    r[i] = W.get();
    u[i] = W.get();
  }
  std::vector<double> sigma(l + 1, 0.0);
  std::vector<T> gam(l + 1, T(0.0)), gam_p(l + 1, T(0.0)), gam_pp(l + 1, T(0.0));
  std::vector<std::vector<T> > tau(l + 1, std::vector<T>(l + 1, T(0.0)));
  T rho0 = 1, rho1, alpha = 0, omega = 1, beta;
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(u[0], x);
  B.sub(b, u[0], r[0]);
  sigma[0] = B.norm2sq(r[0]);
  B.copy(r0, r[0]);
  B.zero(u[0]);
  for (int i = 1; i <= l; i++) {
    B.zero(u[i]);
    B.zero(r[i]);
  }
  int k;
  for (k = 0; k < max_iter; k += l) {
    rho0 *= -omega;
    for (int j = 0; j < l; j++) {  // BiCG part
      rho1 = B.dot(r0, r[j]);
      beta = alpha * rho1 / rho0;
      rho0 = rho1;
      for (int i = 0; i <= j; i++) B.xpay(r[i], -beta, u[i]);  // u_i = r_i - beta u_i
      const T r0u = A.apply_dot(u[j + 1], u[j], r0);           // u_{j+1} = A u_j, <r0,u_{j+1}>
      alpha = rho0 / r0u;
      for (int i = 0; i <= j; i++) B.axpy(-alpha, u[i + 1], r[i]);  // r_i = r_i - alpha u_{i+1}
      A.apply(r[j + 1], r[j]);
      B.axpy(alpha, u[0], x);
    }
    for (int j = 1; j <= l; j++) {  // MR part: modified Gram-Schmidt
      for (int i = 1; i < j; i++) {
        tau[i][j] = B.dot(r[i], r[j]) / sigma[i];
        B.axpy(-tau[i][j], r[i], r[j]);
      }
      double dn[3];
      GLBX(glb_dot_norm(A.ctx, Traits<T>::dtype, size, r[j], r[0], dn));  // <r_j,r_0>, |r_j|^2
      sigma[j] = dn[2];
      gam_p[j] = Traits<T>::unpack(dn) / sigma[j];
    }
    gam[l] = gam_p[l];
    omega = gam[l];
    for (int j = l - 1; j > 0; j--) {
      gam[j] = gam_p[j];
      for (int i = j + 1; i <= l; i++) gam[j] = gam[j] - tau[j][i] * gam[i];
    }
    for (int j = 1; j < l; j++) {
      gam_pp[j] = gam[j + 1];
      for (int i = j + 1; i < l; i++) gam_pp[j] = gam_pp[j] + tau[j][i] * gam[i + 1];
    }
    B.axpy(gam[1], r[0], x);
    B.axpy(-gam[l], u[l], u[0]);
    B.axpy(-gam_p[l], r[l], r[0]);
    for (int j = 1; j < l; j++) {
      B.axpy(-gam[j], u[j], u[0]);
      B.axpy(gam_pp[j], r[j], x);
      B.axpy(-gam_p[j], r[j], r[0]);
    }
    sigma[0] = B.norm2sq(r[0]);
    print_verbosity_resid(verb, nm.str(), k + l, A.ops, sqrt(sigma[0]) / bsqrt);
    if (sqrt(sigma[0]) < eps * bsqrt) break;
  }
  inf.success = !(k >= max_iter - 1);
  k++;
  A.apply(u[0], x);
  const double truersq = B.diffnorm2sq(u[0], b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, nm.str(), inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = nm.str();
  return inf;
}

// ------------------------------------------------------------------------------------------ GMRES
// generic_gmres.cpp:40-383 / :429-774.  Arnoldi with modified Gram-Schmidt on the device; the
// (iter x iter) UN-conjugated normal equations H^T H y = beta H^T e1 (generic_gmres.cpp:590-602)
// and their Gauss-Jordan solution (generic_gelim.cpp) stay on the host; explicit residual with a
// second operator application per step (generic_gmres.cpp:653).
template <typename T>
inversion_info gmres_dev(T* phi, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                         inversion_verbose_struct* verb) {
  inversion_info inf;
  if (size < max_iter) max_iter = size;
  max_iter++;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  T *res = W.get(), *tmp = W.get(), *tmp2 = W.get();
  std::vector<const void*> q;
  std::vector<std::vector<T> > h;
  std::vector<T> y(max_iter, T(0.0)), bhTy(max_iter, T(0.0));
  A.apply(tmp, phi);
  B.sub(b, tmp, res);
  const double beta = sqrt(B.norm2sq(res));
  const double bres = sqrt(B.norm2sq(b));
  T* q0 = W.get();
  B.rdiv(res, beta, q0);
  q.push_back(q0);
  B.copy(tmp2, phi);  // if the very first elimination fails phi is returned unchanged
  double localres = 0.0;
  int iter;
  std::vector<double> coef;
  for (iter = 1; iter < max_iter; iter++) {
    T* qn = W.get();
    h.push_back(std::vector<T>(max_iter + 1, T(0.0)));
    A.apply(qn, (T*)q[iter - 1]);
    for (int j = 0; j < iter; j++) {
      h[iter - 1][j] = B.dot((const T*)q[j], qn);
      B.axpy(-h[iter - 1][j], (const T*)q[j], qn);
    }
    const double hn = sqrt(B.norm2sq(qn));
    h[iter - 1][iter] = hn;
    B.rdiv(qn, hn, qn);
    q.push_back(qn);
    for (int i = 0; i < iter; i++) bhTy[i] = beta * h[i][0];
    std::vector<std::vector<T> > hTh(iter, std::vector<T>(iter));
    std::vector<T*> rows(iter);
    for (int i = 0; i < iter; i++) {
      rows[i] = hTh[i].data();
      for (int j = 0; j < iter; j++) {
        T s = 0.0;
        for (int kk = 0; kk < iter + 1; kk++) s = s + h[i][kk] * h[j][kk];
        hTh[i][j] = s;
      }
    }
    if (!gaussian_elimination(y.data(), bhTy.data(), rows.data(), iter)) break;
    coef.resize(2 * iter);
    for (int j = 0; j < iter; j++) Traits<T>::pack(y[j], &coef[2 * j]);
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, iter, coef.data(), q.data(), 0, tmp));  // tmp = sum_j q_j y_j
    B.add(phi, tmp, tmp2);
    A.apply(res, tmp2);
    localres = sqrt(B.diffnorm2sq(b, res));
    print_verbosity_resid(verb, "GMRES", iter, A.ops, localres / bres);
    if (localres < eps * bres) break;
  }
  B.copy(phi, tmp2);
  if (iter == max_iter) {
    inf.success = false;
    iter--;
  } else {
    inf.success = true;
  }
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "GMRES", inf.success, iter, inf.ops_count, localres / bres);
  inf.resSq = localres * localres;
  inf.iter = iter;
  if (IsComplex<T>::value && !inf.success) inf.iter--;  // generic_gmres.cpp:768: complex overload only
  inf.name = "GMRES";
  return inf;
}

// ------------------------------------------------------------------------------------------ CG-M
// generic_cg_m.cpp:23-309 / :312-599.  phi is a HOST array of n_shift DEVICE vectors.
template <typename T>
inversion_info cg_m_dev(T** phi, T* b, int n_shift, int size, int check_every, int max_iter, double eps,
                        double* shifts, void (*fn)(T*, T*, void*), void* extra, bool worst_first,
                        inversion_verbose_struct* verb) {
  inversion_info inf(n_shift);
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  std::vector<T> alpha_s(n_shift, T(0.0)), beta_s(n_shift, T(1.0)), zeta_s(n_shift, T(1.0)), zeta_prev(n_shift, T(1.0));
  std::vector<T*> p_s(n_shift);
  std::vector<int> mapping(n_shift);
  for (int s = 0; s < n_shift; s++) {
    p_s[s] = W.get();
    mapping[s] = s;
  }
  T *r = W.get(), *p = W.get(), *Ap = W.get();
  int live = n_shift;
  T beta = 1.0, alpha = 0.0, beta_prev;
  const double bsqrt = sqrt(B.norm2sq(b));
  for (int s = 0; s < n_shift; s++) {
    B.copy(p_s[s], b);
    B.zero(phi[s]);
  }
  B.copy(p, b);
  B.copy(r, b);
  T pAp = A.apply_dot(Ap, p, p);
  double rsq = B.norm2sq(r), rsqNew = 0.0;
  std::vector<double> c0, c1;
  int k;
  for (k = 0; k < max_iter; k++) {
    beta_prev = beta;
    beta = -rsq / pAp;
    c0.resize(2 * live);
    for (int s = 0; s < live; s++) {
      const T z_old = zeta_s[s];
      zeta_s[s] = (zeta_s[s] * zeta_prev[s] * beta_prev) /
                  (beta * alpha * (zeta_prev[s] - zeta_s[s]) + zeta_prev[s] * beta_prev * (1.0 - shifts[s] * beta));
      zeta_prev[s] = z_old;
      beta_s[s] = beta * zeta_s[s] / zeta_prev[s];
      Traits<T>::pack(beta_s[s], &c0[2 * s]);
    }
    // x_s = x_s - beta_s p_s for every live shift, one launch
    GLBX(glb_cgm_update_x(A.ctx, Traits<T>::dtype, size, live, c0.data(), (const void* const*)p_s.data(),
                          (void* const*)phi));
    rsqNew = B.axpy_norm(beta, Ap, r);  // r = r + beta Ap ; |r|^2
    print_verbosity_resid(verb, "CG-M", k + 1, A.ops, sqrt(rsqNew) / bsqrt);
    if (k % check_every == 0) {
      for (int s = 0; s < live; s++) {
        if (zeta_mag(zeta_s[s]) * sqrt(rsqNew) < eps * bsqrt) {
          live--;
          if (live != s) {
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
    if ((worst_first && std::abs(zeta_s[0]) * sqrt(rsqNew) < eps * bsqrt) || live == 0 || k == max_iter - 1) break;
    alpha = rsqNew / rsq;
    rsq = rsqNew;
    c0.resize(2 * live);
    c1.resize(2 * live);
    for (int s = 0; s < live; s++) {
      alpha_s[s] = alpha * zeta_s[s] * beta_s[s] / (zeta_prev[s] * beta);
      Traits<T>::pack(zeta_s[s], &c0[2 * s]);
      Traits<T>::pack(alpha_s[s], &c1[2 * s]);
    }
    // p_s = zeta_s r + alpha_s p_s for every live shift, r read once
    if (live > 0)
      GLBX(glb_cgm_update_p(A.ctx, Traits<T>::dtype, size, live, c0.data(), c1.data(), r, (void* const*)p_s.data()));
    B.xpay(r, alpha, p);
    pAp = A.apply_dot(Ap, p, p);
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
  std::vector<double> relres(n_shift);
  for (int s = 0; s < n_shift; s++) {
    A.apply(Ap, phi[s]);
    B.axpy(T(shifts[s]), phi[s], Ap);  // Ap = Ap + shift*phi  (generic_cg_m.cpp:568)
    inf.resSqmrhs[s] = B.diffnorm2sq(Ap, b);
    relres[s] = sqrt(inf.resSqmrhs[s]) / bsqrt;
  }
  inf.ops_count = A.ops;
  print_verbosity_summary_multi(verb, "CG-M", inf.success, k, inf.ops_count, relres.data(), n_shift);
  inf.resSq = 0.0;  // generic_cg_m.cpp:596 returns a truersq that is never assigned
  inf.iter = k;
  inf.name = "CG-M";
  return inf;
}

// ------------------------------------------------------------------------------------------ restarts
// e.g. generic_cg.cpp:235-275.  GMRES's final success test compares an ABSOLUTE residual
// (generic_gmres.cpp:807); every other solver a relative one.
template <typename T, typename Inner>
inversion_info restarted(const std::string& label, T* b, int size, int max_iter, double res, glb_context* ctx,
                         inversion_verbose_struct* verb, bool absolute_quirk, Inner inner) {
  Blas<T> B = {ctx, (size_t)size};
  const double bsqrt = sqrt(B.norm2sq(b));
  inversion_verbose_struct verb_rest;
  shuffle_verbosity_restart(&verb_rest, verb);
  inversion_info inf;
  int iter = 0, ops = 0;
  do {
    inf = inner(&verb_rest);
    iter += inf.iter;
    ops += inf.ops_count;
    print_verbosity_restart(verb, label, iter, ops, sqrt(inf.resSq) / bsqrt);
  } while (iter < max_iter && inf.success == false && sqrt(inf.resSq) / bsqrt > res);
  inf.iter = iter;
  inf.ops_count = ops;
  print_verbosity_summary(verb, label, inf.success, iter, inf.ops_count, sqrt(inf.resSq) / bsqrt);
  inf.name = label;
  if (absolute_quirk)
    inf.success = !(sqrt(inf.resSq) > res);
  else
    inf.success = !(sqrt(inf.resSq) / bsqrt > res);
  return inf;
}

template <typename T>
glb_context* ctx_of(void (*fn)(T*, T*, void*), void* extra) {
  void (*std_cb)(T*, T*, void*) = &glb200_apply_dev;
  return (fn == std_cb) ? glb_op_context((glb_operator*)extra) : glb200_default_context();
}

std::string label(const char* base, int rf) {
  std::ostringstream ss;
  ss << base << "(" << rf << ")";
  return ss.str();
}

}  // namespace

void glb200_force_host_scalars(bool force) { g_force_host_scalars = force; }

// ------------------------------------------------------------------------------------------ exports
#define GLB200_DEF_BASIC(NAME, CORE, ALG)                                                                         \
  inversion_info NAME(double* phi, double* phi0, int size, int max_iter, double res,                              \
                      void (*mv)(double*, double*, void*), void* extra, inversion_verbose_struct* verb) {         \
    try {                                                                                                         \
      return CORE<double>(phi, phi0, size, max_iter, res, mv, extra, verb);                                       \
    } catch (const std::exception& e) {                                                                           \
      return failed(ALG, e);                                                                                      \
    }                                                                                                             \
  }                                                                                                               \
  inversion_info NAME(zcplx* phi, zcplx* phi0, int size, int max_iter, double res, void (*mv)(zcplx*, zcplx*, void*), \
                      void* extra, inversion_verbose_struct* verb) {                                              \
    try {                                                                                                         \
      return CORE<zcplx>(phi, phi0, size, max_iter, res, mv, extra, verb);                                        \
    } catch (const std::exception& e) {                                                                           \
      return failed(ALG, e);                                                                                      \
    }                                                                                                             \
  }

#define GLB200_DEF_RESTART(NAME, CORE, ALG, QUIRK)                                                                \
  inversion_info NAME(double* phi, double* phi0, int size, int max_iter, double res, int rf,                      \
                      void (*mv)(double*, double*, void*), void* extra, inversion_verbose_struct* verb) {         \
    try {                                                                                                         \
      return restarted<double>(label(ALG, rf), phi0, size, max_iter, res, ctx_of<double>(mv, extra), verb, QUIRK, \
                               [&](inversion_verbose_struct* v) {                                                 \
                                 return CORE<double>(phi, phi0, size, rf, res, mv, extra, v);                     \
                               });                                                                                \
    } catch (const std::exception& e) {                                                                           \
      return failed(ALG, e);                                                                                      \
    }                                                                                                             \
  }                                                                                                               \
  inversion_info NAME(zcplx* phi, zcplx* phi0, int size, int max_iter, double res, int rf,                        \
                      void (*mv)(zcplx*, zcplx*, void*), void* extra, inversion_verbose_struct* verb) {           \
    try {                                                                                                         \
      return restarted<zcplx>(label(ALG, rf), phi0, size, max_iter, res, ctx_of<zcplx>(mv, extra), verb, QUIRK,   \
                              [&](inversion_verbose_struct* v) {                                                  \
                                return CORE<zcplx>(phi, phi0, size, rf, res, mv, extra, v);                       \
                              });                                                                                 \
    } catch (const std::exception& e) {                                                                           \
      return failed(ALG, e);                                                                                      \
    }                                                                                                             \
  }

GLB200_DEF_BASIC(minv_vector_cg_dev, cg_dev, "CG")
GLB200_DEF_RESTART(minv_vector_cg_restart_dev, cg_dev, "CG", false)
GLB200_DEF_BASIC(minv_vector_cr_dev, cr_dev, "CR")
GLB200_DEF_RESTART(minv_vector_cr_restart_dev, cr_dev, "CR", false)
GLB200_DEF_BASIC(minv_vector_gcr_dev, gcr_dev, "GCR")
GLB200_DEF_RESTART(minv_vector_gcr_restart_dev, gcr_dev, "GCR", false)
GLB200_DEF_BASIC(minv_vector_bicgstab_dev, bicgstab_dev, "BiCGStab")
GLB200_DEF_RESTART(minv_vector_bicgstab_restart_dev, bicgstab_dev, "BiCGStab", false)
GLB200_DEF_BASIC(minv_vector_gmres_dev, gmres_dev, "GMRES")
GLB200_DEF_RESTART(minv_vector_gmres_restart_dev, gmres_dev, "GMRES", true)

#define GLB200_DEF_BICGL(T)                                                                                       \
  inversion_info minv_vector_bicgstab_l_dev(T* phi, T* phi0, int size, int max_iter, double res, int l,           \
                                            void (*mv)(T*, T*, void*), void* extra,                               \
                                            inversion_verbose_struct* verb) {                                     \
    try {                                                                                                         \
      return bicgstab_l_dev<T>(phi, phi0, size, max_iter, res, l, mv, extra, verb);                               \
    } catch (const std::exception& e) {                                                                           \
      return failed("BiCGStab-l", e);                                                                             \
    }                                                                                                             \
  }                                                                                                               \
  inversion_info minv_vector_bicgstab_l_restart_dev(T* phi, T* phi0, int size, int max_iter, double res, int rf,  \
                                                    int l, void (*mv)(T*, T*, void*), void* extra,                \
                                                    inversion_verbose_struct* verb) {                             \
    try {                                                                                                         \
      std::ostringstream ss;                                                                                      \
      ss << "BiCGStab-" << l << "(" << rf << ")";                                                                 \
      return restarted<T>(ss.str(), phi0, size, max_iter, res, ctx_of<T>(mv, extra), verb, false,                 \
                          [&](inversion_verbose_struct* v) {                                                      \
                            return bicgstab_l_dev<T>(phi, phi0, size, rf, res, l, mv, extra, v);                  \
                          });                                                                                     \
    } catch (const std::exception& e) {                                                                           \
      return failed("BiCGStab-l", e);                                                                             \
    }                                                                                                             \
  }                                                                                                               \
  inversion_info minv_vector_cg_m_dev(T** phi, T* phi0, int n_shift, int size, int resid_freq_check, int max_iter, \
                                      double eps, double* shifts, void (*mv)(T*, T*, void*), void* extra,         \
                                      bool worst_first, inversion_verbose_struct* verb) {                         \
    try {                                                                                                         \
      return cg_m_dev<T>(phi, phi0, n_shift, size, resid_freq_check, max_iter, eps, shifts, mv, extra,            \
                         worst_first, verb);                                                                      \
    } catch (const std::exception& e) {                                                                           \
      return failed("CG-M", e);                                                                                   \
    }                                                                                                             \
  }
GLB200_DEF_BICGL(double)
GLB200_DEF_BICGL(zcplx)

// ------------------------------------------------------------------------------------------ VPGCR exports
#define GLB200_DEF_VPGCR(T)                                                                                        \
  inversion_info minv_vector_gcr_var_precond_dev(T* phi, T* phi0, int size, int max_iter, double res,              \
                                                 void (*mv)(T*, T*, void*), void* extra,                           \
                                                 void (*pc)(T*, T*, int, void*, inversion_verbose_struct*),        \
                                                 void* pc_info, inversion_verbose_struct* verb) {                  \
    try {                                                                                                          \
      return gcr_var_precond_dev<T>(phi, phi0, size, max_iter, res, mv, extra, pc, pc_info, verb);                 \
    } catch (const std::exception& e) {                                                                            \
      return failed("VPGCR", e);                                                                                   \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_gcr_var_precond_restart_dev(T* phi, T* phi0, int size, int max_iter, double res,      \
                                                         int rf, void (*mv)(T*, T*, void*), void* extra,           \
                                                         void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), \
                                                         void* pc_info, inversion_verbose_struct* verb) {          \
    try {                                                                                                          \
      return restarted<T>(label("Variably Preconditioned Restarted GCR", rf), phi0, size, max_iter, res,           \
                          ctx_of<T>(mv, extra), verb, false, [&](inversion_verbose_struct* v) {                    \
                            return gcr_var_precond_dev<T>(phi, phi0, size, rf, res, mv, extra, pc, pc_info, v);    \
                          });                                                                                      \
    } catch (const std::exception& e) {                                                                            \
      return failed("VPGCR", e);                                                                                   \
    }                                                                                                              \
  }
GLB200_DEF_VPGCR(double)
GLB200_DEF_VPGCR(zcplx)

namespace {
// ------------------------------------------------------------------------------------------ SOR, MinRes
// generic_sor.cpp:24-117 / :122-209.  x_{n+1} = x_n + omega (b - A x_n); the convergence test looks at the residual
// of x_n, and on convergence the reference returns x_n (its xnew is dropped), so x is advanced after the test.
template <typename T>
inversion_info sor_dev(T* x, T* b, int size, int max_iter, double eps, double omega, void (*fn)(T*, T*, void*),
                       void* extra, inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  T *Ax = W.get(), *t = W.get();
  std::ostringstream ss;
  ss << "SOR_" << omega;
  const double bsqrt = sqrt(B.norm2sq(b));
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    A.apply(Ax, x);
    rsq = B.diffnorm2sq(Ax, b);  // |check|^2 with check = Ax - b
    print_verbosity_resid(verb, ss.str(), k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) break;
    B.sub(b, Ax, t);
    B.axpy(T(omega), t, x);  // x = x + omega (b - Ax)
  }
  inf.success = (k != max_iter);
  if (inf.success && !IsComplex<T>::value) k++;  // only the real overload counts the last iteration (:83 vs :180)
  A.apply(Ax, x);
  const double truersq = B.diffnorm2sq(Ax, b);
  inf.ops_count = A.ops;
  // generic_sor.cpp:104: the summary line is printed from invif.resSq before it is assigned (0)
  print_verbosity_summary(verb, ss.str(), inf.success, k, inf.ops_count, sqrt(inf.resSq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  std::ostringstream ss2;
  ss2 << "SOR omega=" << omega;
  inf.name = ss2.str();
  return inf;
}

// generic_minres.cpp:22-117 / :128-232 (Saad 5.3.2 with a relaxation factor): p = A r, alpha = omega <p,r>/|p|^2,
// x += alpha r, r -= alpha p.  <p,r> and |p|^2 come from one pass, the update and |r|^2 from another.
template <typename T>
inversion_info minres_dev(T* x, T* b, int size, int max_iter, double eps, double omega, void (*fn)(T*, T*, void*),
                          void* extra, inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  T *p = W.get(), *r = W.get();
  std::ostringstream ss;
  ss << "MR_" << omega;
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(p, x);
  B.sub(b, p, r);
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    A.apply(p, r);
    double dn[3];
    GLBX(glb_dot_norm(A.ctx, Traits<T>::dtype, size, p, r, dn));  // <p,r>, |p|^2
    T alpha = Traits<T>::unpack(dn) / dn[2];
    alpha *= omega;
    rsq = B.update_xr_norm(alpha, r, x, -alpha, p, r);
    print_verbosity_resid(verb, ss.str(), k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) break;
  }
  inf.success = (k != max_iter);
  if (inf.success && !IsComplex<T>::value) k++;  // :91 vs :204
  A.apply(p, x);
  const double truersq = B.diffnorm2sq(p, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, ss.str(), inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "Minimum Residual (MinRes)";
  return inf;
}

}  // namespace

#define GLB200_DEF_RELAX(T)                                                                                       \
  inversion_info minv_vector_sor_dev(T* phi, T* phi0, int size, int max_iter, double eps, double omega,           \
                                     void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {    \
    try {                                                                                                         \
      return sor_dev<T>(phi, phi0, size, max_iter, eps, omega, mv, extra, verb);                                  \
    } catch (const std::exception& e) {                                                                           \
      return failed("SOR", e);                                                                                    \
    }                                                                                                             \
  }                                                                                                               \
  inversion_info minv_vector_minres_dev(T* phi, T* phi0, int size, int max_iter, double eps, double omega,        \
                                        void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) { \
    try {                                                                                                         \
      return minres_dev<T>(phi, phi0, size, max_iter, eps, omega, mv, extra, verb);                               \
    } catch (const std::exception& e) {                                                                           \
      return failed("MinRes", e);                                                                                 \
    }                                                                                                             \
  }                                                                                                               \
  inversion_info minv_vector_minres_dev(T* phi, T* phi0, int size, int max_iter, double eps,                      \
                                        void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) { \
    return minv_vector_minres_dev(phi, phi0, size, max_iter, eps, 1.0, mv, extra, verb);                          \
  }
GLB200_DEF_RELAX(double)
GLB200_DEF_RELAX(zcplx)

// ------------------------------------------------------------------------------------------ power iteration
// generic_poweriter.cpp:23-88
eigenvalue_info eig_vector_poweriter_dev(double* eig, double* phi0, int size, int max_iter, double relres,
                                         void (*fn)(double*, double*, void*), void* extra) {
  eigenvalue_info eigif;
  eigif.relative_diff = 0.0;
  eigif.iter = 0;
  eigif.success = false;
  eigif.name = "Power Iteration";
  try {
    DevOp<double> A = make_op<double>(fn, extra, size);
    Blas<double> B = {A.ctx, (size_t)size};
    Work<double> W(B);
    double *x = W.get(), *q = W.get();
    B.copy(x, phi0);
    double beta = 0.0, beta_new = sqrt(B.norm2sq(x));
    B.rdiv(x, beta_new, q);
    int k;
    for (k = 0; k < max_iter; k++) {
      beta = beta_new;
      A.apply(x, q);
      beta_new = sqrt(B.norm2sq(x));
      if (fabs(beta - beta_new) < relres) break;
      B.rdiv(x, beta_new, q);
    }
    eigif.success = (k != max_iter);
    *eig = beta_new;
    eigif.relative_diff = fabs(beta - beta_new);
    eigif.iter = k;
  } catch (const std::exception& e) {
    std::cerr << "[glb200] Power Iteration aborted: " << e.what() << std::endl;
  }
  return eigif;
}

// generic_inverter.cpp:18-190 on device vectors: the enum dispatch the MG smoother uses
template <typename T>
static inversion_info dispatch_dev(T* lhs, T* rhs, int size, minv_inverter type, minv_inverter_params& p,
                                   void (*mv)(T*, T*, void*), void* extra, inversion_verbose_struct* verb) {
  switch (type) {
    case MINV_CG:
      return p.restart ? minv_vector_cg_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_cg_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_CR:
      return p.restart ? minv_vector_cr_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_cr_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_GCR:
      return p.restart ? minv_vector_gcr_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_gcr_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_BICGSTAB:
      return p.restart
                 ? minv_vector_bicgstab_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                 : minv_vector_bicgstab_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_BICGSTAB_L:
      return p.restart ? minv_vector_bicgstab_l_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq,
                                                            p.bicgstabl_l, mv, extra, verb)
                       : minv_vector_bicgstab_l_dev(lhs, rhs, size, p.max_iters, p.tol, p.bicgstabl_l, mv, extra, verb);
    case MINV_GMRES:
      return p.restart ? minv_vector_gmres_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv, extra, verb)
                       : minv_vector_gmres_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, verb);
    case MINV_SOR:     // restarting makes no sense for these two (generic_inverter.cpp:86-95)
      return minv_vector_sor_dev(lhs, rhs, size, p.max_iters, p.tol, p.sor_omega, mv, extra, verb);
    case MINV_MINRES:
      return minv_vector_minres_dev(lhs, rhs, size, p.max_iters, p.tol, p.minres_omega, mv, extra, verb);
    default:
      return inversion_info();
  }
}
inversion_info minv_unpreconditioned_dev(double* lhs, double* rhs, int size, minv_inverter type, minv_inverter_params& p,
                                         void (*mv)(double*, double*, void*), void* extra,
                                         inversion_verbose_struct* verb) {
  return dispatch_dev<double>(lhs, rhs, size, type, p, mv, extra, verb);
}
inversion_info minv_unpreconditioned_dev(zcplx* lhs, zcplx* rhs, int size, minv_inverter type, minv_inverter_params& p,
                                         void (*mv)(zcplx*, zcplx*, void*), void* extra,
                                         inversion_verbose_struct* verb) {
  return dispatch_dev<zcplx>(lhs, rhs, size, type, p, mv, extra, verb);
}

// =====================================================================================================
// the rest of the solver family on the same kernels -- preconditioned CG / flexible CG /
// BiCGStab (generic_cg_precond.cpp, generic_cg_flex_precond.cpp, generic_bicgstab_precond.cpp), the
// multishift CR and BiCGStab (generic_cr_m.cpp, generic_bicgstab_m.cpp), the enum dispatch
// (generic_inverter_precond.cpp) and the stock preconditioners of generic_precond.cpp.  Pure host-shell
// work: every vector operation is an existing C-ABI call.
// =====================================================================================================
namespace {

// ------------------------------------------------------------------------------------------ PCG
// generic_cg_precond.cpp:23-145 / :150-275
template <typename T>
inversion_info cg_precond_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                              void (*precond)(T*, T*, int, void*, inversion_verbose_struct*), void* precond_info,
                              inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  inversion_verbose_struct verb_prec;
  shuffle_verbosity_precond(&verb_prec, verb);
  T *r = W.get(), *p = W.get(), *Ap = W.get(), *z = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(p, x);
  B.sub(b, p, r);
  B.zero(z);
  precond(z, r, size, precond_info, &verb_prec);
  B.copy(p, z);
  A.apply(Ap, p);
  T zdotr = B.dot(z, r);
  if (IsComplex<T>::value) {  // generic_cg_precond.cpp:199 : the complex overload announces its loop
    printf("Starting loop!\n");
    fflush(stdout);
  }
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    const T alpha = zdotr / B.dot(p, Ap);
    rsq = B.update_xr_norm(alpha, p, x, -alpha, Ap, r);
    print_verbosity_resid(verb, "PCG", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt || k == max_iter - 1) break;
    B.zero(z);
    precond(z, r, size, precond_info, &verb_prec);
    const T zdotr_new = B.dot(r, z);
    const T beta = zdotr_new / zdotr;
    zdotr = zdotr_new;
    B.xpay(z, beta, p);  // p = z + beta p
    A.apply(Ap, p);
  }
  inf.success = !(k == max_iter - 1);
  k++;
  A.apply(Ap, x);
  const double truersq = B.diffnorm2sq(Ap, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "PCG", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "Preconditioned CG";
  return inf;
}

// ------------------------------------------------------------------------------------------ FPCG
// generic_cg_flex_precond.cpp:26-166 / :223-362.  Like VPGCR every direction is kept; the coefficients
// beta_ij = -<Ap_i, z>/<p_i, Ap_i> (:310) all use the same z: one batched pass.
template <typename T>
inversion_info cg_flex_precond_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*), void* extra,
                                   void (*precond)(T*, T*, int, void*, inversion_verbose_struct*), void* precond_info,
                                   inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  inversion_verbose_struct verb_prec;
  shuffle_verbosity_precond(&verb_prec, verb);
  T *r = W.get(), *z = W.get(), *Az = W.get();
  std::vector<const void*> ps, Aps;
  std::vector<T> pAp;
  T* p = W.get();
  T* Ap = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(p, x);
  B.sub(b, p, r);
  B.zero(z);
  precond(z, r, size, precond_info, &verb_prec);
  B.copy(p, z);
  A.apply(Ap, p);
  double rsq = 0.0;
  int k;
  std::vector<double> dots, coef;
  for (k = 0; k < max_iter; k++) {
    ps.push_back(p);
    Aps.push_back(Ap);
    pAp.push_back(B.dot(p, Ap));
    const T alpha = B.dot(p, r) / pAp.back();
    rsq = B.update_xr_norm(alpha, p, x, -alpha, Ap, r);
    print_verbosity_resid(verb, "FPCG", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt || k == max_iter - 1) break;
    B.zero(z);
    precond(z, r, size, precond_info, &verb_prec);
    A.apply(Az, z);
    dots.resize(2 * (k + 1));
    coef.resize(2 * (k + 1));
    GLBX(glb_multi_dot(A.ctx, Traits<T>::dtype, size, k + 1, Aps.data(), z, dots.data()));
    for (int ii = 0; ii <= k; ii++) {
      const T beta = -Traits<T>::unpack(&dots[2 * ii]) / pAp[ii];
      Traits<T>::pack(beta, &coef[2 * ii]);
    }
    p = W.get();
    Ap = W.get();
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, k + 1, coef.data(), ps.data(), z, p));
    GLBX(glb_lincomb(A.ctx, Traits<T>::dtype, size, k + 1, coef.data(), Aps.data(), Az, Ap));
  }
  inf.success = !(k == max_iter - 1);
  k++;
  A.apply(Az, x);
  const double truersq = B.diffnorm2sq(Az, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "FPCG", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "Flexibly Preconditioned CG";
  return inf;
}

// generic_cg_flex_precond.cpp:171-218 / :367-415 : the restart loop spends one more apply on the true residual
template <typename T>
inversion_info cg_flex_precond_restart_dev(T* x, T* b, int size, int max_iter, double res, int rf,
                                           void (*fn)(T*, T*, void*), void* extra,
                                           void (*precond)(T*, T*, int, void*, inversion_verbose_struct*),
                                           void* precond_info, inversion_verbose_struct* verb) {
  glb_context* ctx = ctx_of<T>(fn, extra);
  Blas<T> B = {ctx, (size_t)size};
  const double bsqrt = sqrt(B.norm2sq(b));
  const std::string name = label("Flexibly Preconditioned Restarted CG", rf);
  inversion_verbose_struct verb_rest;
  shuffle_verbosity_restart(&verb_rest, verb);
  inversion_info inf;
  int iter = 0, ops = 0;
  do {
    inf = cg_flex_precond_dev<T>(x, b, size, rf, res, fn, extra, precond, precond_info, &verb_rest);
    iter += inf.iter;
    ops += inf.ops_count;
    print_verbosity_restart(verb, name, iter, ops, sqrt(inf.resSq) / bsqrt);
  } while (iter < max_iter && inf.success == false && sqrt(inf.resSq) / bsqrt > res);
  DevOp<T> A = make_op<T>(fn, extra, size);
  Work<T> W(B);
  T* Ax = W.get();
  A.apply(Ax, x);
  ops++;
  const double truersq = B.diffnorm2sq(Ax, b);
  inf.resSq = truersq;
  if (IsComplex<T>::value) {  // :398 prints before, the real overload (:209-211) after the counts are stored
    print_verbosity_summary(verb, name, inf.success, iter, inf.ops_count, sqrt(truersq) / bsqrt);
    inf.iter = iter;
    inf.ops_count = ops;
  } else {
    inf.iter = iter;
    inf.ops_count = ops;
    print_verbosity_summary(verb, name, inf.success, iter, inf.ops_count, sqrt(truersq) / bsqrt);
  }
  inf.name = name;
  inf.success = !(sqrt(inf.resSq) / bsqrt > res);
  return inf;
}

// ------------------------------------------------------------------------------------------ PBiCGStab
// generic_bicgstab_precond.cpp:22-176 / :226-380 (flexible BiCGStab).  Quirks kept: the loop has no
// k == max_iter-1 exit, failure is k == max_iter and then k is NOT incremented; name "BiCGStab".
template <typename T>
inversion_info bicgstab_precond_dev(T* x, T* b, int size, int max_iter, double eps, void (*fn)(T*, T*, void*),
                                    void* extra, void (*precond)(T*, T*, int, void*, inversion_verbose_struct*),
                                    void* precond_info, inversion_verbose_struct* verb) {
  inversion_info inf;
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  inversion_verbose_struct verb_prec;
  shuffle_verbosity_precond(&verb_prec, verb);
  T *r = W.get(), *r0 = W.get(), *p = W.get(), *s = W.get(), *pt = W.get(), *Apt = W.get(), *st = W.get(),
    *Ast = W.get();
  const double bsqrt = sqrt(B.norm2sq(b));
  A.apply(Apt, x);
  B.sub(b, Apt, r);
  B.copy(r0, r);
  B.copy(p, r);
  T rho = B.dot(r0, r);
  double rsq = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    B.zero(pt);
    precond(pt, p, size, precond_info, &verb_prec);
    const T r0Apt = A.apply_dot(Apt, pt, r0);
    const T alpha = rho / r0Apt;
    B.axpyz(-alpha, Apt, r, s);  // s = r - alpha A ptilde
    B.zero(st);
    precond(st, s, size, precond_info, &verb_prec);
    double AsAs = 0.0;
    const T sAst = A.apply_dot_norm(Ast, st, s, &AsAs);  // <s, A stilde>, |A stilde|^2
    const T omega = sAst / T(AsAs);                      // :316 dot(s,Astilde)/dot(Astilde,Astilde)
    // x = x + alpha ptilde + omega stilde  (:319-322: one expression, left to right) ; r = s - omega A stilde
    B.axpy(alpha, pt, x);
    B.axpy(omega, st, x);
    B.axpyz(-omega, Ast, s, r);
    rsq = B.norm2sq(r);
    print_verbosity_resid(verb, "Preconditioned BiCGStab", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) break;
    const T rhoNew = B.dot(r0, r);
    const T beta = rhoNew / rho * (alpha / omega);
    rho = rhoNew;
    // p = r + beta (p - omega A ptilde)
    double cb[2], cmo[2];
    Traits<T>::pack(beta, cb);
    Traits<T>::pack(omega, cmo);
    GLBX(glb_bicgstab_pupdate(A.ctx, Traits<T>::dtype, size, r, cb, cmo, Apt, p));
  }
  if (k == max_iter) {
    inf.success = false;
  } else {
    k++;
    inf.success = true;
  }
  A.apply(Apt, x);
  const double truersq = B.diffnorm2sq(Apt, b);
  inf.ops_count = A.ops;
  print_verbosity_summary(verb, "Preconditioned BiCGStab", inf.success, k, inf.ops_count, sqrt(truersq) / bsqrt);
  inf.resSq = truersq;
  inf.iter = k;
  inf.name = "BiCGStab";
  return inf;
}

}  // namespace

namespace {

// shared tail of the multishift solvers: undo the permutation of phi[] / shifts[] (generic_cg_m.cpp:533-558)
template <typename T>
void undo_shift_permutation(T** phi, double* shifts, std::vector<int>& mapping, int n_shift) {
  for (int s = 0; s < n_shift; s++) {
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
}

// ------------------------------------------------------------------------------------------ CR-M
// generic_cr_m.cpp:24-318 / :323-622.  Same shifted recurrences as CG-M (zeta, beta_s, alpha_s) on top of the
// CR base iteration; both overloads test abs(zeta).
template <typename T>
inversion_info cr_m_dev(T** phi, T* b, int n_shift, int size, int check_every, int max_iter, double eps, double* shifts,
                        void (*fn)(T*, T*, void*), void* extra, bool worst_first, inversion_verbose_struct* verb) {
  inversion_info inf(n_shift);
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  std::vector<T> alpha_s(n_shift, T(0.0)), beta_s(n_shift, T(1.0)), zeta_s(n_shift, T(1.0)), zeta_prev(n_shift, T(1.0));
  std::vector<T*> p_s(n_shift);
  std::vector<int> mapping(n_shift);
  for (int s = 0; s < n_shift; s++) {
    p_s[s] = W.get();
    mapping[s] = s;
  }
  T *r = W.get(), *Ar = W.get(), *p = W.get(), *Ap = W.get();
  int live = n_shift;
  T beta = 1.0, alpha = 0.0, beta_prev;
  const double bsqrt = sqrt(B.norm2sq(b));
  for (int s = 0; s < n_shift; s++) {
    B.copy(p_s[s], b);
    B.zero(phi[s]);
  }
  B.copy(p, b);
  B.copy(r, b);
  A.apply(Ap, p);
  B.copy(Ar, Ap);
  double Apsq = B.norm2sq(Ap);
  double rsq = B.norm2sq(r);
  std::vector<double> c0, c1;
  int k;
  for (k = 0; k < max_iter; k++) {
    beta_prev = beta;
    beta = -B.dot(Ap, r) / Apsq;
    c0.resize(2 * live);
    for (int s = 0; s < live; s++) {
      const T z_old = zeta_s[s];
      zeta_s[s] = (zeta_s[s] * zeta_prev[s] * beta_prev) /
                  (beta * alpha * (zeta_prev[s] - zeta_s[s]) + zeta_prev[s] * beta_prev * (1.0 - shifts[s] * beta));
      zeta_prev[s] = z_old;
      beta_s[s] = beta * zeta_s[s] / zeta_prev[s];
      Traits<T>::pack(beta_s[s], &c0[2 * s]);
    }
    if (live > 0)
      GLBX(glb_cgm_update_x(A.ctx, Traits<T>::dtype, size, live, c0.data(), (const void* const*)p_s.data(),
                            (void* const*)phi));
    rsq = B.axpy_norm(beta, Ap, r);  // r = r + beta Ap ; |r|^2
    print_verbosity_resid(verb, "CR-M", k + 1, A.ops, sqrt(rsq) / bsqrt);
    if (k % check_every == 0) {
      for (int s = 0; s < live; s++) {
        if (std::abs(zeta_s[s]) * sqrt(rsq) < eps * bsqrt) {
          live--;
          if (live != s) {
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
    if ((worst_first && std::abs(zeta_s[0]) * sqrt(rsq) < eps * bsqrt) || live == 0 || k == max_iter - 1) break;
    const T ApAr = A.apply_dot(Ar, r, Ap);  // Ar = A r, <Ap, Ar>
    alpha = -ApAr / Apsq;
    c0.resize(2 * live);
    c1.resize(2 * live);
    for (int s = 0; s < live; s++) {
      alpha_s[s] = alpha * zeta_s[s] * beta_s[s] / (zeta_prev[s] * beta);
      Traits<T>::pack(zeta_s[s], &c0[2 * s]);
      Traits<T>::pack(alpha_s[s], &c1[2 * s]);
    }
    if (live > 0)
      GLBX(glb_cgm_update_p(A.ctx, Traits<T>::dtype, size, live, c0.data(), c1.data(), r, (void* const*)p_s.data()));
    double ca[2];
    Traits<T>::pack(alpha, ca);
    // p = r + alpha p ; Ap = Ar + alpha Ap ; |Ap|^2
    GLBX(glb_update_p_ap_norm(A.ctx, Traits<T>::dtype, size, r, Ar, ca, p, Ap, &Apsq));
  }
  inf.success = !(k == max_iter - 1);
  k++;
  undo_shift_permutation(phi, shifts, mapping, n_shift);
  std::vector<double> relres(n_shift);
  for (int s = 0; s < n_shift; s++) {
    A.apply(Ap, phi[s]);
    B.axpy(T(shifts[s]), phi[s], Ap);
    inf.resSqmrhs[s] = B.diffnorm2sq(Ap, b);
    relres[s] = sqrt(inf.resSqmrhs[s]) / bsqrt;
  }
  inf.ops_count = A.ops;
  print_verbosity_summary_multi(verb, "CR-M", inf.success, k, inf.ops_count, relres.data(), n_shift);
  inf.resSq = 0.0;  // truersq is never assigned (generic_cr_m.cpp:617)
  inf.iter = k;
  inf.name = "CR-M";
  return inf;
}

inline bool is_nan_d(double v) { return v != v; }

// ------------------------------------------------------------------------------------------ BiCGStab-M
// generic_bicgstab_m.cpp:26-397 / :402-774 (Jegerlehner's shifted BiCGStab).  w^dag = conj(r_0) in the complex
// overload (:494), r_0 in the real one; dot() conjugates its first argument, so <w^dag, v> = sum r_0 v.
template <typename T>
inversion_info bicgstab_m_dev(T** phi, T* b, int n_shift, int size, int check_every, int max_iter, double eps,
                              double* shifts, void (*fn)(T*, T*, void*), void* extra, bool worst_first,
                              inversion_verbose_struct* verb) {
  inversion_info inf(n_shift);
  DevOp<T> A = make_op<T>(fn, extra, size);
  Blas<T> B = {A.ctx, (size_t)size};
  Work<T> W(B);
  const T one = 1.0, zero = 0.0;
  std::vector<T> alpha_s(n_shift, zero), beta_s(n_shift, one), zeta_s(n_shift, one), zeta_prev(n_shift, one),
      chi_s(n_shift, zero), rho_s(n_shift, one), rho_prev(n_shift, one);
  std::vector<T*> s_s(n_shift);
  std::vector<int> mapping(n_shift);
  for (int n = 0; n < n_shift; n++) {
    s_s[n] = W.get();
    mapping[n] = n;
  }
  T *r = W.get(), *r_prev = W.get(), *s = W.get(), *As = W.get(), *w = W.get(), *Aw = W.get(), *wdag = W.get();
  int live = n_shift;
  T beta = 1.0, alpha = 0.0, beta_prev, chi = 0.0;
  bool breakdown = false;
  const double bsqrt = sqrt(B.norm2sq(b));
  for (int n = 0; n < n_shift; n++) {
    B.copy(s_s[n], b);
    B.zero(phi[n]);
  }
  B.copy(s, b);
  B.copy(r, b);
  B.copy(r_prev, r);
  A.apply(As, s);
  GLBX(glb_conj(A.ctx, Traits<T>::dtype, size, r, wdag));  // wdag = conj(r)  (a copy for real fields)
  T delta = B.dot(wdag, r), delta_prev = delta;
  T psi = B.dot(wdag, As) / delta;
  double rsqNew = 0.0;
  int k;
  for (k = 0; k < max_iter; k++) {
    beta_prev = beta;
    beta = -1.0 / psi;
    B.axpyz(beta, As, r, w);  // w = r + beta As
    double AwAw = 0.0;
    const T wAw = A.apply_dot_norm(Aw, w, w, &AwAw);  // <w, Aw>, |Aw|^2
    chi = conj_of(wAw) / AwAw;                        // :521 dot(Aw, w)/norm2sq(Aw)
    B.copy(r_prev, r);
    B.axpyz(-chi, Aw, w, r);  // r = w - chi Aw
    for (int n = 0; n < live; n++) {
      const T z_old = zeta_s[n];
      zeta_s[n] = (zeta_s[n] * zeta_prev[n] * beta_prev) /
                  (beta * alpha * (zeta_prev[n] - zeta_s[n]) + zeta_prev[n] * beta_prev * (1.0 - shifts[n] * beta));
      zeta_prev[n] = z_old;
      beta_s[n] = beta * zeta_s[n] / zeta_prev[n];
      chi_s[n] = chi / (1.0 + chi * shifts[n]);
      const T rho_old = rho_s[n];
      rho_s[n] = rho_s[n] / (1.0 + chi * shifts[n]);
      rho_prev[n] = rho_old;
      // x_n = x_n - beta_n s_n + (chi_n rho_n^prev zeta_n) w   (:552, left to right)
      B.axpy(-beta_s[n], s_s[n], phi[n]);
      B.axpy(chi_s[n] * rho_prev[n] * zeta_s[n], w, phi[n]);
    }
    rsqNew = B.norm2sq(r);
    print_verbosity_resid(verb, "BICGSTAB-M", k + 1, A.ops, sqrt(rsqNew) / bsqrt);
    if (k % check_every == 0) {
      for (int n = 0; n < live; n++) {
        if (std::abs(zeta_s[n] * rho_s[n]) * sqrt(rsqNew) < eps * bsqrt) {
          live--;
          if (live != n) {
            std::swap(mapping[live], mapping[n]);
            std::swap(phi[live], phi[n]);
            std::swap(s_s[live], s_s[n]);
            std::swap(alpha_s[live], alpha_s[n]);
            std::swap(beta_s[live], beta_s[n]);
            std::swap(zeta_s[live], zeta_s[n]);
            std::swap(zeta_prev[live], zeta_prev[n]);
            std::swap(chi_s[live], chi_s[n]);
            std::swap(rho_s[live], rho_s[n]);
            std::swap(rho_prev[live], rho_prev[n]);
            std::swap(shifts[live], shifts[n]);
            n--;
          }
        }
      }
    }
    if ((worst_first && (std::abs(zeta_s[0] * rho_s[0]) * sqrt(rsqNew) < eps * bsqrt)) || is_nan_d(rsqNew) || live == 0 ||
        k == max_iter - 1)
      break;
    delta_prev = delta;
    delta = B.dot(wdag, r);
    alpha = -beta * delta / (delta_prev * chi);
    for (int n = 0; n < live; n++) alpha_s[n] = alpha * zeta_s[n] * beta_s[n] / (zeta_prev[n] * beta);
    // s = r + alpha (s - chi As)
    double ca[2], cc[2];
    Traits<T>::pack(alpha, ca);
    Traits<T>::pack(chi, cc);
    GLBX(glb_bicgstab_pupdate(A.ctx, Traits<T>::dtype, size, r, ca, cc, As, s));
    for (int n = 0; n < live; n++) {
      if (std::abs(shifts[n]) != 0.0) {
        // s_n = (zeta_n rho_n) r + alpha_n (s_n - (chi_n/beta_n) ((zeta_n rho_n^prev) w - (zeta_n^prev rho_n^prev) r_prev))
        double c[10];
        Traits<T>::pack(zeta_s[n] * rho_s[n], c);
        Traits<T>::pack(alpha_s[n], c + 2);
        Traits<T>::pack(chi_s[n] / beta_s[n], c + 4);
        Traits<T>::pack(zeta_s[n] * rho_prev[n], c + 6);
        Traits<T>::pack(zeta_prev[n] * rho_prev[n], c + 8);
        GLBX(glb_bicgstabm_update_s(A.ctx, Traits<T>::dtype, size, c, r, w, r_prev, s_s[n]));
      } else {
        B.copy(s_s[n], s);
      }
    }
    A.apply(As, s);
    psi = B.dot(wdag, As) / delta;
    if (std::abs(psi) == 0) {
      breakdown = true;
      break;
    }
  }
  inf.success = !(k == max_iter - 1 || breakdown);
  k++;
  undo_shift_permutation(phi, shifts, mapping, n_shift);
  std::vector<double> relres(n_shift);
  for (int n = 0; n < n_shift; n++) {
    A.apply(As, phi[n]);
    B.axpy(T(shifts[n]), phi[n], As);
    inf.resSqmrhs[n] = B.diffnorm2sq(As, b);
    relres[n] = sqrt(inf.resSqmrhs[n]) / bsqrt;
  }
  inf.ops_count = A.ops;
  print_verbosity_summary_multi(verb, "BICGSTAB-M", inf.success, k, inf.ops_count, relres.data(), n_shift);
  inf.resSq = 0.0;
  inf.iter = k;
  inf.name = "BICGSTAB-M";
  return inf;
}

}  // namespace

// ------------------------------------------------------------------------------------------ 8f-4 exports
#define GLB200_DEF_PRECOND_FAMILY(T)                                                                               \
  inversion_info minv_vector_cg_precond_dev(T* phi, T* phi0, int size, int max_iter, double eps,                   \
                                            void (*mv)(T*, T*, void*), void* extra,                                \
                                            void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), void* pci,  \
                                            inversion_verbose_struct* verb) {                                      \
    try {                                                                                                          \
      return cg_precond_dev<T>(phi, phi0, size, max_iter, eps, mv, extra, pc, pci, verb);                          \
    } catch (const std::exception& e) {                                                                            \
      return failed("PCG", e);                                                                                     \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_cg_flex_precond_dev(T* phi, T* phi0, int size, int max_iter, double eps,              \
                                                 void (*mv)(T*, T*, void*), void* extra,                           \
                                                 void (*pc)(T*, T*, int, void*, inversion_verbose_struct*),        \
                                                 void* pci, inversion_verbose_struct* verb) {                      \
    try {                                                                                                          \
      return cg_flex_precond_dev<T>(phi, phi0, size, max_iter, eps, mv, extra, pc, pci, verb);                     \
    } catch (const std::exception& e) {                                                                            \
      return failed("FPCG", e);                                                                                    \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_cg_flex_precond_restart_dev(T* phi, T* phi0, int size, int max_iter, double res,      \
                                                         int rf, void (*mv)(T*, T*, void*), void* extra,           \
                                                         void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), \
                                                         void* pci, inversion_verbose_struct* verb) {              \
    try {                                                                                                          \
      return cg_flex_precond_restart_dev<T>(phi, phi0, size, max_iter, res, rf, mv, extra, pc, pci, verb);         \
    } catch (const std::exception& e) {                                                                            \
      return failed("FPCG", e);                                                                                    \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_bicgstab_precond_dev(T* phi, T* phi0, int size, int max_iter, double eps,             \
                                                  void (*mv)(T*, T*, void*), void* extra,                          \
                                                  void (*pc)(T*, T*, int, void*, inversion_verbose_struct*),       \
                                                  void* pci, inversion_verbose_struct* verb) {                     \
    try {                                                                                                          \
      return bicgstab_precond_dev<T>(phi, phi0, size, max_iter, eps, mv, extra, pc, pci, verb);                    \
    } catch (const std::exception& e) {                                                                            \
      return failed("Preconditioned BiCGStab", e);                                                                 \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_bicgstab_precond_restart_dev(T* phi, T* phi0, int size, int max_iter, double res,     \
                                                          int rf, void (*mv)(T*, T*, void*), void* extra,          \
                                                          void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), \
                                                          void* pci, inversion_verbose_struct* verb) {             \
    try {                                                                                                          \
      return restarted<T>(label("Preconditioned Restarted BiCGStab", rf), phi0, size, max_iter, res,               \
                          ctx_of<T>(mv, extra), verb, false, [&](inversion_verbose_struct* v) {                    \
                            return bicgstab_precond_dev<T>(phi, phi0, size, rf, res, mv, extra, pc, pci, v);       \
                          });                                                                                      \
    } catch (const std::exception& e) {                                                                            \
      return failed("Preconditioned BiCGStab", e);                                                                 \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_cr_m_dev(T** phi, T* phi0, int n_shift, int size, int rfc, int max_iter, double eps,  \
                                      double* shifts, void (*mv)(T*, T*, void*), void* extra, bool worst_first,    \
                                      inversion_verbose_struct* verb) {                                            \
    try {                                                                                                          \
      return cr_m_dev<T>(phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra, worst_first, verb);      \
    } catch (const std::exception& e) {                                                                            \
      return failed("CR-M", e);                                                                                    \
    }                                                                                                              \
  }                                                                                                                \
  inversion_info minv_vector_bicgstab_m_dev(T** phi, T* phi0, int n_shift, int size, int rfc, int max_iter,        \
                                            double eps, double* shifts, void (*mv)(T*, T*, void*), void* extra,    \
                                            bool worst_first, inversion_verbose_struct* verb) {                    \
    try {                                                                                                          \
      return bicgstab_m_dev<T>(phi, phi0, n_shift, size, rfc, max_iter, eps, shifts, mv, extra, worst_first, verb); \
    } catch (const std::exception& e) {                                                                            \
      return failed("BICGSTAB-M", e);                                                                              \
    }                                                                                                              \
  }                                                                                                                \
  void identity_preconditioner_dev(T* lhs, T* rhs, int size, void*, inversion_verbose_struct*) {                   \
    glb_vec_copy(glb200_default_context(), Traits<T>::dtype, size, lhs, rhs);                                      \
  }
GLB200_DEF_PRECOND_FAMILY(double)
GLB200_DEF_PRECOND_FAMILY(zcplx)

// generic_precond.cpp:62-77 : n_step GCR iterations on the operator named by the struct
void gcr_preconditioner_dev(double* lhs, double* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  gcr_precond_struct_real* g = (gcr_precond_struct_real*)extra_data;
  minv_vector_gcr_dev(lhs, rhs, size, g->n_step, g->rel_res, g->matrix_vector, g->matrix_extra_data, verb);
}
void gcr_preconditioner_dev(zcplx* lhs, zcplx* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  gcr_precond_struct_complex* g = (gcr_precond_struct_complex*)extra_data;
  minv_vector_gcr_dev(lhs, rhs, size, g->n_step, g->rel_res, g->matrix_vector, g->matrix_extra_data, verb);
}

// generic_precond.cpp:45-59 : n_step MinRes iterations (no relaxation) on the operator named by the struct
void minres_preconditioner_dev(double* lhs, double* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  minres_precond_struct_real* m = (minres_precond_struct_real*)extra_data;
  minv_vector_minres_dev(lhs, rhs, size, m->n_step, m->rel_res, m->matrix_vector, m->matrix_extra_data, verb);
}
void minres_preconditioner_dev(zcplx* lhs, zcplx* rhs, int size, void* extra_data, inversion_verbose_struct* verb) {
  minres_precond_struct_complex* m = (minres_precond_struct_complex*)extra_data;
  minv_vector_minres_dev(lhs, rhs, size, m->n_step, m->rel_res, m->matrix_vector, m->matrix_extra_data, verb);
}

// generic_inverter_precond.cpp:18-114
template <typename T>
static inversion_info dispatch_precond_dev(T* lhs, T* rhs, int size, minv_inverter_precond type,
                                           minv_inverter_precond_params& p, void (*mv)(T*, T*, void*), void* extra,
                                           void (*pc)(T*, T*, int, void*, inversion_verbose_struct*), void* pci,
                                           inversion_verbose_struct* verb) {
  switch (type) {
    case MINV_PRE_CG:  // there is no restarted preconditioned CG (generic_inverter_precond.cpp:22)
      return minv_vector_cg_precond_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    case MINV_PRE_FPCG:
      return p.restart ? minv_vector_cg_flex_precond_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv,
                                                                 extra, pc, pci, verb)
                       : minv_vector_cg_flex_precond_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    case MINV_PRE_VPGCR:
      return p.restart ? minv_vector_gcr_var_precond_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq, mv,
                                                                 extra, pc, pci, verb)
                       : minv_vector_gcr_var_precond_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    case MINV_PRE_BICGSTAB:
      return p.restart ? minv_vector_bicgstab_precond_restart_dev(lhs, rhs, size, p.max_iters, p.tol, p.restart_freq,
                                                                  mv, extra, pc, pci, verb)
                       : minv_vector_bicgstab_precond_dev(lhs, rhs, size, p.max_iters, p.tol, mv, extra, pc, pci, verb);
    default:
      return inversion_info();
  }
}
inversion_info minv_preconditioned_dev(double* lhs, double* rhs, int size, minv_inverter_precond type,
                                       minv_inverter_precond_params& p, void (*mv)(double*, double*, void*), void* extra,
                                       void (*pc)(double*, double*, int, void*, inversion_verbose_struct*), void* pci,
                                       inversion_verbose_struct* verb) {
  return dispatch_precond_dev<double>(lhs, rhs, size, type, p, mv, extra, pc, pci, verb);
}
inversion_info minv_preconditioned_dev(zcplx* lhs, zcplx* rhs, int size, minv_inverter_precond type,
                                       minv_inverter_precond_params& p, void (*mv)(zcplx*, zcplx*, void*), void* extra,
                                       void (*pc)(zcplx*, zcplx*, int, void*, inversion_verbose_struct*), void* pci,
                                       inversion_verbose_struct* verb) {
  return dispatch_precond_dev<zcplx>(lhs, rhs, size, type, p, mv, extra, pc, pci, verb);
}
