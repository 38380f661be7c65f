// dev_internal.hpp -- thin C++ conveniences over the C ABI (include/glb200.h) used by the solver
// shells.  Nothing here touches CUDA directly: every operation is one glb_* call.
#ifndef GLB200_DEV_INTERNAL_HPP
#define GLB200_DEV_INTERNAL_HPP

#include <complex>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "glb200.h"
#include "glb200_device.h"

namespace glbx {

typedef std::complex<double> zcplx;

struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

inline void check(int rc, const char* what) {
  if (rc != GLB_OK) throw Error(std::string(what) + ": " + glb_last_error());
}
#define GLBX(call) ::glbx::check((call), #call)

template <typename T>
struct Traits;
template <>
struct Traits<double> {
  enum { dtype = GLB_REAL };
  static void pack(double a, double out[2]) {
    out[0] = a;
    out[1] = 0.0;
  }
  static double unpack(const double in[2]) { return in[0]; }
};
template <>
struct Traits<zcplx> {
  enum { dtype = GLB_COMPLEX };
  static void pack(const zcplx& a, double out[2]) {
    out[0] = a.real();
    out[1] = a.imag();
  }
  static zcplx unpack(const double in[2]) { return zcplx(in[0], in[1]); }
};

// The operator a solver works with: a native glb_operator (fused paths) or an opaque device callback.
template <typename T>
struct DevOp {
  glb_context* ctx;
  glb_operator* native;  // non-null when the callback is glb200_apply_dev
  void (*fn)(T*, T*, void*);
  void* extra;
  size_t n;
  int ops;  // operator applications so far (the reference's invif.ops_count)

  void apply(T* out, T* in) {
    if (native)
      GLBX(glb_op_apply(native, out, in));
    else
      fn(out, in, extra);
    ops++;
  }
  // out = A in and <w,out> in the same pass when native
  T apply_dot(T* out, T* in, T* w);
  // out = A in, <w,out> and |out|^2
  T apply_dot_norm(T* out, T* in, T* w, double* nrm);
};

// BLAS-1 bound to one context, vector length and scalar type
template <typename T>
struct Blas {
  glb_context* ctx;
  size_t n;
  enum { dt = Traits<T>::dtype };

  T* alloc() const {
    void* p = 0;
    GLBX(glb_vec_alloc(ctx, dt, n, &p));
    return (T*)p;
  }
  void release(T* p) const {
    if (p) glb_vec_free(ctx, p);
  }
  void zero(T* a) const { GLBX(glb_vec_zero(ctx, dt, n, a)); }
  void copy(T* dst, const T* src) const { GLBX(glb_vec_copy(ctx, dt, n, dst, src)); }
  T dot(const T* a, const T* b) const {
    double o[2];
    GLBX(glb_dot(ctx, dt, n, a, b, o));
    return Traits<T>::unpack(o);
  }
  double norm2sq(const T* a) const {
    double o;
    GLBX(glb_norm2sq(ctx, dt, n, a, &o));
    return o;
  }
  double diffnorm2sq(const T* a, const T* b) const {
    double o;
    GLBX(glb_diffnorm2sq(ctx, dt, n, a, b, &o));
    return o;
  }
  void sub(const T* a, const T* b, T* out) const { GLBX(glb_sub(ctx, dt, n, a, b, out)); }
  void add(const T* a, const T* b, T* out) const { GLBX(glb_add(ctx, dt, n, a, b, out)); }
  void axpy(T a, const T* x, T* y) const {  // y = y + a x
    double c[2];
    Traits<T>::pack(a, c);
    GLBX(glb_axpy(ctx, dt, n, c, x, y));
  }
  void xpay(const T* x, T a, T* y) const {  // y = x + a y
    double c[2];
    Traits<T>::pack(a, c);
    GLBX(glb_xpay(ctx, dt, n, x, c, y));
  }
  void axpyz(T a, const T* x, const T* y, T* z) const {  // z = y + a x
    double c[2];
    Traits<T>::pack(a, c);
    GLBX(glb_axpyz(ctx, dt, n, c, x, y, z));
  }
  void rdiv(const T* x, double d, T* out) const { GLBX(glb_rdiv(ctx, dt, n, x, d, out)); }
  double axpy_norm(T a, const T* x, T* y) const {
    double c[2], o;
    Traits<T>::pack(a, c);
    GLBX(glb_axpy_norm(ctx, dt, n, c, x, y, &o));
    return o;
  }
  double update_xr_norm(T a, const T* p, T* x, T b, const T* q, T* r) const {
    double ca[2], cb[2], o;
    Traits<T>::pack(a, ca);
    Traits<T>::pack(b, cb);
    GLBX(glb_update_xr_norm(ctx, dt, n, ca, p, x, cb, q, r, &o));
    return o;
  }
};

template <typename T>
inline T DevOp<T>::apply_dot(T* out, T* in, T* w) {
  double d[3];
  if (native) {
    GLBX(glb_op_apply_dot(native, out, in, w, 0, d));
    ops++;
    return Traits<T>::unpack(d);
  }
  apply(out, in);  // opaque device callback: separate reduction pass
  GLBX(glb_dot(ctx, Traits<T>::dtype, n, w, out, d));
  return Traits<T>::unpack(d);
}
template <typename T>
inline T DevOp<T>::apply_dot_norm(T* out, T* in, T* w, double* nrm) {
  double d[3];
  if (native) {
    GLBX(glb_op_apply_dot(native, out, in, w, 1, d));
    ops++;
    *nrm = d[2];
    return Traits<T>::unpack(d);
  }
  apply(out, in);
  GLBX(glb_dot(ctx, Traits<T>::dtype, n, w, out, d));
  GLBX(glb_norm2sq(ctx, Traits<T>::dtype, n, out, nrm));
  return Traits<T>::unpack(d);
}

// RAII set of device work vectors
template <typename T>
struct Work {
  const Blas<T>& b;
  std::vector<T*> v;
  explicit Work(const Blas<T>& bl) : b(bl) {}
  T* get() {
    T* p = b.alloc();
    v.push_back(p);
    return p;
  }
  ~Work() {
    for (size_t i = 0; i < v.size(); i++) b.release(v[i]);
  }
};

}  // namespace glbx

#endif
