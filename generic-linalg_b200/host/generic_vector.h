// generic_vector.h -- HOST-side vector helpers for driver programs: drop-in for the reference's header-only
// generic_vector.h (zero / gaussian / copy / dot / norm2sq / diffnorm2sq / normalize / conj / orthogonal on plain host
// arrays, generic_vector.h:12-249).
//
// These are what a driver uses AROUND a solve -- to clear and fill its host arrays, draw a random right-hand side, or
// check an answer it got back -- and they exist so that a program written against the reference compiles against this
// directory alone.  They are not a compute path of this library: no solver, operator or set-up routine here calls them
// (all vector work of the solvers runs in the CUDA kernels behind include/glb200.h).  Each function performs the same
// left-to-right loop with the same per-element expression as the reference, so values (and the random sequence drawn
// from a given std::mt19937) are the reference's bit for bit.
#ifndef GLB200_GENERIC_VECTOR_H
#define GLB200_GENERIC_VECTOR_H

#include <cmath>
#include <complex>
#include <random>

namespace glb200_hostvec {
// one pass over [0, size) applying f(i)
template <typename F>
inline void each(int size, F f) {
  for (int i = 0; i < size; ++i) f(i);
}
}  // namespace glb200_hostvec

// ---- zero (generic_vector.h:12-32)
template <typename T>
inline void zero(T* v, int size) {
  glb200_hostvec::each(size, [&](int i) { v[i] = 0.0; });
}
template <typename T>
inline void zero(std::complex<T>* v, int size) {
  glb200_hostvec::each(size, [&](int i) { v[i] = 0.0; });
}

// ---- gaussian (generic_vector.h:35-60): unit normal entries; a complex entry is built as complex(draw, draw), whose
// two draws g++ evaluates right to left -- the imaginary part comes first in the generator's sequence
template <typename T>
inline void gaussian(T* v, int size, std::mt19937& generator) {
  std::normal_distribution<> unit_normal(0.0, 1.0);
  glb200_hostvec::each(size, [&](int i) { v[i] = static_cast<T>(unit_normal(generator)); });
}
template <typename T>
inline void gaussian(std::complex<T>* v, int size, std::mt19937& generator) {
  std::normal_distribution<> unit_normal(0.0, 1.0);
  glb200_hostvec::each(size, [&](int i) {
    const T im = static_cast<T>(unit_normal(generator));
    const T re = static_cast<T>(unit_normal(generator));
    v[i] = std::complex<T>(re, im);
  });
}

// ---- copy: dst <- src (generic_vector.h:63-84)
template <typename T>
inline void copy(T* dst, T* src, int size) {
  glb200_hostvec::each(size, [&](int i) { dst[i] = src[i]; });
}
template <typename T>
inline void copy(std::complex<T>* dst, std::complex<T>* src, int size) {
  glb200_hostvec::each(size, [&](int i) { dst[i] = src[i]; });
}

// ---- reductions, summed left to right (generic_vector.h:87-169)
template <typename T>
inline T dot(T* a, T* b, int size) {  // sum a_i b_i
  T acc = (T)0.0;
  glb200_hostvec::each(size, [&](int i) { acc = acc + a[i] * b[i]; });
  return acc;
}
template <typename T>
inline std::complex<T> dot(std::complex<T>* a, std::complex<T>* b, int size) {  // sum conj(a_i) b_i
  std::complex<T> acc = (T)0.0;
  glb200_hostvec::each(size, [&](int i) { acc = acc + std::conj(a[i]) * b[i]; });
  return acc;
}
template <typename T>
inline T norm2sq(T* a, int size) {
  T acc = (T)0.0;
  glb200_hostvec::each(size, [&](int i) { acc = acc + a[i] * a[i]; });
  return acc;
}
template <typename T>
inline T norm2sq(std::complex<T>* a, int size) {
  T acc = (T)0.0;
  glb200_hostvec::each(size, [&](int i) { acc = acc + std::real(std::conj(a[i]) * a[i]); });
  return acc;
}
template <typename T>
inline T diffnorm2sq(T* a, T* b, int size) {  // |a - b|^2
  T acc = (T)0.0;
  glb200_hostvec::each(size, [&](int i) { acc = acc + (a[i] - b[i]) * (a[i] - b[i]); });
  return acc;
}
template <typename T>
inline T diffnorm2sq(std::complex<T>* a, std::complex<T>* b, int size) {
  T acc = (T)0.0;
  glb200_hostvec::each(size, [&](int i) { acc = acc + std::real(std::conj(a[i] - b[i]) * (a[i] - b[i])); });
  return acc;
}

// ---- normalize: v *= 1/sqrt(|v|^2) unless that factor is not positive (generic_vector.h:172-202)
template <typename T>
inline void normalize(T* v, int size) {
  const T scale = 1.0 / sqrt(norm2sq<T>(v, size));
  if (scale > 0.0) glb200_hostvec::each(size, [&](int i) { v[i] *= scale; });
}
template <typename T>
inline void normalize(std::complex<T>* v, int size) {
  const T scale = 1.0 / sqrt(norm2sq<T>(v, size));
  if (scale > 0.0) glb200_hostvec::each(size, [&](int i) { v[i] *= scale; });
}

// ---- conj in place; nothing to do for real data (generic_vector.h:204-219)
template <typename T>
inline void conj(T*, int) {}
template <typename T>
inline void conj(std::complex<T>* v, int size) {
  glb200_hostvec::each(size, [&](int i) { v[i] = std::conj(v[i]); });
}

// ---- orthogonal: a <- a - (<b,a>/|b|^2) b (generic_vector.h:221-249)
template <typename T>
inline void orthogonal(T* a, T* b, int size) {
  const T alpha = -dot<T>(b, a, size) / norm2sq<T>(b, size);
  glb200_hostvec::each(size, [&](int i) { a[i] = a[i] + alpha * b[i]; });
}
template <typename T>
inline void orthogonal(std::complex<T>* a, std::complex<T>* b, int size) {
  const std::complex<T> alpha = -dot<T>(b, a, size) / norm2sq<T>(b, size);
  glb200_hostvec::each(size, [&](int i) { a[i] = a[i] + alpha * b[i]; });
}

#endif
