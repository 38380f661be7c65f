// generic_traits.h -- RealType<T>: the real scalar behind T (drop-in for generic_traits.h:12-63).
#ifndef GLB200_GENERIC_TRAITS_H
#define GLB200_GENERIC_TRAITS_H
#include <complex>

template <typename T>
struct RealType;

#define GLB200_REAL_TRAIT(R)                              \
  template <>                                             \
  struct RealType<R> {                                    \
    typedef R Type;                                       \
    static Type Real(Type v) { return v; }                \
  };                                                      \
  template <>                                             \
  struct RealType<std::complex<R> > {                     \
    typedef R Type;                                       \
    static Type Real(std::complex<R> v) { return v.real(); } \
  };
GLB200_REAL_TRAIT(float)
GLB200_REAL_TRAIT(double)
#undef GLB200_REAL_TRAIT

#endif
