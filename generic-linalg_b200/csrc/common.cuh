// common.cuh -- device helpers shared by every kernel of libglb200.
//
//  * exact arithmetic: all field arithmetic goes through __dmul_rn/__dadd_rn so that nvcc can
//    never contract a multiply and an add into an FMA.  The reference is compiled for baseline
//    x86-64 (no FMA, reference Makefile:18), so evaluating the same expression tree without
//    contraction reproduces its results bit for bit.  These kernels are HBM-bound (<1 flop/byte),
//    the extra FP64 issue slots are free.
//  * complex<double> is carried as double2 (16-byte vector loads/stores).
//  * reductions: per-thread serial accumulation over a fixed grid-stride, fixed shuffle tree in
//    the warp, fixed tree over the warps, per-block partials, and a deterministic final pass by
//    whichever block arrives last ("last block done") -- run-to-run reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glb {

// ------------------------------------------------------------------ exact scalar ops
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// ------------------------------------------------------------------ field traits
// T = double  (real)   or   T = double2 (complex, x=re, y=im).
typedef double2 cplx;

__device__ __forceinline__ cplx mk(double re, double im) { return make_double2(re, im); }

// std::complex<double> operator* as g++ emits it for finite operands: (ac-bd, ad+bc)
__device__ __forceinline__ cplx fmul(cplx a, cplx b) {
  return mk(xsub(xmul(a.x, b.x), xmul(a.y, b.y)), xadd(xmul(a.x, b.y), xmul(a.y, b.x)));
}
__device__ __forceinline__ double fmul(double a, double b) { return xmul(a, b); }
// conj(a) * b
__device__ __forceinline__ cplx fcmul(cplx a, cplx b) {
  return mk(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}
__device__ __forceinline__ double fcmul(double a, double b) { return xmul(a, b); }
__device__ __forceinline__ cplx fadd(cplx a, cplx b) { return mk(xadd(a.x, b.x), xadd(a.y, b.y)); }
__device__ __forceinline__ double fadd(double a, double b) { return xadd(a, b); }
__device__ __forceinline__ cplx fsub(cplx a, cplx b) { return mk(xsub(a.x, b.x), xsub(a.y, b.y)); }
__device__ __forceinline__ double fsub(double a, double b) { return xsub(a, b); }
__device__ __forceinline__ cplx fneg(cplx a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ double fneg(double a) { return -a; }
// real scalar times field element (double * complex<double> is component-wise)
__device__ __forceinline__ cplx fscale(double s, cplx a) { return mk(xmul(s, a.x), xmul(s, a.y)); }
__device__ __forceinline__ double fscale(double s, double a) { return xmul(s, a); }
// field element divided by a real (complex<double> / double is component-wise)
__device__ __forceinline__ cplx frdiv(cplx a, double d) { return mk(xdiv(a.x, d), xdiv(a.y, d)); }
__device__ __forceinline__ double frdiv(double a, double d) { return xdiv(a, d); }
// real(conj(a)*a)
__device__ __forceinline__ double fnorm(cplx a) { return xadd(xmul(a.x, a.x), xmul(a.y, a.y)); }
__device__ __forceinline__ double fnorm(double a) { return xmul(a, a); }

template <typename T>
struct Field;
template <>
struct Field<double> {
  static constexpr int NCOMP = 1;
  __device__ __forceinline__ static double zero() { return 0.0; }
  __device__ __forceinline__ static double from(const double* a) { return a[0]; }
  // accumulate conj(a)*b into acc[0..NCOMP)
  __device__ __forceinline__ static void dot_acc(double* acc, double a, double b) { acc[0] += xmul(a, b); }
};
template <>
struct Field<cplx> {
  static constexpr int NCOMP = 2;
  __device__ __forceinline__ static cplx zero() { return mk(0.0, 0.0); }
  __device__ __forceinline__ static cplx from(const double* a) { return mk(a[0], a[1]); }
  __device__ __forceinline__ static void dot_acc(double* acc, cplx a, cplx b) {
    cplx t = fcmul(a, b);
    acc[0] += t.x;
    acc[1] += t.y;
  }
};

// Complex division as libgcc's __divdc3 performs it on its ordinary path (Smith's algorithm, no
// FMA): this is what `double / std::complex<double>` and `complex / complex` compile to in the
// reference (e.g. alpha = rsq/dot(p,Ap), generic_cg.cpp:326).  The overflow/underflow rescue
// branches of libgcc are not needed for solver scalars and are omitted.
__device__ __forceinline__ cplx cdiv(cplx n, cplx z) {
  const double a = n.x, b = n.y, c = z.x, d = z.y;
  if (fabs(c) < fabs(d)) {
    const double ratio = xdiv(c, d);
    const double denom = xadd(xmul(c, ratio), d);
    return mk(xdiv(xadd(xmul(a, ratio), b), denom), xdiv(xsub(xmul(b, ratio), a), denom));
  } else {
    const double ratio = xdiv(d, c);
    const double denom = xadd(xmul(d, ratio), c);
    return mk(xdiv(xadd(xmul(b, ratio), a), denom), xdiv(xsub(b, xmul(a, ratio)), denom));
  }
}
__device__ __forceinline__ double cdiv(double n, double z) { return xdiv(n, z); }

// ------------------------------------------------------------------ cache-hinted 16B / 8B access
__device__ __forceinline__ cplx ld_stream(const cplx* p) {  // read-once data: do not keep in L1
  cplx r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
// same as a plain load (L1 allocating), but as volatile asm: keeps its place in a hand-ordered batch of loads
__device__ __forceinline__ cplx ld_keep(const cplx* p) {
  cplx r;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_keep(const double* p) {
  double r;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ cplx ld_cached(const cplx* p) { return *p; }
__device__ __forceinline__ double ld_cached(const double* p) { return *p; }
__device__ __forceinline__ void st_stream(cplx* p, cplx v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// 1 or 2 adjacent complex sites per access: 16-byte or 32-byte (LDG.256 / STG.256 on sm_100a).
// The 2-site form needs a 32-byte aligned address (even site index on an even-width row).
template <int SPT>
__device__ __forceinline__ void ldv(const cplx* p, cplx (&v)[SPT]) {
  if (SPT == 2) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0].x), "=d"(v[0].y), "=d"(v[SPT - 1].x), "=d"(v[SPT - 1].y)
                 : "l"(p));
  } else {
    v[0] = *p;
  }
}
template <int SPT>
__device__ __forceinline__ void ldv_nc(const cplx* p, cplx (&v)[SPT]) {  // read-only data path
  if (SPT == 2) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0].x), "=d"(v[0].y), "=d"(v[SPT - 1].x), "=d"(v[SPT - 1].y)
                 : "l"(p));
  } else {
    v[0] = __ldg(p);
  }
}
template <int SPT>
__device__ __forceinline__ void stv(cplx* p, const cplx (&v)[SPT]) {
  if (SPT == 2) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0].x), "d"(v[0].y), "d"(v[SPT - 1].x),
                 "d"(v[SPT - 1].y)
                 : "memory");
  } else {
    *p = v[0];
  }
}

// cp.async (LDGSTS): 16-byte global -> shared copies that bypass the register file; a per-thread
// queue of commit groups gives deep prefetch without spending registers on staging.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ warp shuffles of wide values
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplx shfl_up_c(cplx v, int d) {
  return mk(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ cplx shfl_down_c(cplx v, int d) {
  return mk(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ double shfl_up_c(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_down_c(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// ------------------------------------------------------------------ deterministic reductions
constexpr int MAX_RED = 8;          // doubles reduced by one kernel (static kernels)
constexpr int MAX_PARTIAL_BLOCKS = 4096;

struct ReduceWs {
  double* partials;      // [nred][MAX_PARTIAL_BLOCKS]  device
  unsigned int* ticket;  // device counter, self-resetting
  double* result_dev;    // device copy of the final sums
  double* result_host;   // mapped pinned host copy (device-visible address), may be null
};

// Sum NRED per-thread values over the block (fixed tree).  Result valid in thread 0.
template <int NRED>
__device__ __forceinline__ void block_sum(double (&v)[NRED], double* smem /* NRED*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int r = 0; r < NRED; r++) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v[r] += shfl_xor_d(v[r], m);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < NRED; r++) smem[r * 32 + warp] = v[r];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r < NRED; r++) {
      double t = (lane < nwarp) ? smem[r * 32 + lane] : 0.0;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) t += shfl_xor_d(t, m);
      v[r] = t;
    }
  }
}

// Grid-wide deterministic sum.  Every thread of every block calls this with its per-thread
// values.  Returns true in ALL threads of the one block that arrived last, after which
// total[] (in thread 0 of that block only) holds the grid totals and have been stored to
// ws.result_dev / ws.result_host.  The caller may then run a scalar epilogue in thread 0.
template <int NRED>
__device__ __forceinline__ bool grid_sum(double (&v)[NRED], const ReduceWs& ws, double (&total)[NRED]) {
  __shared__ double s_red[NRED * 32];
  __shared__ bool s_last;
  block_sum<NRED>(v, s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < NRED; r++) ws.partials[r * MAX_PARTIAL_BLOCKS + blockIdx.x] = v[r];
    __threadfence();
    const unsigned int t = atomicInc(ws.ticket, gridDim.x - 1);  // wraps to 0 after the last block
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc[NRED];
#pragma unroll
  for (int r = 0; r < NRED; r++) {
    acc[r] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
      acc[r] += __ldcg(&ws.partials[r * MAX_PARTIAL_BLOCKS + b]);
  }
  __syncthreads();  // s_red reuse
  block_sum<NRED>(acc, s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < NRED; r++) {
      total[r] = acc[r];
      ws.result_dev[r] = acc[r];
      if (ws.result_host) ws.result_host[r] = acc[r];
    }
  }
  return true;
}

}  // namespace glb
