// blas1.cu -- device BLAS-1 for the Krylov shells: generic_vector.h:12-169 plus the open-coded
// axpy-type loops of every solver, fused so that each solver step is one pass over HBM.
//
// One streaming kernel template (ew_kernel) drives small functors.  Each thread moves 32 bytes
// per vector per step when alignment allows (LDG.256/STG.256), grid = a few blocks per SM with a
// grid-stride loop, reductions through glb::grid_sum (deterministic).  Element-wise arithmetic
// reproduces the reference's expressions exactly (no FMA contraction).
#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

int blas_grid(const glb_context* ctx, size_t n, int threads, int per_thread) {
  size_t want = (n + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
  size_t cap = (size_t)ctx->sm_count * 8;
  if (cap > (size_t)MAX_PARTIAL_BLOCKS) cap = MAX_PARTIAL_BLOCKS;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

template <typename T, int W>
struct alignas(sizeof(T) * W) Pack {
  T v[W];
};

template <int NV>
struct VecPtrs {
  void* p[NV];
};

// F::NV vectors; bit k of F::RD / F::WR says vector k is read / written; F::NRED reductions.
template <typename T, typename F, int W>
__global__ void __launch_bounds__(256) ew_kernel(F f, VecPtrs<F::NV> ptrs, size_t n, ReduceWs red) {
  constexpr int NRED = F::NRED > 0 ? F::NRED : 1;
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x * W;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * W; i < n; i += stride) {
    Pack<T, W> v[F::NV];
#pragma unroll
    for (int k = 0; k < F::NV; k++)
      if ((F::RD >> k) & 1) v[k] = *reinterpret_cast<const Pack<T, W>*>((const T*)ptrs.p[k] + i);
#pragma unroll
    for (int w = 0; w < W; w++) {
      T e[F::NV];
#pragma unroll
      for (int k = 0; k < F::NV; k++) e[k] = v[k].v[w];
      f.elem(e, acc);
#pragma unroll
      for (int k = 0; k < F::NV; k++) v[k].v[w] = e[k];
    }
#pragma unroll
    for (int k = 0; k < F::NV; k++)
      if ((F::WR >> k) & 1) *reinterpret_cast<Pack<T, W>*>((T*)ptrs.p[k] + i) = v[k];
  }
  if (F::NRED > 0) {
    double total[NRED];
    grid_sum<NRED>(acc, red, total);
  }
}

template <typename T, typename F>
static int run_ew(glb_context* ctx, const F& f, const VecPtrs<F::NV>& ptrs, size_t n, bool to_host) {
  constexpr int WMAX = 32 / sizeof(T);
  bool wide = (n % WMAX == 0);
  for (int k = 0; k < F::NV; k++) wide = wide && (((uintptr_t)ptrs.p[k] & 31u) == 0);
  ReduceWs red = ctx->red;
  if (!to_host) red.result_host = nullptr;
  ProfScope prof(ctx, PROF_EW, (double)n * sizeof(T) * (__builtin_popcount(F::RD) + __builtin_popcount(F::WR)));
  if (wide) {
    const int grid = blas_grid(ctx, n / WMAX, 256, 2);
    ew_kernel<T, F, WMAX><<<grid, 256, 0, ctx->stream>>>(f, ptrs, n, red);
  } else {
    const int grid = blas_grid(ctx, n, 256, 4);
    ew_kernel<T, F, 1><<<grid, 256, 0, ctx->stream>>>(f, ptrs, n, red);
  }
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

// wait for the stream and fetch k reduction results from the mapped host buffer
static int fetch_results(glb_context* ctx, double* out, int k) {
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < k; i++) out[i] = ctx->result_host_ptr[i];
  if (ctx->nranks > 1) return allreduce_sum(ctx, out, k);
  return GLB_OK;
}

// ------------------------------------------------------------------------- functors
template <typename T>
struct FDot {  // <x,y>
  static constexpr int NV = 2, RD = 3, WR = 0, NRED = Field<T>::NCOMP;
  __device__ void elem(T (&e)[NV], double* acc) const { Field<T>::dot_acc(acc, e[0], e[1]); }
};
template <typename T>
struct FNorm {  // |x|^2
  static constexpr int NV = 1, RD = 1, WR = 0, NRED = 1;
  __device__ void elem(T (&e)[NV], double* acc) const { acc[0] += fnorm(e[0]); }
};
template <typename T>
struct FDiffNorm {  // |x-y|^2, generic_vector.h:148,160
  static constexpr int NV = 2, RD = 3, WR = 0, NRED = 1;
  __device__ void elem(T (&e)[NV], double* acc) const { acc[0] += fnorm(fsub(e[0], e[1])); }
};
template <typename T>
struct FDotNorm {  // <x,y>, |x|^2
  static constexpr int NV = 2, RD = 3, WR = 0, NRED = Field<T>::NCOMP + 1;
  __device__ void elem(T (&e)[NV], double* acc) const {
    Field<T>::dot_acc(acc, e[0], e[1]);
    acc[Field<T>::NCOMP] += fnorm(e[0]);
  }
};
template <typename T>
struct FSub {  // out = a - b
  static constexpr int NV = 3, RD = 3, WR = 4, NRED = 0;
  __device__ void elem(T (&e)[NV], double*) const { e[2] = fsub(e[0], e[1]); }
};
template <typename T>
struct FAdd {  // out = a + b
  static constexpr int NV = 3, RD = 3, WR = 4, NRED = 0;
  __device__ void elem(T (&e)[NV], double*) const { e[2] = fadd(e[0], e[1]); }
};
template <typename T>
struct FAxpy {  // y = y + a*x        vectors: x, y
  static constexpr int NV = 2, RD = 3, WR = 2, NRED = 0;
  T a;
  __device__ void elem(T (&e)[NV], double*) const { e[1] = fadd(e[1], fmul(a, e[0])); }
};
template <typename T>
struct FXpay {  // y = x + a*y        vectors: x, y
  static constexpr int NV = 2, RD = 3, WR = 2, NRED = 0;
  T a;
  __device__ void elem(T (&e)[NV], double*) const { e[1] = fadd(e[0], fmul(a, e[1])); }
};
template <typename T>
struct FAxpyz {  // z = y + a*x       vectors: x, y, z
  static constexpr int NV = 3, RD = 3, WR = 4, NRED = 0;
  T a;
  __device__ void elem(T (&e)[NV], double*) const { e[2] = fadd(e[1], fmul(a, e[0])); }
};
template <typename T>
struct FRdiv {  // out = x / d        vectors: x, out
  static constexpr int NV = 2, RD = 1, WR = 2, NRED = 0;
  double d;
  __device__ void elem(T (&e)[NV], double*) const { e[1] = frdiv(e[0], d); }
};
template <typename T>
struct FAxpyNorm {  // y = y + a*x ; |y|^2
  static constexpr int NV = 2, RD = 3, WR = 2, NRED = 1;
  T a;
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[1] = fadd(e[1], fmul(a, e[0]));
    acc[0] += fnorm(e[1]);
  }
};
template <typename T>
struct FUpdateXR {  // x = x + a*p ; r = r + b*q ; |r|^2      vectors: p, x, q, r
  static constexpr int NV = 4, RD = 15, WR = 10, NRED = 1;
  T a, b;
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[1] = fadd(e[1], fmul(a, e[0]));
    e[3] = fadd(e[3], fmul(b, e[2]));
    acc[0] += fnorm(e[3]);
  }
};
template <typename T>
struct FUpdatePAp {  // p = r + beta*p ; Ap = Ar + beta*Ap ; |Ap|^2   vectors: r, Ar, p, Ap
  static constexpr int NV = 4, RD = 15, WR = 12, NRED = 1;
  T beta;
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[2] = fadd(e[0], fmul(beta, e[2]));
    e[3] = fadd(e[1], fmul(beta, e[3]));
    acc[0] += fnorm(e[3]);
  }
};
template <typename T>
struct FBicgUpdate {  // vectors: p, s, As, r0, x, r       (generic_bicgstab.cpp:274-295)
  static constexpr int NV = 6, RD = 31, WR = 48, NRED = 1 + Field<T>::NCOMP;
  T alpha, omega;
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[4] = fadd(fadd(e[4], fmul(alpha, e[0])), fmul(omega, e[1]));  // x = x + alpha*p + omega*s
    e[5] = fsub(e[1], fmul(omega, e[2]));                            // r = s - omega*As
    acc[0] += fnorm(e[5]);
    Field<T>::dot_acc(acc + 1, e[3], e[5]);                          // <r0, r>
  }
};
template <typename T>
struct FBicgP {  // p = r + beta*(p - omega*Ap)     vectors: r, Ap, p    (generic_bicgstab.cpp:300-303)
  static constexpr int NV = 3, RD = 7, WR = 4, NRED = 0;
  T beta, omega;
  __device__ void elem(T (&e)[NV], double*) const { e[2] = fadd(e[0], fmul(beta, fsub(e[2], fmul(omega, e[1])))); }
};
template <typename T>
struct FRscale {  // out = s*x, real s       vectors: x, out   (generic_vector.h normalize)
  static constexpr int NV = 2, RD = 1, WR = 2, NRED = 0;
  double s;
  __device__ void elem(T (&e)[NV], double*) const { e[1] = fscale(s, e[0]); }
};
template <typename T>
struct FConj {  // out = conj(x)   (a copy for real fields)     vectors: x, out    (generic_vector.h conj<>)
  static constexpr int NV = 2, RD = 1, WR = 2, NRED = 0;
  __device__ static double cj(double a) { return a; }
  __device__ static cplx cj(cplx a) { return mk(a.x, -a.y); }
  __device__ void elem(T (&e)[NV], double*) const { e[1] = cj(e[0]); }
};
template <typename T>
struct FBicgMS {  // s_n = c0 r + c1 (s_n - c2 (c3 w - c4 r_prev))   vectors: r, w, r_prev, s_n  (generic_bicgstab_m.cpp:705)
  static constexpr int NV = 4, RD = 15, WR = 8, NRED = 0;
  T c0, c1, c2, c3, c4;
  __device__ void elem(T (&e)[NV], double*) const {
    const T inner = fsub(fmul(c3, e[1]), fmul(c4, e[2]));
    e[3] = fadd(fmul(c0, e[0]), fmul(c1, fsub(e[3], fmul(c2, inner))));
  }
};

template <typename T>
static T coef(const double a[2]);
template <>
double coef<double>(const double a[2]) {
  return a[0];
}
template <>
cplx coef<cplx>(const double a[2]) {
  return make_double2(a[0], a[1]);
}

// ------------------------------------------------------------------------- multi-vector kernels
constexpr int MAXK = 16;
template <typename T>
struct MultiArgs {
  const T* X[MAXK];
  T* Y[MAXK];
  T c0[MAXK];
  T c1[MAXK];
  int k;
};

// out[NC*j ..] = <X[j], y> for up to MAXK stored vectors in ONE pass over y (the Gram-Schmidt sweeps of GCR / VPGCR /
// the flexible solvers, generic_gcr.cpp:284-292).
// Shape: the 8 warps of a block sweep the SAME elements, warp w owning the vectors [w*VPW, (w+1)*VPW) -- a thread
// carries 2*VPW accumulators instead of 32, so the registers go to loads in flight (U elements x (VPW + 1) 16-byte
// loads per thread) and not to sums: one thread summing all 16 vectors needs 102 registers, keeps ~4 loads in flight
// and reaches 35 % of the HBM peak (profiles/r02/config5_profile_before_gs_kernels.jsonl).  y is fetched from DRAM once
// per block (the other warps hit L1).  Warp sums -> per-block partials -> the block that arrives last adds the
// partials of every vector in block order: fixed summation order, run-to-run reproducible.
constexpr int MD_U = 4;  // elements per lane and chunk
template <typename T, int VPW>
__global__ void __launch_bounds__(256, 3) multi_dot_kernel(MultiArgs<T> a, const T* __restrict__ y, size_t n, ReduceWs red) {
  constexpr int NC = Field<T>::NCOMP;
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = warp * VPW;
  double acc[VPW * NC];
#pragma unroll
  for (int i = 0; i < VPW * NC; i++) acc[i] = 0.0;
  if (j0 < a.k) {
    const T* xp[VPW];
#pragma unroll
    for (int v = 0; v < VPW; v++) xp[v] = a.X[(j0 + v < a.k) ? j0 + v : j0];
    const size_t chunk = (size_t)32 * MD_U;
    for (size_t c0 = (size_t)blockIdx.x * chunk; c0 < n; c0 += (size_t)gridDim.x * chunk) {
      T yv[MD_U], xv[VPW][MD_U];
      if (c0 + chunk <= n) {
        // full chunk: every load is issued (volatile asm keeps them ahead of the arithmetic) before the first use
#pragma unroll
        for (int u = 0; u < MD_U; u++) {
          const size_t i = c0 + (size_t)u * 32 + lane;
#pragma unroll
          for (int v = 0; v < VPW; v++) xv[v][u] = ld_stream(xp[v] + i);
          yv[u] = ld_keep(y + i);
        }
      } else {
#pragma unroll
        for (int u = 0; u < MD_U; u++) {
          const size_t i = c0 + (size_t)u * 32 + lane;
          const bool in = i < n;
          yv[u] = in ? y[i] : Field<T>::zero();
#pragma unroll
          for (int v = 0; v < VPW; v++) xv[v][u] = in ? xp[v][i] : Field<T>::zero();
        }
      }
#pragma unroll
      for (int u = 0; u < MD_U; u++)
#pragma unroll
        for (int v = 0; v < VPW; v++) Field<T>::dot_acc(acc + v * NC, xv[v][u], yv[u]);
    }
  }
#pragma unroll
  for (int r = 0; r < VPW * NC; r++) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc[r] += shfl_xor_d(acc[r], m);
  }
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < VPW; v++) {
      const bool live = j0 + v < a.k;
#pragma unroll
      for (int c = 0; c < NC; c++)
        red.partials[(size_t)((j0 + v) * NC + c) * MAX_PARTIAL_BLOCKS + blockIdx.x] = live ? acc[v * NC + c] : 0.0;
    }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(red.ticket, gridDim.x - 1);  // wraps to 0 after the last block
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // final pass: warp w finishes the sums of its own vectors
#pragma unroll
  for (int r = 0; r < VPW * NC; r++) {
    const int slot = j0 * NC + r;
    double t = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&red.partials[(size_t)slot * MAX_PARTIAL_BLOCKS + b]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) t += shfl_xor_d(t, m);
    if (lane == 0) {
      red.result_dev[slot] = t;
      if (red.result_host) red.result_host[slot] = t;
    }
  }
}

// out = init + c0[0]*X[0] + c0[1]*X[1] + ...   (sequential accumulation, generic_gcr.cpp:284-292).  The loop over the
// vectors is unrolled over all MAXK slots so that every load of an element is in flight before the first multiply;
// the sum still runs in the order j = 0, 1, ... (bit-identical to the serial loop).
template <typename T>
__global__ void __launch_bounds__(256) lincomb_kernel(MultiArgs<T> a, const T* __restrict__ init, T* __restrict__ out,
                                                      size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    T xv[MAXK];
#pragma unroll
    for (int j = 0; j < MAXK; j++)
      if (j < a.k) xv[j] = ld_stream(a.X[j] + i);
    T v = init ? init[i] : Field<T>::zero();
#pragma unroll
    for (int j = 0; j < MAXK; j++)
      if (j < a.k) v = fadd(v, fmul(a.c0[j], xv[j]));
    out[i] = v;
  }
}

// x[s] = x[s] - beta_s[s]*p_s[s]           (generic_cg_m.cpp:414-417; c0 carries beta_s)
template <typename T>
__global__ void __launch_bounds__(256) cgm_x_kernel(MultiArgs<T> a, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    for (int j = 0; j < a.k; j++) a.Y[j][i] = fsub(a.Y[j][i], fmul(a.c0[j], a.X[j][i]));
  }
}
// p_s[s] = zeta[s]*r + alpha_s[s]*p_s[s]   (generic_cg_m.cpp:501-512; c0 = zeta, c1 = alpha_s), r read once
template <typename T>
__global__ void __launch_bounds__(256) cgm_p_kernel(MultiArgs<T> a, const T* r, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const T rv = r[i];
    for (int j = 0; j < a.k; j++) a.Y[j][i] = fadd(fmul(a.c0[j], rv), fmul(a.c1[j], a.Y[j][i]));
  }
}

// ------------------------------------------------------------------------- device-resident CG kernels
// x = x + alpha p ; r = r - alpha Ap ; rsqNew = |r|^2, with alpha = rsq/<p,Ap> taken from the CG
// state, and the stopping test of generic_cg.cpp:339 evaluated by the last block.
template <typename T, int W>
__global__ void __launch_bounds__(256)
cg_update_kernel(CgState* st, double* hist, const T* p, T* x, const T* Ap, T* r, size_t n, ReduceWs red, int defer,
                 P2PRed pr) {
  if (st->done) return;
  T alpha;
  {
    const double rsq = st->rsq_old;
    if (Field<T>::NCOMP == 2) {
      const cplx a = cdiv(mk(rsq, 0.0), mk(st->pAp_re, st->pAp_im));
      alpha = *reinterpret_cast<const T*>(&a);
    } else {
      const double a = xdiv(rsq, st->pAp_re);
      alpha = *reinterpret_cast<const T*>(&a);
    }
  }
  const T nalpha = fneg(alpha);
  double acc[1] = {0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x * W;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * W; i < n; i += stride) {
    Pack<T, W> vp = *reinterpret_cast<const Pack<T, W>*>(p + i);
    Pack<T, W> vx = *reinterpret_cast<const Pack<T, W>*>(x + i);
    Pack<T, W> vq = *reinterpret_cast<const Pack<T, W>*>(Ap + i);
    Pack<T, W> vr = *reinterpret_cast<const Pack<T, W>*>(r + i);
#pragma unroll
    for (int w = 0; w < W; w++) {
      vx.v[w] = fadd(vx.v[w], fmul(alpha, vp.v[w]));   // phi = phi + alpha*p
      vr.v[w] = fadd(vr.v[w], fmul(nalpha, vq.v[w]));  // r = r - alpha*Ap
      acc[0] += fnorm(vr.v[w]);
    }
    *reinterpret_cast<Pack<T, W>*>(x + i) = vx;
    *reinterpret_cast<Pack<T, W>*>(r + i) = vr;
  }
  double total[1];
  if (!grid_sum<1>(acc, red, total)) return;  // only the block that arrived last goes on
  if (defer == 2) p2p_allreduce_block(pr, total, 1);  // slabs over peer memory: finish the sum here
  if (threadIdx.x == 0) {
    if (defer == 1) {  // slab run over NCCL: the sum over ranks and the recurrence step follow on the stream
      st->partial[0] = total[0];
      return;
    }
    const double rsq_new = total[0];
    st->rsq_new = rsq_new;
    const int k = st->iter;  // 0-based iteration index of the reference loop
    st->iter = k + 1;
    if (hist != nullptr && k < st->hist_cap) hist[k] = rsq_new;
    const bool conv = sqrt(rsq_new) < st->eps * st->bnorm;
    const bool last = (k == st->max_iter - 1);
    if (conv || last) {
      st->done = 1;
      st->hit_max = (k == st->max_iter - 1) ? 1 : 0;  // generic_cg.cpp:356 tests k alone
    }
  }
}

// slab runs: fold the rank-summed |r|^2 into the recurrence (what the last block does on one rank)
__global__ void cg_post_update_kernel(CgState* st, double* hist) {
  if (st->done) return;
  const double rsq_new = st->partial[0];
  st->rsq_new = rsq_new;
  const int k = st->iter;
  st->iter = k + 1;
  if (hist != nullptr && k < st->hist_cap) hist[k] = rsq_new;
  const bool conv = sqrt(rsq_new) < st->eps * st->bnorm;
  const bool last = (k == st->max_iter - 1);
  if (conv || last) {
    st->done = 1;
    st->hit_max = last ? 1 : 0;
  }
}
// slab runs: fold the rank-summed <p,Ap> into the recurrence
__global__ void cg_post_apply_kernel(CgState* st) {
  if (st->done) return;
  st->pAp_re = st->partial[1];
  st->pAp_im = st->partial[2];
  st->rsq_old = st->rsq_new;
}
// slab runs: the new direction p = r + beta p_old on the `nrows` lowest and highest rows of the slab,
// written to the send buffers that go to the neighbouring ranks' ghost rows
template <typename T>
__global__ void cg_boundary_kernel(const CgState* st, const T* r, const T* pold, T* send_lo, T* send_hi,
                                   size_t row_elems, int nrows, size_t local_elems) {
  if (st->done) return;
  const double b = xdiv(st->rsq_new, st->rsq_old);
  const size_t n = row_elems * nrows;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
    const bool hi = i >= n;
    const size_t j = hi ? i - n : i;
    const size_t src = hi ? local_elems - n + j : j;
    (hi ? send_hi : send_lo)[j] = fadd(r[src], fscale(b, pold[src]));
  }
}

// same, but the rows go straight into the neighbours' ghost rows over NVLink and the last block raises
// their flags (compute + halo push in one kernel)
template <typename T>
__global__ void cg_boundary_push_kernel(const CgState* st, const T* r, const T* pold, T* dst_down_hi, T* dst_up_lo,
                                        size_t row_elems, int nrows, size_t local_elems,
                                        unsigned long long* flag_down_hi, unsigned long long* flag_up_lo,
                                        unsigned long long seq, unsigned int* ticket) {
  // NB: no early exit on st->done -- the neighbours' kernels wait for these flags
  const bool done = st->done != 0;
  const double b = done ? 0.0 : xdiv(st->rsq_new, st->rsq_old);
  const size_t n = row_elems * nrows;
  if (!done) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
      const bool hi = i >= n;
      const size_t j = hi ? i - n : i;
      const size_t src = hi ? local_elems - n + j : j;
      (hi ? dst_up_lo : dst_down_hi)[j] = fadd(r[src], fscale(b, pold[src]));
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    if (t == gridDim.x - 1) {
      __threadfence_system();
      st_release_sys(flag_down_hi, seq);
      st_release_sys(flag_up_lo, seq);
    }
  }
}

// p = r + beta p with beta = rsqNew/rsq from the CG state (for operators without the fused input)
template <typename T>
__global__ void __launch_bounds__(256) cg_xpay_kernel(const CgState* st, const T* r, T* p, size_t n) {
  if (st->done) return;
  const double b = xdiv(st->rsq_new, st->rsq_old);
  T beta;
  if (Field<T>::NCOMP == 2) {
    const cplx bc = mk(b, 0.0);
    beta = *reinterpret_cast<const T*>(&bc);
  } else {
    beta = *reinterpret_cast<const T*>(&b);
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = fadd(r[i], fmul(beta, p[i]));
}

int launch_cg_post_update(glb_context* ctx, void* st, double* hist) {
  cg_post_update_kernel<<<1, 1, 0, ctx->stream>>>((CgState*)st, hist);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}
int launch_cg_post_apply(glb_context* ctx, void* st) {
  cg_post_apply_kernel<<<1, 1, 0, ctx->stream>>>((CgState*)st);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}
int launch_cg_boundary(glb_context* ctx, const void* st, const void* r, const void* pold, void* send_lo, void* send_hi,
                       size_t row_elems, int nrows, size_t local_elems) {
  const int grid = blas_grid(ctx, 2 * row_elems * nrows, 256, 1);
  cg_boundary_kernel<cplx><<<grid, 256, 0, ctx->stream>>>((const CgState*)st, (const cplx*)r, (const cplx*)pold,
                                                          (cplx*)send_lo, (cplx*)send_hi, row_elems, nrows, local_elems);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_cg_boundary_push(glb_context* ctx, const void* st, const void* r, const void* pold, size_t row_elems, int nrows,
                            size_t local_elems, const HaloTargets& t) {
  const int grid = blas_grid(ctx, 2 * row_elems * nrows, 256, 1);
  cg_boundary_push_kernel<cplx><<<grid, 256, 0, ctx->stream>>>(
      (const CgState*)st, (const cplx*)r, (const cplx*)pold, (cplx*)t.dst_down_hi, (cplx*)t.dst_up_lo, row_elems, nrows,
      local_elems, t.flag_down_hi, t.flag_up_lo, t.wait.seq, t.ticket);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_cg_update(glb_context* ctx, int dtype, void* st, double* hist, const void* p, void* x, const void* Ap,
                     void* r, size_t n, int defer) {
  P2PRed pr{};
  if (defer == 2) pr = comm_p2p_red(ctx);
  ReduceWs red = ctx->red;
  red.result_host = nullptr;
  ProfScope prof(ctx, PROF_CG_UPDATE);
  if (dtype == GLB_COMPLEX) {
    const bool wide = (n % 2 == 0) && ((((uintptr_t)p | (uintptr_t)x | (uintptr_t)Ap | (uintptr_t)r) & 31u) == 0);
    if (wide) {
      const int grid = blas_grid(ctx, n / 2, 256, 2);
      cg_update_kernel<cplx, 2><<<grid, 256, 0, ctx->stream>>>((CgState*)st, hist, (const cplx*)p, (cplx*)x,
                                                               (const cplx*)Ap, (cplx*)r, n, red, defer, pr);
    } else {
      const int grid = blas_grid(ctx, n, 256, 4);
      cg_update_kernel<cplx, 1><<<grid, 256, 0, ctx->stream>>>((CgState*)st, hist, (const cplx*)p, (cplx*)x,
                                                               (const cplx*)Ap, (cplx*)r, n, red, defer, pr);
    }
  } else {
    const int grid = blas_grid(ctx, n, 256, 4);
    cg_update_kernel<double, 1><<<grid, 256, 0, ctx->stream>>>((CgState*)st, hist, (const double*)p, (double*)x,
                                                               (const double*)Ap, (double*)r, n, red, defer, pr);
  }
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_cg_xpay(glb_context* ctx, int dtype, const void* st, const void* r, void* p, size_t n) {
  const int grid = blas_grid(ctx, n, 256, 4);
  if (dtype == GLB_COMPLEX)
    cg_xpay_kernel<cplx><<<grid, 256, 0, ctx->stream>>>((const CgState*)st, (const cplx*)r, (cplx*)p, n);
  else
    cg_xpay_kernel<double><<<grid, 256, 0, ctx->stream>>>((const CgState*)st, (const double*)r, (double*)p, n);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

}  // namespace glb

// =============================================================================== C ABI
using namespace glb;

// run fn(T{}) with T = cplx or double
template <typename Fn>
static int dispatch(int dtype, Fn&& fn) {
  if (dtype == GLB_COMPLEX) return fn(cplx{});
  if (dtype == GLB_REAL) return fn(double{});
  return fail(GLB_ERR_ARG, "bad dtype");
}

template <template <typename> class F, int NVEC, typename Init>
static int ew_call(glb_context* ctx, int dtype, size_t n, bool to_host, const VecPtrs<NVEC>& ptrs, Init&& init) {
  return dispatch(dtype, [&](auto tag) {
    typedef decltype(tag) T;
    F<T> f;
    init(f);
    return run_ew<T, F<T>>(ctx, f, ptrs, n, to_host);
  });
}
struct NoInit {
  template <typename F>
  void operator()(F&) const {}
};

extern "C" {

int glb_dot(glb_context* ctx, int dtype, size_t n, const void* x, const void* y, double out[2]) {
  VecPtrs<2> p{{(void*)x, (void*)y}};
  int rc = ew_call<FDot, 2>(ctx, dtype, n, true, p, NoInit());
  if (rc) return rc;
  out[1] = 0.0;
  return fetch_results(ctx, out, dtype == GLB_COMPLEX ? 2 : 1);
}

int glb_norm2sq(glb_context* ctx, int dtype, size_t n, const void* x, double* out) {
  VecPtrs<1> p{{(void*)x}};
  int rc = ew_call<FNorm, 1>(ctx, dtype, n, true, p, NoInit());
  if (rc) return rc;
  return fetch_results(ctx, out, 1);
}

int glb_diffnorm2sq(glb_context* ctx, int dtype, size_t n, const void* x, const void* y, double* out) {
  VecPtrs<2> p{{(void*)x, (void*)y}};
  int rc = ew_call<FDiffNorm, 2>(ctx, dtype, n, true, p, NoInit());
  if (rc) return rc;
  return fetch_results(ctx, out, 1);
}

int glb_dot_norm(glb_context* ctx, int dtype, size_t n, const void* x, const void* y, double out[3]) {
  VecPtrs<2> p{{(void*)x, (void*)y}};
  int rc = ew_call<FDotNorm, 2>(ctx, dtype, n, true, p, NoInit());
  if (rc) return rc;
  double tmp[3] = {0, 0, 0};
  rc = fetch_results(ctx, tmp, dtype == GLB_COMPLEX ? 3 : 2);
  if (dtype == GLB_COMPLEX) {
    out[0] = tmp[0], out[1] = tmp[1], out[2] = tmp[2];
  } else {
    out[0] = tmp[0], out[1] = 0.0, out[2] = tmp[1];
  }
  return rc;
}

int glb_multi_dot(glb_context* ctx, int dtype, size_t n, int k, const void* const* X, const void* y, double* out) {
  if (k < 0) return fail(GLB_ERR_ARG, "glb_multi_dot: k < 0");
  for (int base = 0; base < k; base += MAXK) {
    const int kk = (k - base < MAXK) ? k - base : MAXK;
    int rc = dispatch(dtype, [&](auto tag) {
      typedef decltype(tag) T;
      MultiArgs<T> a{};
      a.k = kk;
      for (int j = 0; j < kk; j++) a.X[j] = (const T*)X[base + j];
      // one 128-element chunk per block and step; 3 resident blocks per SM keep ~150 KB of loads in flight per SM
      size_t want = (n + (size_t)32 * MD_U - 1) / ((size_t)32 * MD_U);
      const size_t cap = (size_t)ctx->sm_count * 3;
      const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
      ProfScope prof(ctx, PROF_MULTI_DOT, (double)n * sizeof(T) * (kk + 1));
      if (kk <= 8)
        multi_dot_kernel<T, 1><<<grid, 256, 0, ctx->stream>>>(a, (const T*)y, n, ctx->red);
      else
        multi_dot_kernel<T, 2><<<grid, 256, 0, ctx->stream>>>(a, (const T*)y, n, ctx->red);
      GLB_LAUNCH_CHECK();
      return GLB_OK;
    });
    if (rc) return rc;
    const int nc = dtype == GLB_COMPLEX ? 2 : 1;
    double tmp[MAXK * 2];
    rc = fetch_results(ctx, tmp, MAXK * nc);
    if (rc) return rc;
    for (int j = 0; j < kk; j++) {
      out[2 * (base + j)] = tmp[j * nc];
      out[2 * (base + j) + 1] = (nc == 2) ? tmp[j * nc + 1] : 0.0;
    }
  }
  return GLB_OK;
}

int glb_sub(glb_context* ctx, int dtype, size_t n, const void* a, const void* b, void* out) {
  VecPtrs<3> p{{(void*)a, (void*)b, out}};
  return ew_call<FSub, 3>(ctx, dtype, n, false, p, NoInit());
}
int glb_add(glb_context* ctx, int dtype, size_t n, const void* a, const void* b, void* out) {
  VecPtrs<3> p{{(void*)a, (void*)b, out}};
  return ew_call<FAdd, 3>(ctx, dtype, n, false, p, NoInit());
}
int glb_axpy(glb_context* ctx, int dtype, size_t n, const double a[2], const void* x, void* y) {
  VecPtrs<2> p{{(void*)x, y}};
  return ew_call<FAxpy, 2>(ctx, dtype, n, false, p, [&](auto& f) { f.a = coef<decltype(f.a)>(a); });
}
int glb_xpay(glb_context* ctx, int dtype, size_t n, const void* x, const double a[2], void* y) {
  VecPtrs<2> p{{(void*)x, y}};
  return ew_call<FXpay, 2>(ctx, dtype, n, false, p, [&](auto& f) { f.a = coef<decltype(f.a)>(a); });
}
int glb_axpyz(glb_context* ctx, int dtype, size_t n, const double a[2], const void* x, const void* y, void* z) {
  VecPtrs<3> p{{(void*)x, (void*)y, z}};
  return ew_call<FAxpyz, 3>(ctx, dtype, n, false, p, [&](auto& f) { f.a = coef<decltype(f.a)>(a); });
}
int glb_rdiv(glb_context* ctx, int dtype, size_t n, const void* x, double d, void* out) {
  VecPtrs<2> p{{(void*)x, out}};
  return ew_call<FRdiv, 2>(ctx, dtype, n, false, p, [&](auto& f) { f.d = d; });
}
int glb_bicgstab_pupdate(glb_context* ctx, int dtype, size_t n, const void* r, const double beta[2],
                         const double omega[2], const void* Ap, void* p) {
  VecPtrs<3> v{{(void*)r, (void*)Ap, p}};
  return ew_call<FBicgP, 3>(ctx, dtype, n, false, v, [&](auto& f) {
    f.beta = coef<decltype(f.beta)>(beta);
    f.omega = coef<decltype(f.omega)>(omega);
  });
}

int glb_rscale(glb_context* ctx, int dtype, size_t n, const void* x, double sc, void* out) {
  VecPtrs<2> v{{(void*)x, out}};
  return ew_call<FRscale, 2>(ctx, dtype, n, false, v, [&](auto& f) { f.s = sc; });
}

int glb_conj(glb_context* ctx, int dtype, size_t n, const void* x, void* out) {
  VecPtrs<2> v{{(void*)x, out}};
  return ew_call<FConj, 2>(ctx, dtype, n, false, v, NoInit());
}

int glb_bicgstabm_update_s(glb_context* ctx, int dtype, size_t n, const double c[10], const void* r, const void* w,
                           const void* r_prev, void* s_n) {
  VecPtrs<4> v{{(void*)r, (void*)w, (void*)r_prev, s_n}};
  return ew_call<FBicgMS, 4>(ctx, dtype, n, false, v, [&](auto& f) {
    f.c0 = coef<decltype(f.c0)>(c);
    f.c1 = coef<decltype(f.c1)>(c + 2);
    f.c2 = coef<decltype(f.c2)>(c + 4);
    f.c3 = coef<decltype(f.c3)>(c + 6);
    f.c4 = coef<decltype(f.c4)>(c + 8);
  });
}

int glb_axpy_norm(glb_context* ctx, int dtype, size_t n, const double a[2], const void* x, void* y, double* nrm) {
  VecPtrs<2> p{{(void*)x, y}};
  int rc = ew_call<FAxpyNorm, 2>(ctx, dtype, n, true, p, [&](auto& f) { f.a = coef<decltype(f.a)>(a); });
  if (rc) return rc;
  return fetch_results(ctx, nrm, 1);
}

int glb_update_xr_norm(glb_context* ctx, int dtype, size_t n, const double a[2], const void* p, void* x,
                       const double b[2], const void* q, void* r, double* rsq) {
  VecPtrs<4> v{{(void*)p, x, (void*)q, r}};
  int rc = ew_call<FUpdateXR, 4>(ctx, dtype, n, true, v, [&](auto& f) {
    f.a = coef<decltype(f.a)>(a);
    f.b = coef<decltype(f.b)>(b);
  });
  if (rc) return rc;
  return fetch_results(ctx, rsq, 1);
}

int glb_update_p_ap_norm(glb_context* ctx, int dtype, size_t n, const void* r, const void* Ar, const double beta[2],
                         void* p, void* Ap, double* apsq) {
  VecPtrs<4> v{{(void*)r, (void*)Ar, p, Ap}};
  int rc = ew_call<FUpdatePAp, 4>(ctx, dtype, n, true, v, [&](auto& f) { f.beta = coef<decltype(f.beta)>(beta); });
  if (rc) return rc;
  return fetch_results(ctx, apsq, 1);
}

int glb_bicgstab_update(glb_context* ctx, int dtype, size_t n, const double alpha[2], const void* p,
                        const double omega[2], const void* s, const void* As, const void* r0, void* x, void* r,
                        double out[3]) {
  VecPtrs<6> v{{(void*)p, (void*)s, (void*)As, (void*)r0, x, r}};
  int rc = ew_call<FBicgUpdate, 6>(ctx, dtype, n, true, v, [&](auto& f) {
    f.alpha = coef<decltype(f.alpha)>(alpha);
    f.omega = coef<decltype(f.omega)>(omega);
  });
  if (rc) return rc;
  out[2] = 0.0;
  return fetch_results(ctx, out, dtype == GLB_COMPLEX ? 3 : 2);
}

int glb_lincomb(glb_context* ctx, int dtype, size_t n, int k, const double* coefs, const void* const* X,
                const void* init, void* out) {
  if (k < 0) return fail(GLB_ERR_ARG, "glb_lincomb: k < 0");
  if (k == 0) {
    if (init) return (init == out) ? GLB_OK : glb_vec_copy(ctx, dtype, n, out, init);
    return glb_vec_zero(ctx, dtype, n, out);
  }
  const void* cur_init = init;
  for (int base = 0; base < k; base += MAXK) {
    const int kk = (k - base < MAXK) ? k - base : MAXK;
    int rc = dispatch(dtype, [&](auto tag) {
      typedef decltype(tag) T;
      MultiArgs<T> a{};
      a.k = kk;
      for (int j = 0; j < kk; j++) {
        a.X[j] = (const T*)X[base + j];
        a.c0[j] = coef<T>(coefs + 2 * (base + j));
      }
      const int grid = blas_grid(ctx, n, 256, 2);
      ProfScope prof(ctx, PROF_LINCOMB, (double)n * sizeof(T) * (kk + (cur_init ? 1 : 0) + 1));
      lincomb_kernel<T><<<grid, 256, 0, ctx->stream>>>(a, (const T*)cur_init, (T*)out, n);
      GLB_LAUNCH_CHECK();
      return GLB_OK;
    });
    if (rc) return rc;
    cur_init = out;  // later chunks keep accumulating onto the running sum
  }
  return GLB_OK;
}

int glb_cgm_update_x(glb_context* ctx, int dtype, size_t n, int ns, const double* beta_s, const void* const* p_s,
                     void* const* x) {
  for (int base = 0; base < ns; base += MAXK) {
    const int kk = (ns - base < MAXK) ? ns - base : MAXK;
    int rc = dispatch(dtype, [&](auto tag) {
      typedef decltype(tag) T;
      MultiArgs<T> a{};
      a.k = kk;
      for (int j = 0; j < kk; j++) {
        a.X[j] = (const T*)p_s[base + j];
        a.Y[j] = (T*)x[base + j];
        a.c0[j] = coef<T>(beta_s + 2 * (base + j));
      }
      const int grid = blas_grid(ctx, n, 256, 2);
      cgm_x_kernel<T><<<grid, 256, 0, ctx->stream>>>(a, n);
      GLB_LAUNCH_CHECK();
      return GLB_OK;
    });
    if (rc) return rc;
  }
  return GLB_OK;
}

int glb_cgm_update_p(glb_context* ctx, int dtype, size_t n, int ns, const double* zeta, const double* alpha_s,
                     const void* r, void* const* p_s) {
  for (int base = 0; base < ns; base += MAXK) {
    const int kk = (ns - base < MAXK) ? ns - base : MAXK;
    int rc = dispatch(dtype, [&](auto tag) {
      typedef decltype(tag) T;
      MultiArgs<T> a{};
      a.k = kk;
      for (int j = 0; j < kk; j++) {
        a.Y[j] = (T*)p_s[base + j];
        a.c0[j] = coef<T>(zeta + 2 * (base + j));
        a.c1[j] = coef<T>(alpha_s + 2 * (base + j));
      }
      const int grid = blas_grid(ctx, n, 256, 2);
      cgm_p_kernel<T><<<grid, 256, 0, ctx->stream>>>(a, (const T*)r, n);
      GLB_LAUNCH_CHECK();
      return GLB_OK;
    });
    if (rc) return rc;
  }
  return GLB_OK;
}

}  // extern "C"
