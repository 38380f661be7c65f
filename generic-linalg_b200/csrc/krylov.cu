// krylov.cu -- device-resident loops of minv_vector_bicgstab (generic_bicgstab.cpp:258-308 complex, :75-125 real)
// and minv_vector_cr (generic_cr.cpp:246-286 complex, :76-116 real): no host round trip per iteration.
//
// The host-scalar shells (host/dev_solvers.cpp) read every inner product back before they can form alpha / omega /
// beta: three (BiCGStab) or four (CR) stream synchronisations per iteration, which is what bounds them below ~2048^2.
// Here the scalars never leave the GPU.  Every reduction kernel of the library leaves its grid totals in
// ReduceWs::result_dev; the NEXT kernel on the stream forms the scalar it needs from those totals in its prologue
// (a couple of divisions, done redundantly by every thread), and the last block of a reducing kernel runs the
// recurrence / stopping test in its epilogue.  Scalars that must survive more than one kernel live in KrylovState.
// The vector kernels and their launch geometry are the shells' own (same functor arithmetic, same grid, same
// reduction tree), so given the same inputs the iterates are bit-identical to the host-scalar shells and the
// iteration counts are the reference's.
//
// Per iteration (bytes per site, complex): BiCGStab 48 + 64 + 112 + 48 + 80 = 352 (SURVEY 8 d-bytes fused minimum: 336),
// CR 96 + 80 + 96 = 272 (<Ap,r> rides on the p / Ap update).  The host enqueues BATCH iterations at a time -- directly or as ONE CUDA-graph launch
// (the kernel arguments never change between iterations) -- and polls the state of the previous batch while the next
// one runs; kernels past the stopping point return at once, the operator applies through CgState::done.
#include <cstdlib>

#include "cg_state.cuh"
#include "runtime.hpp"

namespace glb {

int op_apply_fused(glb_operator* op, void* out, const void* in, const ApplyFusion& f);

namespace {

template <typename T, int W>
struct alignas(sizeof(T) * W) KPack {
  T v[W];
};
template <int NV>
struct KPtrs {
  void* p[NV];
};

__device__ __forceinline__ void put(double* d, cplx v) {
  d[0] = v.x;
  d[1] = v.y;
}
__device__ __forceinline__ void put(double* d, double v) {
  d[0] = v;
  d[1] = 0.0;
}
__device__ __forceinline__ cplx fconj(cplx a) { return mk(a.x, -a.y); }
__device__ __forceinline__ double fconj(double a) { return a; }
__device__ __forceinline__ void lift(double r, cplx* out) { *out = mk(r, 0.0); }  // T(r)
__device__ __forceinline__ void lift(double r, double* out) { *out = r; }

// the stopping test every solver of the family shares (generic_bicgstab.cpp:289, generic_cr.cpp:265): returns true
// when the loop ends with this iteration
__device__ __forceinline__ bool stop_test(KrylovState* st, double* hist, double rsq) {
  st->cg.rsq_new = rsq;
  const int k = st->cg.iter;  // 0-based index of the reference's loop
  st->cg.iter = k + 1;
  if (hist != nullptr && k < st->cg.hist_cap) hist[k] = rsq;
  const bool conv = sqrt(rsq) < st->cg.eps * st->cg.bnorm;
  const bool last = (k == st->cg.max_iter - 1);
  if (conv || last) {
    st->cg.hit_max = last ? 1 : 0;  // the reference tests k alone after the loop
    st->cg.done = 1;
    return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------ functors
// interface: prologue(st, res) loads / forms the scalars (every thread); persist(st) by one thread of the grid;
// elem() per element; epilogue(total, st, hist) by thread 0 of the block that finished the grid sum

// BiCGStab step 4-5 (generic_bicgstab.cpp:261-267): alpha = rho / <r0,Ap> ; s = r - alpha Ap      vectors: Ap, r, s
template <typename T>
struct FBicgS {
  static constexpr int NV = 3, RD = 3, WR = 4, NRED = 0;
  T alpha, nalpha;
  __device__ void prologue(const KrylovState* st, const double* res) {
    alpha = cdiv(Field<T>::from(st->rho), Field<T>::from(res));  // res = <r0,Ap> of the apply before this kernel
    nalpha = fneg(alpha);
  }
  __device__ void persist(KrylovState* st) const { put(st->alpha, alpha); }
  __device__ void elem(T (&e)[NV], double*) const { e[2] = fadd(e[1], fmul(nalpha, e[0])); }
  __device__ void epilogue(const double*, KrylovState*, double*) const {}
};

// steps 6-9 (generic_bicgstab.cpp:270-298): omega = <As,s>/<As,As> ; x += alpha p + omega s ; r = s - omega As ;
// |r|^2, <r0,r> ; stopping test ; beta = rhoNew/rho * (alpha/omega)                       vectors: p, s, As, r0, x, r
template <typename T>
struct FBicgXR {
  static constexpr int NV = 6, RD = 31, WR = 48, NRED = 1 + Field<T>::NCOMP;
  T alpha, omega;
  __device__ void prologue(const KrylovState* st, const double* res) {
    // res = <s,As> (NCOMP doubles), |As|^2 ; generic_bicgstab.cpp:271 divides dot(As,s) by the complex dot(As,As)
    T den;
    lift(res[Field<T>::NCOMP], &den);
    omega = cdiv(fconj(Field<T>::from(res)), den);
    alpha = Field<T>::from(st->alpha);
  }
  __device__ void persist(KrylovState*) const {}
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[4] = fadd(fadd(e[4], fmul(alpha, e[0])), fmul(omega, e[1]));  // phi = phi + alpha*p + omega*s
    e[5] = fsub(e[1], fmul(omega, e[2]));                            // r = s - omega*As
    acc[0] += fnorm(e[5]);
    Field<T>::dot_acc(acc + 1, e[3], e[5]);                          // <r0, r>
  }
  __device__ void epilogue(const double* total, KrylovState* st, double* hist) const {
    put(st->omega, omega);
    if (stop_test(st, hist, total[0])) return;
    const T rho = Field<T>::from(st->rho);
    const T rho_new = Field<T>::from(total + 1);
    put(st->beta, fmul(cdiv(rho_new, rho), cdiv(alpha, omega)));  // beta = rhoNew/rho*(alpha/omega)
    put(st->rho, rho_new);
  }
};

// step 10 (generic_bicgstab.cpp:300-303): p = r + beta*(p - omega*Ap)                      vectors: r, Ap, p
template <typename T>
struct FBicgP {
  static constexpr int NV = 3, RD = 7, WR = 4, NRED = 0;
  T beta, omega;
  __device__ void prologue(const KrylovState* st, const double*) {
    beta = Field<T>::from(st->beta);
    omega = Field<T>::from(st->omega);
  }
  __device__ void persist(KrylovState*) const {}
  __device__ void elem(T (&e)[NV], double*) const { e[2] = fadd(e[0], fmul(beta, fsub(e[2], fmul(omega, e[1])))); }
  __device__ void epilogue(const double*, KrylovState*, double*) const {}
};

// CR: alpha = <Ap,r>/|Ap|^2 ; x += alpha p ; r -= alpha Ap ; |r|^2 ; stopping test (generic_cr.cpp:249-265)
//                                                                                          vectors: p, x, Ap, r
template <typename T>
struct FCrXR {
  static constexpr int NV = 4, RD = 15, WR = 10, NRED = 1;
  T a, b;
  __device__ void prologue(const KrylovState* st, const double*) {
    a = frdiv(Field<T>::from(st->alpha), st->rho[0]);  // <Ap,r> / |Ap|^2 ; complex / double: component-wise
    b = fneg(a);
  }
  __device__ void persist(KrylovState*) const {}
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[1] = fadd(e[1], fmul(a, e[0]));
    e[3] = fadd(e[3], fmul(b, e[2]));
    acc[0] += fnorm(e[3]);
  }
  __device__ void epilogue(const double* total, KrylovState* st, double* hist) const { stop_test(st, hist, total[0]); }
};

// CR: beta = -<Ap,Ar>/|Ap|^2 ; p = r + beta p ; Ap = Ar + beta Ap ; |Ap|^2 (generic_cr.cpp:275-285), and the next
// iteration's <Ap,r> (:249) while Ap and r are in registers -- same elements per thread, same order and same
// reduction tree as the separate dot kernel of the shell, so the same bits      vectors: r, Ar, p, Ap
template <typename T>
struct FCrPAp {
  static constexpr int NV = 4, RD = 15, WR = 12, NRED = 1 + Field<T>::NCOMP;
  T beta;
  __device__ void prologue(const KrylovState* st, const double* res) {
    beta = frdiv(fneg(Field<T>::from(res)), st->rho[0]);
  }
  __device__ void persist(KrylovState*) const {}
  __device__ void elem(T (&e)[NV], double* acc) const {
    e[2] = fadd(e[0], fmul(beta, e[2]));
    e[3] = fadd(e[1], fmul(beta, e[3]));
    acc[0] += fnorm(e[3]);
    Field<T>::dot_acc(acc + 1, e[3], e[0]);
  }
  __device__ void epilogue(const double* total, KrylovState* st, double*) const {
    st->rho[0] = total[0];
    put(st->alpha, Field<T>::from(total + 1));
  }
};

// The streaming kernel of blas1.cu (32 bytes per vector per thread and step, grid-stride loop, deterministic grid
// sum) with the scalar prologue / epilogue around it.
template <typename T, typename F, int W>
__global__ void __launch_bounds__(256)
ews_kernel(F f, KPtrs<F::NV> ptrs, size_t n, ReduceWs red, KrylovState* st, double* hist) {
  if (st->cg.done) return;
  f.prologue(st, red.result_dev);
  if (blockIdx.x == 0 && threadIdx.x == 0) f.persist(st);
  constexpr int NRED = F::NRED > 0 ? F::NRED : 1;
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x * W;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * W; i < n; i += stride) {
    KPack<T, W> v[F::NV];
#pragma unroll
    for (int k = 0; k < F::NV; k++)
      if ((F::RD >> k) & 1) v[k] = *reinterpret_cast<const KPack<T, W>*>((const T*)ptrs.p[k] + i);
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
      if ((F::WR >> k) & 1) *reinterpret_cast<KPack<T, W>*>((T*)ptrs.p[k] + i) = v[k];
  }
  if (F::NRED > 0) {
    // every other block has read result_dev and the state in its prologue before it took its ticket, so the block
    // that finishes the sum may overwrite both
    double total[NRED];
    if (grid_sum<NRED>(acc, red, total) && threadIdx.x == 0) f.epilogue(total, st, hist);
  }
}

template <typename T, typename F>
int run_ews(glb_context* ctx, const KPtrs<F::NV>& ptrs, size_t n, KrylovState* st, double* hist) {
  constexpr int WMAX = 32 / sizeof(T);
  bool wide = (n % WMAX == 0);
  for (int k = 0; k < F::NV; k++) wide = wide && (((uintptr_t)ptrs.p[k] & 31u) == 0);
  ReduceWs red = ctx->red;
  red.result_host = nullptr;
  F f{};
  ProfScope prof(ctx, PROF_EW, (double)n * sizeof(T) * (__builtin_popcount(F::RD) + __builtin_popcount(F::WR)));
  if (wide) {
    const int grid = blas_grid(ctx, n / WMAX, 256, 2);
    ews_kernel<T, F, WMAX><<<grid, 256, 0, ctx->stream>>>(f, ptrs, n, red, st, hist);
  } else {
    const int grid = blas_grid(ctx, n, 256, 4);
    ews_kernel<T, F, 1><<<grid, 256, 0, ctx->stream>>>(f, ptrs, n, red, st, hist);
  }
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

// out = A in with <w,out> (and |out|^2) left in result_dev; returns at once after the stopping test fired
int apply_red(glb_operator* op, KrylovState* st, void* out, const void* in, const void* w, bool want_norm) {
  ApplyFusion f;
  f.w = w;
  f.w_is_input = (w == in);
  f.want_norm = want_norm;
  f.to_host = false;
  f.cg_state = (const double*)st;
  f.cg_role = 0;
  return op_apply_fused(op, out, in, f);
}

int g_graph_mode = -1;       // -1: read GLB_KRYLOV_GRAPH on first use
int g_last_used_graph = 0;   // diagnostic: the last solve replayed a CUDA graph
bool graph_mode() {
  if (g_graph_mode < 0) {
    const char* e = getenv("GLB_KRYLOV_GRAPH");
    g_graph_mode = (e && *e) ? (atoi(e) != 0 ? 1 : 0) : 1;
  }
  return g_graph_mode == 1;
}

// Enqueue iterations in batches until the state says the loop has ended (same protocol as glb_cg_solve): the state
// after batch i is copied to a pinned slot asynchronously and looked at while batch i+1 runs.  enqueue_iteration(i)
// gets the index of the iteration it enqueues (only its parity matters: BATCH is even, a captured batch starts even).
template <typename Enq>
int run_batches(glb_context* ctx, KrylovState* d_st, int max_iter, KrylovState* fin, Enq&& enqueue_iteration) {
  const int BATCH = 8;
  KrylovState* h_st = (KrylovState*)ctx->h_table;
  static_assert(2 * sizeof(KrylovState) <= 4096, "pinned scratch of the context is 4 KiB");
  int rc = GLB_OK;
  cudaGraphExec_t exec = nullptr;
  unsigned long long graph_kernels = 0;  // kernels one replay launches (glb_kernel_launches counts kernels, not graphs)
  g_last_used_graph = 0;
  // the first batch is launched directly (it also loads every kernel); later batches replay ONE graph
  bool try_graph = graph_mode() && !ctx->prof_on && max_iter > BATCH;
  int enq = 0;
  bool finished = false, have_pending = false;
  while (!finished) {
    if (exec) {
      if (cudaGraphLaunch(exec, ctx->stream) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "cudaGraphLaunch");
        goto out;
      }
      g_launches.fetch_add(graph_kernels, std::memory_order_relaxed);
      g_last_used_graph = 1;
      enq += BATCH;
    } else {
      for (int b = 0; b < BATCH; b++) {
        if ((rc = enqueue_iteration(enq))) goto out;
        enq++;
      }
    }
    {
      const int slot = (enq / BATCH) & 1;
      if (have_pending) {
        if (cudaEventSynchronize(ctx->ev_a) != cudaSuccess) {
          rc = fail(GLB_ERR_CUDA, "cudaEventSynchronize");
          goto out;
        }
        if (h_st[slot ^ 1].cg.done) finished = true;
      }
      if (cudaMemcpyAsync(&h_st[slot], d_st, sizeof(KrylovState), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaEventRecord(ctx->ev_a, ctx->stream) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "state readback");
        goto out;
      }
      have_pending = true;
      if (enq >= max_iter + BATCH) finished = true;  // everything that could run has been enqueued
    }
    if (try_graph && !finished) {
      // one batch as a graph: the arguments of the kernels do not change from one iteration to the next
      try_graph = false;
      cudaGraph_t graph = nullptr;
      const unsigned long long l0 = g_launches.load(std::memory_order_relaxed);
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        int erc = GLB_OK;
        for (int b = 0; b < BATCH && erc == GLB_OK; b++) erc = enqueue_iteration(b);  // enq is a multiple of BATCH here
        const cudaError_t e1 = cudaStreamEndCapture(ctx->stream, &graph);
        if (erc != GLB_OK || e1 != cudaSuccess || graph == nullptr ||
            cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
          exec = nullptr;
        if (graph) cudaGraphDestroy(graph);
      }
      // the launches counted while capturing were recorded, not run
      graph_kernels = g_launches.load(std::memory_order_relaxed) - l0;
      g_launches.fetch_sub(graph_kernels, std::memory_order_relaxed);
      cudaGetLastError();  // a failed capture must not poison the direct path
    }
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    rc = fail(GLB_ERR_CUDA, "cudaStreamSynchronize");
    goto out;
  }
  if (cudaMemcpy(fin, d_st, sizeof(KrylovState), cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = fail(GLB_ERR_CUDA, "final state readback");
out:
  if (exec) cudaGraphExecDestroy(exec);
  return rc;
}

template <typename T>
int krylov_solve_t(glb_operator* op, int alg, void* d_x, const void* d_b, int max_iter, double eps, glb_cg_report* rep,
                   double* rsq_hist, int hist_cap) {
  glb_context* ctx = op->ctx;
  const int dt = op->dtype;
  const size_t n = glb_op_local_size(op);
  const int nvec = (alg == GLB_KRYLOV_BICGSTAB) ? 6 : 4;
  void* v[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  KrylovState* d_st = nullptr;
  double* d_hist = nullptr;
  KrylovState* h_st = (KrylovState*)ctx->h_table;
  KrylovState fin{};
  int rc = GLB_OK;
#define KR_TRY(x)      \
  do {                 \
    rc = (x);          \
    if (rc) goto done; \
  } while (0)
  for (int i = 0; i < nvec; i++) KR_TRY(glb_vec_alloc(ctx, dt, n, &v[i]));
  if (cudaMallocAsync((void**)&d_st, sizeof(KrylovState), ctx->stream) != cudaSuccess) {
    rc = fail(GLB_ERR_CUDA, "cudaMallocAsync(KrylovState)");
    goto done;
  }
  if (hist_cap > 0 && rsq_hist) {
    if (cudaMallocAsync((void**)&d_hist, sizeof(double) * hist_cap, ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "cudaMallocAsync(hist)");
      goto done;
    }
  }
  {
    KrylovState init{};
    double bsq = 0.0;
    KR_TRY(glb_norm2sq(ctx, dt, n, d_b, &bsq));
    init.cg.bnorm = sqrt(bsq);
    init.cg.eps = eps;
    init.cg.max_iter = max_iter;
    init.cg.hist_cap = d_hist ? hist_cap : 0;
    rep->bnorm = init.cg.bnorm;
    if (alg == GLB_KRYLOV_BICGSTAB) {
      void *r = v[0], *r0 = v[1], *p = v[2], *Ap = v[3], *s = v[4], *As = v[5];
      // set-up of generic_bicgstab.cpp:243-256: r = b - A x, r0 = p = r, rho = <r0,r>, Ap = A p
      KR_TRY(glb_op_apply(op, Ap, d_x));
      KR_TRY(glb_sub(ctx, dt, n, d_b, Ap, r));
      KR_TRY(glb_vec_copy(ctx, dt, n, r0, r));
      KR_TRY(glb_vec_copy(ctx, dt, n, p, r));
      double rho[2] = {0.0, 0.0};
      KR_TRY(glb_dot(ctx, dt, n, r0, r, rho));
      init.rho[0] = rho[0];
      init.rho[1] = rho[1];
      h_st[0] = init;
      if (cudaMemcpyAsync(d_st, &h_st[0], sizeof(KrylovState), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "upload KrylovState");
        goto done;
      }
      KR_TRY(apply_red(op, d_st, Ap, p, r0, false));  // <r0,Ap> stays in result_dev for the first s kernel
      const KPtrs<3> k1{{Ap, r, s}};
      const KPtrs<6> k3{{p, s, As, r0, d_x, r}};
      const KPtrs<3> k4{{r, Ap, p}};
      KR_TRY(run_batches(ctx, d_st, max_iter, &fin, [&](int) -> int {
        int e;
        if ((e = run_ews<T, FBicgS<T>>(ctx, k1, n, d_st, nullptr))) return e;
        if ((e = apply_red(op, d_st, As, s, s, true))) return e;  // As = A s ; <s,As>, |As|^2
        if ((e = run_ews<T, FBicgXR<T>>(ctx, k3, n, d_st, d_hist))) return e;
        if ((e = run_ews<T, FBicgP<T>>(ctx, k4, n, d_st, nullptr))) return e;
        return apply_red(op, d_st, Ap, p, r0, false);  // Ap = A p ; <r0,Ap>
      }));
      rep->ops = 2 + fin.cg.iter + (fin.cg.iter > 0 ? fin.cg.iter - 1 : 0);  // generic_bicgstab.cpp:270,305
    } else {
      void *r = v[0], *Ar = v[1], *p = v[2], *Ap = v[3];
      // set-up of generic_cr.cpp:229-243: r = b - A x, p = r, Ap = A p, Ar = Ap, |Ap|^2
      KR_TRY(glb_op_apply(op, p, d_x));
      KR_TRY(glb_sub(ctx, dt, n, d_b, p, r));
      KR_TRY(glb_vec_copy(ctx, dt, n, p, r));
      KR_TRY(glb_op_apply(op, Ap, p));
      KR_TRY(glb_vec_copy(ctx, dt, n, Ar, Ap));
      double apsq = 0.0, apr[2] = {0.0, 0.0};
      KR_TRY(glb_norm2sq(ctx, dt, n, Ap, &apsq));
      KR_TRY(glb_dot(ctx, dt, n, Ap, r, apr));  // <Ap,r> of the first iteration; later ones come from the p/Ap update
      init.rho[0] = apsq;
      init.alpha[0] = apr[0];
      init.alpha[1] = apr[1];
      h_st[0] = init;
      if (cudaMemcpyAsync(d_st, &h_st[0], sizeof(KrylovState), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "upload KrylovState");
        goto done;
      }
      const KPtrs<4> k1{{p, d_x, Ap, r}};
      const KPtrs<4> k3{{r, Ar, p, Ap}};
      KR_TRY(run_batches(ctx, d_st, max_iter, &fin, [&](int) -> int {
        int e;
        if ((e = run_ews<T, FCrXR<T>>(ctx, k1, n, d_st, d_hist))) return e;
        if ((e = apply_red(op, d_st, Ar, r, Ap, false))) return e;  // Ar = A r ; <Ap,Ar>
        return run_ews<T, FCrPAp<T>>(ctx, k3, n, d_st, nullptr);
      }));
      rep->ops = 2 + (fin.cg.iter > 0 ? fin.cg.iter - 1 : 0);  // generic_cr.cpp:272
    }
    rep->iterations = fin.cg.iter;
    rep->hit_max_iter = fin.cg.hit_max;
    rep->rsq = fin.cg.rsq_new;
    if (d_hist) {
      const int m = fin.cg.iter < hist_cap ? fin.cg.iter : hist_cap;
      if (m > 0 && cudaMemcpy(rsq_hist, d_hist, sizeof(double) * m, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "history readback");
        goto done;
      }
    }
  }
done:
  if (rc != GLB_OK) cudaStreamSynchronize(ctx->stream);  // nothing may still be using the vectors freed below
  if (d_hist) cudaFreeAsync(d_hist, ctx->stream);
  if (d_st) cudaFreeAsync(d_st, ctx->stream);
  for (int i = nvec - 1; i >= 0; i--) glb_vec_free(ctx, v[i]);
  return rc;
#undef KR_TRY
}

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" int glb_krylov_solve_supported(const glb_operator* op, int alg) {
  if (!op || (alg != GLB_KRYLOV_BICGSTAB && alg != GLB_KRYLOV_CR)) return 0;
  if (op->ctx->nranks != 1) return 0;  // slabs: the host-scalar shells (rank-wide sums through allreduce_sum)
  if (op->composite || op->kind == OPK_GAMMA5) return 0;  // no fused reductions in the apply
  return 1;
}

extern "C" int glb_krylov_solve(glb_operator* op, int alg, void* d_x, const void* d_b, int max_iter, double eps,
                                glb_cg_report* rep, double* rsq_hist, int hist_cap) {
  if (!op || !d_x || !d_b || !rep) return fail(GLB_ERR_ARG, "glb_krylov_solve: null argument");
  if (max_iter < 1) return fail(GLB_ERR_ARG, "glb_krylov_solve: max_iter must be >= 1");
  if (!glb_krylov_solve_supported(op, alg))
    return fail(GLB_ERR_STATE, "glb_krylov_solve: operator / algorithm not supported (see glb_krylov_solve_supported)");
  GLB_CUDA(cudaSetDevice(op->ctx->device));
  if (op->dtype == GLB_COMPLEX) return krylov_solve_t<cplx>(op, alg, d_x, d_b, max_iter, eps, rep, rsq_hist, hist_cap);
  return krylov_solve_t<double>(op, alg, d_x, d_b, max_iter, eps, rep, rsq_hist, hist_cap);
}

extern "C" int glb_krylov_graph_mode(int on) {
  const int prev = graph_mode() ? 1 : 0;
  if (on >= 0) g_graph_mode = on ? 1 : 0;
  return prev;
}

extern "C" int glb_krylov_last_used_graph(void) { return g_last_used_graph; }
