// cg.cu -- device-resident conjugate gradient: the loop of minv_vector_cg
// (generic_cg.cpp:159-206 real, :307-354 complex) with no host round trip per iteration.
//
// Per iteration the stream carries
//   K3  cg_update_kernel : alpha = rsq/<p,Ap>; x += alpha p; r -= alpha Ap; rsqNew = |r|^2;
//                          last block: iter++, stopping test, history
//   K1  direction + apply: beta = rsqNew/rsq; p = r + beta p   fused into the operator apply
//                          (staggered family, single rank) or a separate xpay kernel
//   K2  (normal operator): Ap = D^dag (D p) with <p,Ap> accumulated in the apply's epilogue;
//                          last block publishes <p,Ap> and rsq <- rsqNew
// Fused minimum traffic (SURVEY 8 d-bytes): 192 B/site for a single-apply Hermitian operator,
// 272 B/site for D^dag D.  The host enqueues BATCH iterations at a time and polls the state of
// the previous batch while the next one runs; kernels of iterations past the stopping point
// return at once, so the iteration count is exactly the reference's.
#include <vector>

#include "cg_state.cuh"
#include "cgstep.cuh"
#include "runtime.hpp"

namespace glb {

static double g_last_pred_err = 0.0;  // diagnostic of the last single-kernel solve (glb_cg_last_pred_err)

int launch_cg_update(glb_context* ctx, int dtype, void* st, double* hist, const void* p, void* x, const void* Ap,
                     void* r, size_t n, int defer);
int launch_cg_post_update(glb_context* ctx, void* st, double* hist);
int launch_cg_post_apply(glb_context* ctx, void* st);
int launch_cg_boundary_push(glb_context* ctx, const void* st, const void* r, const void* pold, size_t row_elems, int nrows,
                            size_t local_elems, const HaloTargets& t);
int launch_cg_boundary(glb_context* ctx, const void* st, const void* r, const void* pold, void* send_lo, void* send_hi,
                       size_t row_elems, int nrows, size_t local_elems);
int launch_cg_xpay(glb_context* ctx, int dtype, const void* st, const void* r, void* p, size_t n);
int op_apply_fused(glb_operator* op, void* out, const void* in, const ApplyFusion& f);

static bool can_fuse_direction(const glb_operator* op) {
  if (op->ctx->nranks > 1) return normal_fused_ok(op);  // slabs: only the one-pass D^dag D kernel
  if (op->kind == OPK_STAGGERED && (op->flags & (GLB_STAG_DEO | GLB_STAG_DOE | GLB_STAG_M2MDEODOE))) return false;
  return (op->kind == OPK_STAGGERED || op->kind == OPK_LAPLACE_U1) && op->X >= 2;
}

// p_next = r + beta p_cur ; Ap = A p_next ; <p_next, Ap> -> state.  Returns which buffer holds p.
static int direction_and_apply(glb_operator* op, CgState* d_st, const void* r, void* p_cur, void* p_alt, void* Ap,
                               bool* swapped) {
  glb_context* ctx = op->ctx;
  const size_t n = glb_op_local_size(op);
  *swapped = false;
  if (ctx->nranks > 1) {
    // slab run: boundary rows of the new direction -> neighbours' ghost rows, then the one-pass
    // kernel (own rows formed on the fly, ghost rows read as they are), then the rank sum of <p,Ap>
    const size_t row = (size_t)op->X;
    int rc;
    if (comm_p2p(ctx) && op->ghost_p2p) {
      // peer memory: boundary rows are computed straight into the neighbours' ghost rows, the one-pass
      // kernel waits for its own ghost flags in its prologue and its last block sums <p,Ap> over ranks
      HaloTargets t;
      if ((rc = halo_p2p_begin(op, 2, &t))) return rc;
      if ((rc = launch_cg_boundary_push(ctx, d_st, r, p_cur, row, 2, n, t))) return rc;
      ApplyFusion f;
      f.r = r;
      f.p_old = p_cur;
      f.p_new = p_alt;
      f.cg_state = (const double*)d_st;
      f.w = p_alt;
      f.w_is_input = true;
      f.cg_role = 3;
      f.pr = comm_p2p_red(ctx);
      f.wait = t.wait;
      *swapped = true;
      return launch_normal(op, Ap, nullptr, f);
    }
    if ((rc = launch_cg_boundary(ctx, d_st, r, p_cur, op->send_lo, op->send_hi, row, 2, n))) return rc;
    if ((rc = halo_exchange_ptrs(op, op->send_lo, op->send_hi, 2))) return rc;
    ApplyFusion f;
    f.r = r;
    f.p_old = p_cur;
    f.p_new = p_alt;
    f.cg_state = (const double*)d_st;
    f.w = p_alt;
    f.w_is_input = true;
    f.cg_role = 2;
    *swapped = true;
    if ((rc = launch_normal(op, Ap, nullptr, f))) return rc;
    if ((rc = allreduce_device(ctx, d_st->partial + 1, 2))) return rc;
    return launch_cg_post_apply(ctx, d_st);
  }
  if (can_fuse_direction(op)) {
    ApplyFusion f;
    f.r = r;
    f.p_old = p_cur;
    f.p_new = p_alt;
    f.cg_state = (const double*)d_st;
    *swapped = true;
    if ((op->flags & GLB_STAG_NORMAL) && normal_fused_ok(op)) {
      // one pass: p_alt = r + beta p ; Ap = D^dag D p_alt ; <p_alt,Ap>   (96 B/site)
      f.w = p_alt;
      f.w_is_input = true;
      f.cg_role = 1;
      return launch_normal(op, Ap, nullptr, f);
    }
    if (op->flags & GLB_STAG_NORMAL) {
      // K1: t = D (r + beta p), p_alt = r + beta p
      int rc = launch_staggered(op, op->tmp, nullptr, false, f);
      if (rc) return rc;
      // K2: Ap = D^dag t, <p_alt, Ap>
      ApplyFusion g;
      g.w = p_alt;
      g.cg_state = (const double*)d_st;
      g.cg_role = 1;
      return launch_staggered(op, Ap, op->tmp, true, g);
    }
    f.w = p_alt;  // dot partner = the freshly formed direction = the kernel's own input
    f.w_is_input = true;
    f.cg_role = 1;
    return launch_staggered(op, Ap, nullptr, (op->flags & GLB_STAG_DAGGER) != 0, f);
  }
  int rc = launch_cg_xpay(ctx, op->dtype, d_st, r, p_cur, n);
  if (rc) return rc;
  ApplyFusion g;
  g.w = p_cur;
  g.w_is_input = true;
  g.cg_state = (const double*)d_st;
  g.cg_role = 1;
  return op_apply_fused(op, Ap, p_cur, g);
}


// ---------------------------------------------------------------------------------------------------------
// The single-kernel iteration (cgstep.cu): one launch and one rank-wide reduction per CG iteration, 160 B/site.
// Arena layout of an operator's ghost area on slabs: [parity 0: lo | hi][parity 1: lo | hi][flags + counters],
// each of lo / hi = 3 vectors x 2 rows x X complex.
static int cs_prepare_slab(glb_operator* op) {
  if (op->cs_ready) return GLB_OK;
  const size_t gbytes = (size_t)3 * 2 * op->X * sizeof(cplx);
  size_t off = 0;
  unsigned long long seq0 = 0;
  if (!comm_arena_alloc(op->ctx, 4 * gbytes + 256, &off, &seq0))
    return fail(GLB_ERR_STATE, "peer-memory arena exhausted (GLB_P2P_ARENA_MB)");
  op->cs_off = off;
  op->cs_seq = seq0;
  op->cs_ready = true;
  return GLB_OK;
}

// the slab part of the kernel arguments: this rank's ghost buffers, the neighbours', flags and counters
static void cs_slab_args(glb_operator* op, CgStepArgs* a) {
  glb_context* ctx = op->ctx;
  const int G = ctx->nranks, g = ctx->rank;
  const int up = (g + 1) % G, down = (g + G - 1) % G;
  const size_t gbytes = (size_t)3 * 2 * op->X * sizeof(cplx);
  const size_t off_flag = op->cs_off + 4 * gbytes;
  char* mine = comm_peer(ctx, g);
  a->ghost = (const cplx*)(mine + op->cs_off);
  a->peer_down = (cplx*)(comm_peer(ctx, down) + op->cs_off);
  a->peer_up = (cplx*)(comm_peer(ctx, up) + op->cs_off);
  a->flag_down = (unsigned long long*)(comm_peer(ctx, down) + off_flag + 8);  // its flag_hi
  a->flag_up = (unsigned long long*)(comm_peer(ctx, up) + off_flag);          // its flag_lo
  a->push_count = (unsigned int*)(mine + off_flag + 16);
  a->wait.flag_lo = (const unsigned long long*)(mine + off_flag);
  a->wait.flag_hi = (const unsigned long long*)(mine + off_flag + 8);
  a->wait.seq = 0;
}

static int cg_solve_step(glb_operator* op, void* d_x, const void* d_b, int max_iter, double eps, glb_cg_report* rep,
                         double* rsq_hist, int hist_cap) {
  glb_context* ctx = op->ctx;
  const int dt = op->dtype;
  const size_t n = glb_op_local_size(op);
  const bool slab = ctx->nranks > 1;
  int rc = GLB_OK;
  void* v[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // r0 r1 p0 p1 q0 q1
  CgState* d_st = nullptr;
  double* d_hist = nullptr;
  CgState* h_st = (CgState*)ctx->h_table;
#define CS_TRY(x)      \
  do {                 \
    rc = (x);          \
    if (rc) goto done; \
  } while (0)
  if (slab) CS_TRY(cs_prepare_slab(op));
  for (int i = 0; i < 6; i++) CS_TRY(glb_vec_alloc(ctx, dt, n, &v[i]));
  if (cudaMallocAsync((void**)&d_st, sizeof(CgState), ctx->stream) != cudaSuccess) {
    rc = fail(GLB_ERR_CUDA, "cudaMallocAsync(CgState)");
    goto done;
  }
  if (hist_cap > 0 && rsq_hist) {
    if (cudaMallocAsync((void**)&d_hist, sizeof(double) * hist_cap, ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "cudaMallocAsync(hist)");
      goto done;
    }
  }
  {
    // set-up of generic_cg.cpp:304-321: bnorm, r = b - A x; the first apply (Ap = A p with p = r) and rsq = |r|^2
    // are step 0 of the kernel (alpha = beta = 0, p_old = q_old = 0)
    double bsq = 0.0;
    CS_TRY(glb_norm2sq(ctx, dt, n, d_b, &bsq));
    CS_TRY(glb_op_apply(op, v[2], d_x));
    CS_TRY(glb_sub(ctx, dt, n, d_b, v[2], v[0]));
    CS_TRY(glb_vec_zero(ctx, dt, n, v[2]));
    CS_TRY(glb_vec_zero(ctx, dt, n, v[4]));
    CgState init{};
    init.bnorm = sqrt(bsq);
    init.eps = eps;
    init.max_iter = max_iter;
    init.hist_cap = d_hist ? hist_cap : 0;
    h_st[0] = init;
    if (cudaMemcpyAsync(d_st, &h_st[0], sizeof(CgState), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "upload CgState");
      goto done;
    }
    rep->bnorm = init.bnorm;

    CgStepArgs a{};
    for (int c = 0; c < 2; c++) {
      a.r[c] = (cplx*)v[0 + c];
      a.p[c] = (cplx*)v[2 + c];
      a.q[c] = (cplx*)v[4 + c];
    }
    a.x = (cplx*)d_x;
    a.Ux = op->Ux;
    a.Uy = op->Uy;
    a.X = op->X;
    a.Y = op->Yloc;
    a.mass = op->mass;
    a.st = d_st;
    a.hist = d_hist;
    a.red = ctx->red;
    a.red.result_host = nullptr;
    a.wait.budget = comm_spin_budget(ctx);
    if (a.wait.budget == 0 && !slab) a.wait.budget = 0;  // single rank: the step spin never gives up on its own grid
    const long long max_steps = (long long)max_iter + 1;  // step 0 is the set-up pass
    if (slab) {
      // boundary rows of (r, 0, 0) to the neighbours: what step 0 reads as ghost rows.  Exchange numbers: the set-up
      // push is seq0, step s is seq0 + 1 + s; a range is reserved so that every rank counts alike whatever happens
      cs_slab_args(op, &a);
      const unsigned long long seq0 = op->cs_seq + 1;
      op->cs_seq += 1 + (unsigned long long)max_steps;
      a.seq_base = seq0 + 1;
      CS_TRY(launch_cg_step_halo_init(op, v[0], v[4], v[2], a, seq0, comm_ticket(ctx)));
      a.pr = comm_p2p_red_range(ctx, (unsigned long long)max_steps);
    }
    if (cg_step_persistent()) {
      // ONE launch for the whole solve: the CTAs stay resident and meet at the reduction of every step
      a.nsteps = max_steps > 0x7fffffffLL ? 0x7fffffff : (int)max_steps;
      CS_TRY(launch_cg_step(op, a));
    } else {
      a.nsteps = 1;
      const int BATCH = 8;
      long long enq = 0;
      bool finished = false, have_pending = false;
      while (!finished) {
        for (int b = 0; b < BATCH; b++) {
          CS_TRY(launch_cg_step(op, a));
          enq++;
        }
        const int slot = (int)((enq / BATCH) & 1);
        if (have_pending) {
          if (cudaEventSynchronize(ctx->ev_a) != cudaSuccess) {
            rc = fail(GLB_ERR_CUDA, "cudaEventSynchronize");
            goto done;
          }
          if (h_st[slot ^ 1].done) finished = true;
        }
        if (cudaMemcpyAsync(&h_st[slot], d_st, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaEventRecord(ctx->ev_a, ctx->stream) != cudaSuccess) {
          rc = fail(GLB_ERR_CUDA, "state readback");
          goto done;
        }
        have_pending = true;
        if (finished) break;
        if (enq >= max_steps + BATCH) finished = true;  // everything that could run has been enqueued
      }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "cudaStreamSynchronize");
      goto done;
    }
    if (cudaMemcpy(&h_st[0], d_st, sizeof(CgState), cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "final state readback");
      goto done;
    }
    const CgState fin = h_st[0];
    rep->iterations = fin.iter;
    rep->ops = 2 + (fin.iter > 0 ? fin.iter - 1 : 0);  // the reference's count (generic_cg.cpp:362)
    rep->hit_max_iter = fin.hit_max;
    rep->rsq = fin.rsq_new;
    g_last_pred_err = fin.pred_err;
    if (d_hist) {
      const int m = fin.iter < hist_cap ? fin.iter : hist_cap;
      if (m > 0 && cudaMemcpy(rsq_hist, d_hist, sizeof(double) * m, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "history readback");
        goto done;
      }
    }
  }
done:
  if (d_hist) cudaFreeAsync(d_hist, ctx->stream);
  if (d_st) cudaFreeAsync(d_st, ctx->stream);
  for (int i = 5; i >= 0; i--) glb_vec_free(ctx, v[i]);
  return rc;
#undef CS_TRY
}

}  // namespace glb

using namespace glb;

extern "C" int glb_cg_solve_supported(const glb_operator* op) {
  if (!op) return 0;
  if (op->composite) return 0;  // multi-pass stencil views: no fused epilogue, the host-scalar shell runs them
  if (op->kind == OPK_GAMMA5) return 0;  // gamma_5 alone has no fused reductions either
  return (op->ctx->nranks == 1 || normal_fused_ok(op)) ? 1 : 0;
}

extern "C" int glb_cg_solve(glb_operator* op, void* d_x, const void* d_b, int max_iter, double eps,
                            glb_cg_report* rep, double* rsq_hist, int hist_cap) {
  if (!op || !d_x || !d_b || !rep) return fail(GLB_ERR_ARG, "glb_cg_solve: null argument");
  if (max_iter < 1) return fail(GLB_ERR_ARG, "glb_cg_solve: max_iter must be >= 1");
  glb_context* ctx = op->ctx;
  if (ctx->nranks > 1 && !normal_fused_ok(op))
    return fail(GLB_ERR_STATE, "glb_cg_solve on slabs needs the one-pass D^dag D operator (see glb_cg_solve_supported)");
  if (cg_step_ok(op)) return cg_solve_step(op, d_x, d_b, max_iter, eps, rep, rsq_hist, hist_cap);
  const int dt = op->dtype;
  const size_t n = glb_op_local_size(op);
  int rc;
  void *r = nullptr, *p0 = nullptr, *p1 = nullptr, *Ap = nullptr;
  CgState* d_st = nullptr;
  double* d_hist = nullptr;
  CgState* h_st = nullptr;
  const bool fuse = can_fuse_direction(op);
#define CG_TRY(x) \
  do {            \
    rc = (x);     \
    if (rc) goto done; \
  } while (0)
  rc = GLB_OK;
  CG_TRY(glb_vec_alloc(ctx, dt, n, &r));
  CG_TRY(glb_vec_alloc(ctx, dt, n, &p0));
  if (fuse) CG_TRY(glb_vec_alloc(ctx, dt, n, &p1));
  CG_TRY(glb_vec_alloc(ctx, dt, n, &Ap));
  if (cudaMallocAsync((void**)&d_st, sizeof(CgState), ctx->stream) != cudaSuccess) {
    rc = fail(GLB_ERR_CUDA, "cudaMallocAsync(CgState)");
    goto done;
  }
  if (hist_cap > 0 && rsq_hist) {
    if (cudaMallocAsync((void**)&d_hist, sizeof(double) * hist_cap, ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "cudaMallocAsync(hist)");
      goto done;
    }
  }
  h_st = (CgState*)ctx->h_table;  // pinned scratch owned by the context (two slots used)
  {
    // --- set-up exactly as generic_cg.cpp:304-321: bnorm, r = b - A x, p = r, Ap = A p, rsq
    double bsq = 0.0, rsq = 0.0, dots[3];
    CG_TRY(glb_norm2sq(ctx, dt, n, d_b, &bsq));
    CG_TRY(glb_op_apply(op, p0, d_x));
    CG_TRY(glb_sub(ctx, dt, n, d_b, p0, r));
    CG_TRY(glb_vec_copy(ctx, dt, n, p0, r));
    CG_TRY(glb_op_apply_dot(op, Ap, p0, p0, 0, dots));
    CG_TRY(glb_norm2sq(ctx, dt, n, r, &rsq));
    CgState init{};
    init.rsq_old = rsq;
    init.rsq_new = rsq;
    init.pAp_re = dots[0];
    init.pAp_im = dots[1];
    init.bnorm = sqrt(bsq);
    init.eps = eps;
    init.iter = 0;
    init.max_iter = max_iter;
    init.done = (max_iter <= 0) ? 1 : 0;
    init.hit_max = 0;
    init.hist_cap = d_hist ? hist_cap : 0;
    h_st[0] = init;
    if (cudaMemcpyAsync(d_st, &h_st[0], sizeof(CgState), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "upload CgState");
      goto done;
    }
    rep->bnorm = init.bnorm;

    // --- the loop: enqueue batches, poll the previous batch's state while the next one runs
    const int BATCH = 8;
    void* pc = p0;
    void* pa = p1;
    int enq = 0;
    bool finished = (max_iter <= 0);
    bool have_pending = false;
    while (!finished) {
      for (int b = 0; b < BATCH; b++) {
        if (ctx->nranks > 1 && comm_p2p(ctx)) {
          CG_TRY(launch_cg_update(ctx, dt, d_st, d_hist, pc, d_x, Ap, r, n, 2));  // last block sums over ranks
        } else if (ctx->nranks > 1) {
          CG_TRY(launch_cg_update(ctx, dt, d_st, d_hist, pc, d_x, Ap, r, n, 1));
          CG_TRY(allreduce_device(ctx, d_st->partial, 1));
          CG_TRY(launch_cg_post_update(ctx, d_st, d_hist));
        } else {
          CG_TRY(launch_cg_update(ctx, dt, d_st, d_hist, pc, d_x, Ap, r, n, 0));
        }
        bool swapped = false;
        CG_TRY(direction_and_apply(op, d_st, r, pc, pa, Ap, &swapped));
        if (swapped) std::swap(pc, pa);
        enq++;
      }
      // state after this batch -> pinned slot (enq/BATCH)&1, asynchronously
      const int slot = (enq / BATCH) & 1;
      if (have_pending) {
        // the copy of the PREVIOUS batch was recorded on ev_a: wait for it now (this batch is
        // already queued behind it, so the GPU stays busy)
        if (cudaEventSynchronize(ctx->ev_a) != cudaSuccess) {
          rc = fail(GLB_ERR_CUDA, "cudaEventSynchronize");
          goto done;
        }
        if (h_st[slot ^ 1].done) finished = true;
      }
      if (cudaMemcpyAsync(&h_st[slot], d_st, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaEventRecord(ctx->ev_a, ctx->stream) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "state readback");
        goto done;
      }
      have_pending = true;
      if (finished) break;
      if (enq >= max_iter + BATCH) {  // everything that could run has been enqueued
        finished = true;
      }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "cudaStreamSynchronize");
      goto done;
    }
    if (cudaMemcpy(&h_st[0], d_st, sizeof(CgState), cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = fail(GLB_ERR_CUDA, "final state readback");
      goto done;
    }
    const CgState fin = h_st[0];
    rep->iterations = fin.iter;
    rep->ops = 2 + (fin.iter > 0 ? fin.iter - 1 : 0);  // one apply per iteration that did not stop
    rep->hit_max_iter = fin.hit_max;
    rep->rsq = fin.rsq_new;
    if (d_hist) {
      const int m = fin.iter < hist_cap ? fin.iter : hist_cap;
      if (m > 0 && cudaMemcpy(rsq_hist, d_hist, sizeof(double) * m, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = fail(GLB_ERR_CUDA, "history readback");
        goto done;
      }
    }
  }
done:
  if (d_hist) cudaFreeAsync(d_hist, ctx->stream);
  if (d_st) cudaFreeAsync(d_st, ctx->stream);
  glb_vec_free(ctx, Ap);
  glb_vec_free(ctx, p1);
  glb_vec_free(ctx, p0);
  glb_vec_free(ctx, r);
  return rc;
#undef CG_TRY
}

// largest relative deviation |predicted - exact| / exact of |r|^2 during the last single-kernel CG solve of this
// process (0 when the two-kernel loop ran); measurement aid for the parity tests
extern "C" double glb_cg_last_pred_err(void) { return glb::g_last_pred_err; }
