/* glb200.h -- C ABI of the B200-native solver hot path of generic-linalg.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ and no torch
 * types.  It sits one level BELOW the reference's public C++ API
 * (generic_inverters.h / inverter_struct.h, which return a std::string-bearing
 * struct by value and therefore cannot be a C ABI): the retained C++ solver
 * shells in generic-linalg_b200/host/ call ONLY the functions declared here.
 *
 * What each group replaces in the reference (paths relative to its root):
 *   vectors      : the `new T[size]` / `delete[]` work vectors of every solver
 *                  (e.g. generic_cg.cpp:292-294,370-372) -> device-resident.
 *   operators    : the `void (*matrix_vector)(T* lhs, T* rhs, void* extra_info)`
 *                  callbacks  square_laplacian (square_laplace.cpp:182,
 *                  imag_laplace.cpp:126), square_laplace (operators.cpp:28),
 *                  square_laplace_u1 (:73), square_staggered(_u1) (:127,:184),
 *                  gamma5 / dagger / normal variants (:242,:262,:316,:372,:444),
 *                  apply_stencil_2d (stencil_2d/coarse_stencil.cpp:12).
 *   BLAS-1       : generic_vector.h:12-169 (zero, copy, dot, norm2sq,
 *                  diffnorm2sq) and the open-coded axpy loops of each solver.
 *   fused        : one-pass versions of the loop bodies generic_cg.cpp:326-351,
 *                  generic_cr.cpp:249-285, generic_bicgstab.cpp:261-303,
 *                  generic_gcr.cpp:284-292, generic_cg_m.cpp:414-518.
 *   cg pipeline  : the whole CG loop generic_cg.cpp:324-354 run without any
 *                  host round trip (scalars and the stopping test on device).
 *
 * Conventions
 *   - every function returns GLB_OK (0) or a non-zero error code; the text of
 *     the last error is available from glb_last_error().  There is no CPU
 *     fallback anywhere: without a CUDA device glb_create() fails.
 *   - `dtype` is GLB_REAL (double) or GLB_COMPLEX (interleaved re,im doubles,
 *     layout-compatible with std::complex<double>).
 *   - vector lengths `n` are in ELEMENTS of that dtype (the reference's `size`).
 *   - device pointers are plain `void*` obtained from glb_vec_alloc.
 *   - complex scalars cross the boundary as `const double a[2]` = {re, im};
 *     for GLB_REAL only a[0] is used.
 *   - all work is enqueued on the context's stream; functions that return a
 *     scalar to the host synchronise that stream, the others do not.
 *   - dot products conjugate their FIRST argument (generic_vector.h:102).
 *   - element-wise updates evaluate exactly the reference's expression (same
 *     operation order, no fused multiply-add), so given equal scalars they are
 *     bit-identical to the CPU code; reductions use a fixed tree and are
 *     run-to-run reproducible.
 */
#ifndef GLB200_H
#define GLB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLB_OK 0
#define GLB_ERR_CUDA 1
#define GLB_ERR_ARG 2
#define GLB_ERR_COMM 3
#define GLB_ERR_STATE 4

enum glb_dtype { GLB_REAL = 0, GLB_COMPLEX = 1 };

typedef struct glb_context glb_context;
typedef struct glb_operator glb_operator;

/* ------------------------------------------------------------------ lifecycle */
int glb_create(int device, glb_context** ctx);
int glb_destroy(glb_context* ctx);
const char* glb_last_error(void);
int glb_synchronize(glb_context* ctx);
void* glb_stream(glb_context* ctx);          /* the cudaStream_t all work is enqueued on   */
int glb_device(glb_context* ctx);
int glb_sm_count(glb_context* ctx);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
unsigned long long glb_kernel_launches(void);
/* Per-kernel timing (measurement aid, bench.py roofline leg).  While enabled, every launch of a
   classified kernel is bracketed by two CUDA events on the context's stream.  Classes: 1 one-pass
   D^dag D with the fused CG direction update, 2 one-pass D^dag D, 3 CG x/r update, 4 staggered /
   gauged-Laplace stencil, 5 coarse stencil, 6 Laplace.  glb_prof_enable(ctx,1) clears and starts,
   (ctx,0) stops; glb_prof_read synchronises and returns the durations (ms) of class `cls` in launch
   order (at most cap of them; *n = how many were recorded). */
int glb_prof_enable(glb_context* ctx, int on);
int glb_prof_read(glb_context* ctx, int cls, int cap, float* ms, int* n);
/* classes 8-11: 8 streaming BLAS-1 kernels, 9 multi_dot (batched inner products of the GCR sweeps), 10 lincomb,
   11 multigrid prolong / restrict.  glb_prof_summary: number of launches of class `cls` since glb_prof_enable(ctx,1),
   their summed duration (ms) and their summed algorithmic bytes (tools/bench_mg.py: where a multigrid solve spends
   its time, GB/s per kernel class). */
int glb_prof_summary(glb_context* ctx, int cls, int* launches, double* ms_total, double* bytes_total);

/* ------------------------------------------------------ slab communicator (y-slabs) */
/* One process per GPU.  Rank g of G owns rows [g*Y/G, (g+1)*Y/G) of every lattice
 * vector.  Bootstrapping (exchange of the 128-byte id and of the IPC handles) is done
 * by the host program (torch.distributed in bench.py/tests; any launcher will do).   */
#define GLB_COMM_ID_BYTES 128
int glb_comm_unique_id(char id[GLB_COMM_ID_BYTES]);
int glb_comm_init(glb_context* ctx, int rank, int nranks, const char id[GLB_COMM_ID_BYTES]);
int glb_comm_rank(glb_context* ctx);
int glb_comm_size(glb_context* ctx);
/* NVLink peer-memory fast path for halos and reductions: glb_comm_init exports this rank's
 * arena (CUDA IPC), all-gathers the handles through NCCL and maps every peer; 1 if that worked on
 * every rank (else NCCL send/recv + allreduce are used).  GLB_P2P=0 in the environment disables it. */
int glb_comm_p2p_enabled(glb_context* ctx);
int glb_comm_barrier(glb_context* ctx);

/* ---------------------------------------------------------------------- vectors */
int glb_vec_alloc(glb_context* ctx, int dtype, size_t n, void** dptr);
int glb_vec_free(glb_context* ctx, void* dptr);
int glb_vec_upload(glb_context* ctx, int dtype, size_t n, void* dst_dev, const void* src_host);
int glb_vec_download(glb_context* ctx, int dtype, size_t n, void* dst_host, const void* src_dev);
int glb_vec_zero(glb_context* ctx, int dtype, size_t n, void* d);
int glb_vec_copy(glb_context* ctx, int dtype, size_t n, void* dst, const void* src);
/* pinned host staging buffers (used by the drop-in shells for host<->device copies) */
int glb_host_alloc(glb_context* ctx, size_t bytes, void** hptr);
int glb_host_free(glb_context* ctx, void* hptr);

/* -------------------------------------------------------------------- operators */
/* flags for glb_op_create_staggered */
#define GLB_STAG_DAGGER 1u   /* D^dagger : operators.cpp:372                                   */
#define GLB_STAG_GAMMA5 2u   /* gamma5 D : operators.cpp:262,316                               */
#define GLB_STAG_NORMAL 4u   /* D^dagger D through a temporary : operators.cpp:444             */
/* even/odd pieces (gauged only; not combinable with the flags above) */
#define GLB_STAG_DEO 8u      /* D_eo: hopping term on even sites, odd sites zeroed : operators.cpp:456   */
#define GLB_STAG_DOE 16u     /* D_oe: hopping term on odd sites, even sites zeroed : operators.cpp:494   */
#define GLB_STAG_M2MDEODOE 32u /* m^2 - D_eo D_oe on even sites, odd sites zeroed  : operators.cpp:549   */

/* 5-point periodic Laplacian, Nc colours per site, out = diag*in - sum of 4 neighbours.
 * diag = 4+m2 (square_laplace.cpp:182; operators.cpp:28), or 4+m2+i (imag_laplace.cpp:126).
 * The operator acts on the GLOBAL X x Y lattice; with a communicator each rank holds its slab. */
int glb_op_create_laplace(glb_context* ctx, int dtype, int X, int Y, int Nc, double diag_re, double diag_im,
                          glb_operator** op);
/* real free staggered operator, tests/multishift/multishift.cpp:677 (double; out = D_free in + m in) */
int glb_op_create_staggered_free_real(glb_context* ctx, int X, int Y, double mass, glb_operator** op);
/* gauged Laplacian, operators.cpp:73.  h_links: host array in the reference layout
 * lattice[y*X*2 + x*2 + mu] (complex), GLOBAL lattice; each rank uploads only its slab. */
int glb_op_create_laplace_u1(glb_context* ctx, const void* h_links, int X, int Y, double mass, glb_operator** op);
/* 2-D staggered operator; h_links == NULL gives the free operator (operators.cpp:127,262). */
int glb_op_create_staggered(glb_context* ctx, const void* h_links, int X, int Y, double mass, unsigned flags,
                            glb_operator** op);
/* Same, but the host array holds only THIS RANK's rows plus two ghost rows on each side:
 * rows y0-2, y0-1, y0, ..., y0+Yloc-1, y0+Yloc, y0+Yloc+1 (periodic), i.e. (Yloc+4)*X*2 complex
 * numbers.  Lets a multi-GPU job build its slabs without any rank ever holding the global field. */
int glb_op_create_staggered_local(glb_context* ctx, const void* h_links_local, int X, int Y, double mass,
                                  unsigned flags, glb_operator** op);
/* rows [y0, y0+Yloc) this rank owns of a Y-row lattice */
int glb_slab_bounds(glb_context* ctx, int Y, int* y0, int* Yloc);
/* gamma_5 alone: out = (-1)^(x+y) in  (operators.cpp:242) */
int glb_op_create_gamma5(glb_context* ctx, int X, int Y, glb_operator** op);
/* data-driven stencil, stencil_2d/coarse_stencil.h:33 + coarse_stencil.cpp:29-172 (DIR_ALL).
 * clover: nc*nc*V, hopping: 4 planes, two_link: 8 planes or NULL; host arrays, reference layout. */
int glb_op_create_stencil2d(glb_context* ctx, const void* clover, const void* hopping, const void* two_link, int X,
                            int Y, int nc, const double shift[2], const double eo_shift[2],
                            const double dof_shift[2], glb_operator** op);
int glb_op_destroy(glb_operator* op);
int glb_op_set_mass(glb_operator* op, double mass);
/* stencil2d operators: the three diagonal shifts of stencil_2d (coarse_stencil.h:62-71: shift, eo_shift, dof_shift;
 * applied by coarse_stencil.cpp:153-169).  The reference's set-up changes them in place between the null-vector
 * generation and the final build (aa_mg_square_staggered_u1.cpp:757, :933, :996, :1086); NULL leaves one unchanged. */
int glb_op_set_shifts(glb_operator* op, const double shift[2], const double eo_shift[2], const double dof_shift[2]);
int glb_op_get_shifts(const glb_operator* op, double shift[2], double eo_shift[2], double dof_shift[2]);
/* stencil2d operators: copy the matrices back to host arrays in the reference layout (clover nc*nc*V, hopping
 * 4*nc*nc*V complex; either may be NULL) -- what stencil_2d::clover / ::hopping hold.  On y-slabs V is this rank's
 * X*Yloc sites: every plane holds the local rows only. */
int glb_op_stencil_download(glb_operator* op, void* h_clover, void* h_hopping);
int glb_op_dtype(const glb_operator* op);
size_t glb_op_local_size(const glb_operator* op);   /* elements held by this rank              */
size_t glb_op_global_size(const glb_operator* op);
glb_context* glb_op_context(const glb_operator* op);
/* how many operator applications one glb_op_apply counts for (1; the reference counts the
 * normal operator as one callback too) -- kept for ops_count parity */
/* out = A in.  out is fully overwritten and must not alias in (same contract as the callback). */
int glb_op_apply(glb_operator* op, void* d_out, const void* d_in);
/* out = A in, and in the same pass  dots[0..1] = <w,out> (w may be d_in or any vector; NULL skips),
 * dots[2] = |out|^2 if want_norm.  Host-synchronous. */
int glb_op_apply_dot(glb_operator* op, void* d_out, const void* d_in, const void* d_w, int want_norm,
                     double dots[3]);
/* algorithmic HBM bytes one apply moves (SURVEY section 8 d-bytes); used by bench.py */
double glb_op_bytes_per_apply(const glb_operator* op);

/* ----------------------------------------------------------------------- BLAS-1 */
int glb_dot(glb_context* ctx, int dtype, size_t n, const void* x, const void* y, double out[2]);
int glb_norm2sq(glb_context* ctx, int dtype, size_t n, const void* x, double* out);
int glb_diffnorm2sq(glb_context* ctx, int dtype, size_t n, const void* x, const void* y, double* out);
/* out[0..1] = <x,y>, out[2] = |x|^2 in one pass (GCR alpha, generic_gcr.cpp:255) */
int glb_dot_norm(glb_context* ctx, int dtype, size_t n, const void* x, const void* y, double out[3]);
/* out[2i..2i+1] = <X[i], y>, i < k, one pass over y (GCR / Gram-Schmidt sweeps) */
int glb_multi_dot(glb_context* ctx, int dtype, size_t n, int k, const void* const* X, const void* y, double* out);

int glb_sub(glb_context* ctx, int dtype, size_t n, const void* a, const void* b, void* out);        /* out = a - b   */
int glb_add(glb_context* ctx, int dtype, size_t n, const void* a, const void* b, void* out);        /* out = a + b   */
int glb_axpy(glb_context* ctx, int dtype, size_t n, const double a[2], const void* x, void* y);     /* y = y + a*x   */
int glb_xpay(glb_context* ctx, int dtype, size_t n, const void* x, const double a[2], void* y);     /* y = x + a*y   */
int glb_axpyz(glb_context* ctx, int dtype, size_t n, const double a[2], const void* x, const void* y,
              void* z);                                                                              /* z = y + a*x   */
int glb_rdiv(glb_context* ctx, int dtype, size_t n, const void* x, double d, void* out);            /* out = x / d   */
/* y = y + a*x and |y|^2 in the same pass (generic_cg_m.cpp:423-429) */
int glb_axpy_norm(glb_context* ctx, int dtype, size_t n, const double a[2], const void* x, void* y, double* nrm);
/* out = (init ? init : 0) + c[0]*X[0] + c[1]*X[1] + ... accumulated in that order
 * (generic_gcr.cpp:284-292 with init==out; generic_gmres.cpp:632-648) */
int glb_lincomb(glb_context* ctx, int dtype, size_t n, int k, const double* coef, const void* const* X,
                const void* init, void* out);

/* ------------------------------------------------------- fused solver loop bodies */
/* x = x + a*p ; r = r + b*q ; *rsq = |r|^2   (generic_cg.cpp:328-335 with b = -a, q = Ap) */
int glb_update_xr_norm(glb_context* ctx, int dtype, size_t n, const double a[2], const void* p, void* x,
                       const double b[2], const void* q, void* r, double* rsq);
/* p = r + beta*p ; Ap = Ar + beta*Ap ; *apsq = |Ap|^2   (generic_cr.cpp:279-285) */
int glb_update_p_ap_norm(glb_context* ctx, int dtype, size_t n, const void* r, const void* Ar, const double beta[2],
                         void* p, void* Ap, double* apsq);
/* x = x + alpha*p + omega*s ; r = s - omega*As ; out = {|r|^2, <r0,r>.re, <r0,r>.im}
 * (generic_bicgstab.cpp:274-295) */
int glb_bicgstab_update(glb_context* ctx, int dtype, size_t n, const double alpha[2], const void* p,
                        const double omega[2], const void* s, const void* As, const void* r0, void* x, void* r,
                        double out[3]);
/* p = r + beta*(p - omega*Ap)   (generic_bicgstab.cpp:300-303) */
int glb_bicgstab_pupdate(glb_context* ctx, int dtype, size_t n, const void* r, const double beta[2],
                         const double omega[2], const void* Ap, void* p);
/* out = s*x with a real s (generic_vector.h:173-203 normalize: v *= 1/sqrt(|v|^2)) */
int glb_rscale(glb_context* ctx, int dtype, size_t n, const void* x, double s, void* out);
/* out = conj(x) (complex) / out = x (real)   (generic_vector.h conj<>; generic_bicgstab_m.cpp:494) */
int glb_conj(glb_context* ctx, int dtype, size_t n, const void* x, void* out);
/* multishift BiCGStab, one shift: s_n = c0*r + c1*(s_n - c2*(c3*w - c4*r_prev)) with c = {c0..c4} complex pairs
 * (generic_bicgstab_m.cpp:705: c0 = zeta rho, c1 = alpha_n, c2 = chi_n/beta_n, c3 = zeta rho_prev, c4 = zeta_prev rho_prev) */
int glb_bicgstabm_update_s(glb_context* ctx, int dtype, size_t n, const double c[10], const void* r, const void* w,
                           const void* r_prev, void* s_n);
/* multishift: for s < ns : x[s] = x[s] - beta_s[s]*p_s[s]   (generic_cg_m.cpp:414-417) */
int glb_cgm_update_x(glb_context* ctx, int dtype, size_t n, int ns, const double* beta_s, const void* const* p_s,
                     void* const* x);
/* multishift: for s < ns : p_s[s] = zeta[s]*r + alpha_s[s]*p_s[s]  (generic_cg_m.cpp:501-512), r read once */
int glb_cgm_update_p(glb_context* ctx, int dtype, size_t n, int ns, const double* zeta, const double* alpha_s,
                     const void* r, void* const* p_s);

/* ----------------------------------------------- device-resident CG (no host round trip) */
/* Runs the loop of minv_vector_cg (generic_cg.cpp:159-206 / :307-354) entirely on the device:
 * alpha, beta and the stopping test `sqrt(rsqNew) < eps*bnorm || k == max_iter-1` are
 * evaluated by the kernels themselves; the host only enqueues batches and polls a flag.
 * Fusion (bytes/site, complex):  [p = r + beta p . Ap = A p . <p,Ap>] + [x,r update . |r|^2].
 * x is in/out (initial guess), b the rhs.  On return iters/ops follow the reference's counting
 * (generic_cg.cpp:362,366) EXCLUDING the final true-residual apply, which the shell performs.
 * rsq_hist (host, may be NULL) receives |r|^2 after each iteration, hist_cap entries at most. */
typedef struct glb_cg_report {
  int iterations;      /* value of k+1 at loop exit (what the reference reports as `iter`) */
  int ops;             /* operator applications performed (initial two + one per continued iteration) */
  int hit_max_iter;    /* loop ended because k == max_iter-1 */
  double rsq;          /* last recurrence |r|^2 */
  double bnorm;        /* sqrt(|b|^2) */
} glb_cg_report;
/* 1 if glb_cg_solve can run this operator on this context (always on one rank; on slabs the
 * one-pass D^dag D staggered operator), else the host-scalar shell must be used */
int glb_cg_solve_supported(const glb_operator* op);
int glb_cg_solve(glb_operator* op, void* d_x, const void* d_b, int max_iter, double eps, glb_cg_report* rep,
                 double* rsq_hist, int hist_cap);
/* On the one-pass staggered D^dag D operator (complex, gauged, even X; slabs over peer memory) glb_cg_solve runs
 * the whole iteration as ONE kernel with ONE rank-wide reduction (cgstep.cu): r, x and p updates, the operator and
 * the four inner products |r|^2, <p,Ap>, <r,Ap>, |Ap|^2 in a single pass (160 B/site instead of 192).  beta's
 * numerator |r_new|^2 is predicted from those sums one step ahead (|r - a q|^2 = |r|^2 - 2 Re(a <r,q>) + |a|^2 |q|^2)
 * and replaced by the exact sum for alpha and for the stopping test of generic_cg.cpp:339.  This returns the
 * largest relative |predicted - exact| / exact seen in the last such solve (0 if the two-kernel loop ran);
 * GLB_CGSTEP=0 in the environment selects the two-kernel loop. */
double glb_cg_last_pred_err(void);
/* measurement / test aid: on = 0 makes glb_cg_solve use the two-kernel loop on that operator too, on = 1 (default)
 * the single-kernel iteration; variant > 0 selects schedule and kernel shape (1000*rows per work item [0 = static partition] + 100*consumer
 * warps + 10*stages + blocks per SM).
 * Returns the previous on/off value. */
int glb_cg_step_mode(int on, int variant);

/* ----------------------------------------------- device-resident BiCGStab and CR (csrc/krylov.cu)
 * The loops of minv_vector_bicgstab (generic_bicgstab.cpp:258-308 complex, :75-125 real) and minv_vector_cr
 * (generic_cr.cpp:246-286, :76-116) with alpha / omega / beta and the stopping test kept on the GPU: every kernel
 * forms the scalars it needs from the grid totals its predecessor on the stream left on the device, the host only
 * polls the state once per batch of iterations (later batches are replayed as one CUDA graph).  Same vector kernels,
 * launch geometry and reduction trees as the host-scalar shells: the iterates are bit-identical to theirs.
 * x is in/out (initial guess), b the rhs; the report counts as the reference does, EXCLUDING the final
 * true-residual apply (the shell performs it); rsq_hist (host, may be NULL) receives |r|^2 after each iteration.
 * One rank, any native operator with fused reductions (not gamma5, not the composite stencil views). */
#define GLB_KRYLOV_BICGSTAB 1
#define GLB_KRYLOV_CR 2
int glb_krylov_solve_supported(const glb_operator* op, int alg);
int glb_krylov_solve(glb_operator* op, int alg, void* d_x, const void* d_b, int max_iter, double eps,
                     glb_cg_report* rep, double* rsq_hist, int hist_cap);
/* measurement / test aid: on = 1 (default; GLB_KRYLOV_GRAPH in the environment) replays batches as a CUDA graph,
 * 0 launches every kernel directly, < 0 only queries.  Returns the previous setting.  glb_krylov_last_used_graph:
 * 1 if the last glb_krylov_solve of this process replayed a graph. */
int glb_krylov_graph_mode(int on);
int glb_krylov_last_used_graph(void);

/* ----------------------------------------------- measurement aids of the peer-memory communicator */
/* tools/p2p_bench.py.  Collective over the communicator's ranks (every rank calls with equal arguments).
 * glb_dbg_p2p_bench: iters x [busy-wait busy_us, then sum 6 doubles over ranks with protocol `variant`]; wait_us[i] =
 * time rank-locally spent in sum i (variant 70 = the protocol of the library: one warp per peer; 0 = one warp for all
 * peers; 80 = recursive doubling; 100*k + v = only the first k ranks take part).
 * glb_dbg_p2p_pingpong: one 8-byte word each way between rank 0 and `peer`; rtt_us[i] on rank 0. */
int glb_dbg_p2p_bench(glb_context* ctx, int variant, int iters, double busy_us, float* wait_us);
int glb_dbg_p2p_pingpong(glb_context* ctx, int peer, int iters, float* rtt_us);

/* ----------------------------------------------- partial stencil applies */
/* apply_stencil_2d_eo / _oe / _tb / _bt (coarse_stencil.cpp:395, 560, 725, 1120; DIR_ALL) on a stencil2d operator:
 *   EO, OE: hopping term only, output on even / odd sites, the other parity zeroed;
 *   TB, BT: clover + hopping (+ two-link) from the bottom to the top half of the colour index (rows < nc/2 summed
 *           over c >= nc/2) or the mirror, the other rows zeroed.  No shifts are applied. */
#define GLB_PART_EO 1
#define GLB_PART_OE 2
#define GLB_PART_TB 3
#define GLB_PART_BT 4
int glb_op_apply_part(glb_operator* op, void* d_out, const void* d_in, int part);

/* ----------------------------------------------- preconditioned stencil paths */
/* The composite operators the reference builds from the partial applies of a stencil_2d, as device operators that
 * plug into every solver (glb_op_apply; no fused reductions, several passes through a temporary):
 *   M2MDEODOE  m^2 - D_eo D_oe on even sites, 0 on odd     apply_square_staggered_m2mdeodoe_stencil  operators_stencil.cpp:196
 *   M2MDTBDBT  m^2 - D_tb D_bt on the top colours, 0 below apply_square_staggered_m2mdtbdbt_stencil  mg_complex.cpp:1228
 *   NORMAL_EO  m^2 - D_eo D_oe - D_oe D_eo                 apply_square_staggered_normal_eo_stencil  mg_complex.cpp:1278
 *   NORMAL_TB  m^2 - D_tb D_bt - D_bt D_tb                 apply_square_staggered_normal_tb_stencil  mg_complex.cpp:1306
 *   DAGGER_EO  epsilon(x) D epsilon(x)                      apply_square_staggered_dagger_eo_stencil  mg_complex.cpp:1336
 *   DAGGER_TB  sigma_3 D sigma_3                            apply_square_staggered_dagger_tb_stencil  mg_complex.cpp:1355
 * (m = the stencil's `shift`).  A view shares the matrices and shifts of `base` (read at apply time; set_shifts on
 * either is seen by both) and must not outlive it, unless adopt_base != 0: then destroying the view destroys the base. */
#define GLB_SV_M2MDEODOE 1
#define GLB_SV_M2MDTBDBT 2
#define GLB_SV_NORMAL_EO 3
#define GLB_SV_NORMAL_TB 4
#define GLB_SV_DAGGER_EO 5
#define GLB_SV_DAGGER_TB 6
int glb_op_create_stencil_view(glb_operator* base, int kind, int adopt_base, glb_operator** view);
/* apply_square_staggered_{eo,tb}prec_prepare_stencil (operators_stencil.cpp:179, mg_complex.cpp:1211):
 * rhs_part = shift*rhs_orig - D_eo rhs_orig on even sites (top_bottom = 0) / - D_tb rhs_orig on the top colours (1),
 * 0 on the other half.  One pass. */
int glb_stencil_prec_prepare(glb_operator* op, int top_bottom, void* d_rhs_part, const void* d_rhs_orig);
/* apply_square_staggered_{eo,tb}prec_reconstruct_stencil (operators_stencil.cpp:217, mg_complex.cpp:1252):
 * lhs_full = lhs_part on the solved half, (rhs_other - D_oe lhs_part)/Re(shift) on the other half.  One pass. */
int glb_stencil_prec_reconstruct(glb_operator* op, int top_bottom, void* d_lhs_full, const void* d_lhs_part,
                                 const void* d_rhs_other);

/* ----------------------------------------------- even/odd preconditioning */
/* square_staggered_eoprec_prepare (operators.cpp:528-545): rhs_e = m rhs_orig - D_eo rhs_orig on even sites, 0 on
 * odd sites.  `op` is any gauged staggered operator (its links and mass are used). */
int glb_stag_eoprec_prepare(glb_operator* op, void* d_rhs_e, const void* d_rhs_orig);
/* square_staggered_eoprec_reconstruct (operators.cpp:574-598): lhs_full = lhs_e on even sites,
 * (rhs_o - D_oe lhs_e)/m on odd sites. */
int glb_stag_eoprec_reconstruct(glb_operator* op, void* d_lhs_full, const void* d_lhs_e, const void* d_rhs_o);

/* ----------------------------------------------- multigrid grid transfers */
/* prolong / restrict of multigrid/aa_mg/mg_complex.cpp:372-467 on device vectors.  A transfer
 * couples a fine level (Xf x Yf sites, dof_f values per site, vector index (y*Xf + x)*dof_f + d) to
 * the coarse level of bx x by blocks with nvec values per coarse site (index (yc*Xc + xc)*nvec + v).
 * null_vectors: nvec HOST pointers to complex arrays of Xf*Yf*dof_f elements each -- the reference's
 * mg_operator_struct_complex::null_vectors[level] (mg_complex.h:164) -- copied to the device once.
 * On y-slabs pass the local extent (the transfers are block-local; Yf must be a multiple of by).
 *   prolong : fine   = P coarse            (fine fully overwritten;  mg_complex.cpp:372-418)
 *   restrict: coarse = P^dagger fine       (coarse fully overwritten; mg_complex.cpp:422-467)   */
typedef struct glb_mg_transfer glb_mg_transfer;
int glb_mg_transfer_create(glb_context* ctx, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                           const void* const* null_vectors, glb_mg_transfer** out);
int glb_mg_transfer_destroy(glb_mg_transfer* t);
/* ---- set-up on the device ---- */
/* the transfer from DEVICE-resident null vectors (nvec device pointers, reference layout null_vectors[v][f]) */
int glb_mg_transfer_create_dev(glb_context* ctx, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                               const void* const* d_null_vectors, glb_mg_transfer** out);
/* block_orthonormalize + block_normalize (mg_complex.cpp:191-370) in place on nvec device vectors */
int glb_mg_block_orthonormalize(glb_context* ctx, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                                void* const* d_null_vectors);
/* BLOCK_EO partition of one null vector on the X x Y lattice (Y GLOBAL; on y-slabs the arrays hold this rank's rows,
 * and site parity counts global rows) with dof values per site: its "odd" part moves to odd_out (only those elements
 * of odd_out are written) and is zeroed in even_io.  colour_period = 0: odd SITES (x+y odd; top level,
 * null_partition_staggered, null_gen.cpp:26-35).  colour_period = m > 0: elements with (index % m) >= m/2
 * (null_partition_coarse, null_gen.cpp:109-126, where m = n_vectors[curr_level]). */
int glb_mg_partition(glb_context* ctx, int X, int Y, int dof, int colour_period, void* d_even_io, void* d_odd_out);
/* BLOCK_CORNER partition (null_gen.cpp:74-88, :132-152): the elements of corner class `which` (1..3) move to dst_out
 * and are zeroed in src_io.  colour_period = 0: sites by (x odd, y odd) = 1, (x odd, y even) = 2, (x even, y odd) = 3;
 * colour_period = p > 0: index % p in [p/4, 2p/4) = 1, [2p/4, 3p/4) = 2, [3p/4, p) = 3 (integer divisions). */
int glb_mg_partition_corner(glb_context* ctx, int X, int Y, int dof, int colour_period, int which, void* d_src_io,
                            void* d_dst_out);
/* Galerkin coarse operator P^dag A P of a five-point stencil2d fine operator (what generate_coarse_from_fine_stencil,
 * mg_complex.cpp:827-1026, assembles by probing with 1 + 8 applies per coarse colour): a new stencil2d operator on
 * the coarse lattice of the transfer, nc = nvec, all three coarse shifts zero.  ignore_shifts = 0: the fine shifts
 * are folded into the coarse clover; != 0: they are left out (the caller then sets the coarse shifts with
 * glb_op_set_shifts, aa_mg_square_staggered_u1.cpp:1080-1093).  The coarse lattice needs an even number (>= 2) of
 * sites per direction, as the reference's even/odd probing does.  On y-slabs (slab boundaries on block boundaries)
 * the neighbours' boundary rows of the null vectors travel once through the fine operator's halo path and the
 * result is the slab of the coarse operator. */
int glb_mg_galerkin(glb_mg_transfer* t, glb_operator* fine, int ignore_shifts, glb_operator** coarse);
size_t glb_mg_fine_size(const glb_mg_transfer* t);
size_t glb_mg_coarse_size(const glb_mg_transfer* t);
int glb_mg_prolong(glb_mg_transfer* t, void* d_fine, const void* d_coarse);
int glb_mg_restrict(glb_mg_transfer* t, void* d_coarse, const void* d_fine);

#ifdef __cplusplus
}
#endif
#endif /* GLB200_H */
