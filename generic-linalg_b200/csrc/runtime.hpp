// runtime.hpp -- host-side state behind the opaque handles of include/glb200.h.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/glb200.h"
#include "common.cuh"
#include "p2p.cuh"

namespace glb {

// last error text (thread local; glb_last_error())
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<unsigned long long> g_launches;

#define GLB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return glb::fail(GLB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)

#define GLB_LAUNCH_CHECK()                                                                          \
  do {                                                                                              \
    glb::g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) return glb::fail(GLB_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)

struct Comm;  // comm.cu

// per-kernel timing for bench.py's roofline leg (glb_prof_*): while enabled, the launchers of the
// classified kernels bracket each launch with a pair of CUDA events on the context's stream
enum ProfClass {
  PROF_NORMAL_FUSED = 1,  // normal_kernel with the fused CG direction update (96 B/site)
  PROF_NORMAL = 2,        // normal_kernel, plain D^dag D (64 B/site)
  PROF_CG_UPDATE = 3,     // cg_update_kernel (96 B/site)
  PROF_STAG = 4,          // stag_kernel, any flavour
  PROF_COARSE = 5,        // coarse_kernel / coarse_ring_kernel
  PROF_LAPLACE = 6,
  PROF_CG_STEP = 7,        // cg_step_kernel: the whole CG iteration in one pass (160 B/site)
  PROF_EW = 8,             // streaming BLAS-1 kernels (ew_kernel functors of blas1.cu, ews_kernel of krylov.cu)
  PROF_MULTI_DOT = 9,      // multi_dot_kernel: <X_i, y> for up to 16 stored vectors in one pass (GCR / VPGCR sweeps)
  PROF_LINCOMB = 10,       // lincomb_kernel: out = init + sum_i c_i X_i
  PROF_MG_TRANSFER = 11,   // mg_prolong_kernel / mg_restrict_kernel
  PROF_NCLASS = 12
};
struct ProfRec {
  cudaEvent_t a, b;
  int cls;
  double bytes;  // algorithmic bytes of the launch (every distinct element read once / written once), 0 if not known
};

}  // namespace glb

struct glb_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  // reduction workspace
  glb::ReduceWs red{};
  double* result_host_ptr = nullptr;  // host address of the mapped result buffer
  // small device scratch for pointer/coefficient tables of the multi-vector kernels
  void* d_table = nullptr;
  void* h_table = nullptr;  // pinned
  size_t table_bytes = 0;
  // slab communicator
  int rank = 0, nranks = 1;
  glb::Comm* comm = nullptr;
  // kernel timing (off by default)
  bool prof_on = false;
  std::vector<glb::ProfRec> prof;
};

namespace glb {

constexpr int LINK_GHOST = 2;

enum OpKind {
  OPK_LAPLACE = 0,      // free Laplacian, Nc colours, real or complex
  OPK_LAPLACE_U1 = 1,   // gauged Laplacian
  OPK_STAGGERED = 2,    // staggered family (free or gauged; dagger / gamma5 / normal flags)
  OPK_GAMMA5 = 3,
  OPK_STENCIL = 4
};

}  // namespace glb

struct glb_operator {
  glb_context* ctx = nullptr;
  int kind = 0;
  int dtype = GLB_COMPLEX;
  int X = 0, Y = 0;          // global lattice
  int y0 = 0, Yloc = 0;      // this rank's slab: rows [y0, y0+Yloc)
  int nc = 1;
  double mass = 0.0;
  double diag_re = 0.0, diag_im = 0.0;
  unsigned flags = 0;
  bool has_links = false;
  // links as two site-major planes (SoA), slab-local, plus the U_y row below the slab
  // Ux/Uy point at row 0 of the slab inside *_store, which carries LINK_GHOST periodic ghost rows
  // below and above: U(x,y) is addressable for y in [-LINK_GHOST, Yloc+LINK_GHOST)
  glb::cplx* Ux = nullptr;
  glb::cplx* Uy = nullptr;
  glb::cplx* Uy_lo = nullptr;  // = Uy - X (row y0-1)
  glb::cplx* Ux_store = nullptr;
  glb::cplx* Uy_store = nullptr;
  // temporaries
  void* tmp = nullptr;         // for the normal operator
  // ghost rows for slab decomposition (rows y0-1 and y0+Yloc of the input), nc*X elements each
  void* ghost_lo = nullptr;    // ghost_depth rows: y0-ghost_depth .. y0-1 (lowest first)
  void* ghost_hi = nullptr;    // ghost_depth rows: y0+Yloc .. y0+Yloc+ghost_depth-1
  int ghost_depth = 0;
  bool ghost_p2p = false;      // ghost rows live in the peer-mapped arena (NVLink stores + flags)
  size_t ghost_off = 0;        // offset of this operator's ghost area inside every rank's arena
  size_t ghost_arena_bytes = 0;
  unsigned long long halo_seq = 0;
  void* send_lo = nullptr;     // staging for boundary rows produced on the fly (device CG)
  void* send_hi = nullptr;
  // single-kernel CG iteration on slabs (cgstep.cu): ghost rows of (r, q, p), two parities, in the peer arena
  bool cs_ready = false;
  size_t cs_off = 0;           // [parity 0: lo | hi][parity 1: lo | hi][flag_lo, flag_hi, count0, count1]
  unsigned long long cs_seq = 0;
  // stencil data (slab-local, device)
  glb::cplx* clover = nullptr;
  glb::cplx* hopping = nullptr;   // 4 planes of nc*nc*Vloc
  glb::cplx* two_link = nullptr;  // 8 planes
  bool has_two = false;
  double shift[2] = {0, 0}, eo_shift[2] = {0, 0}, dof_shift[2] = {0, 0};
  // composite views of a stencil2d operator (GLB_SV_*): matrices and shifts are the base's, read at apply time
  int composite = 0;
  glb_operator* base = nullptr;
  bool owns_base = false;
  void* tmp2 = nullptr;
  // launch geometry chosen at creation
  int stencil_blocks = 0;
};

namespace glb {

// RAII bracket around one classified launch; a no-op unless glb_prof_enable(ctx, 1) was called
struct ProfScope {
  glb_context* ctx;
  cudaEvent_t b = nullptr;
  ProfScope(glb_context* c, int cls, double bytes = 0.0) : ctx(c) {
    if (!c->prof_on) return;
    ProfRec r{};
    r.cls = cls;
    r.bytes = bytes;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, c->stream);
    b = r.b;
    c->prof.push_back(r);
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, ctx->stream);
  }
};

inline size_t elem_bytes(int dtype) { return dtype == GLB_COMPLEX ? 16 : 8; }

// blas1.cu
int blas_grid(const glb_context* ctx, size_t n, int threads, int per_thread);

// stencil.cu : launchers.  `w` (optional) is dotted against the output, want_norm adds |out|^2.
struct ApplyFusion {
  // pre-op: in = r + beta * p_old, written to p_new (device-resident CG); beta read from device
  const void* r = nullptr;
  const void* p_old = nullptr;
  void* p_new = nullptr;
  const double* cg_state = nullptr;  // device CgState (see cg.cu), null if unused
  // post-op reductions
  const void* w = nullptr;   // <w, out>
  bool w_is_input = false;   // w == the (possibly fused) input vector
  bool want_norm = false;    // |out|^2
  bool to_host = false;      // copy results to the mapped host buffer
  int cg_role = 0;           // 0 none, 1: epilogue stores <p,Ap> into CgState, 2: rank-local partial (summed on the
                             // stream next), 3: the last block sums over ranks itself (peer memory) and publishes
  P2PRed pr{};               // cg_role 3: descriptor of that reduction
  HaloWait wait{};           // slabs on the peer-memory path: ghost-row flags to wait for in the prologue
};
int launch_staggered(glb_operator* op, void* out, const void* in, bool dagger, const ApplyFusion& f);
// even/odd pieces of the staggered operator (operators.cpp:456-616): parity 0 updates even sites, 1 odd sites;
// post 0: h/2 | 0, post 1: coef*aux - h/2 | 0, post 2: coef*(aux - h/2) | in
int launch_staggered_eo(glb_operator* op, void* out, const void* in, int parity, int post, double coef, const void* aux,
                        const ApplyFusion& f);
int launch_laplace(glb_operator* op, void* out, const void* in, const ApplyFusion& f);
int launch_gamma5(glb_operator* op, void* out, const void* in);
// normal.cu : D^dag D in one pass (single rank, gauged, even X)
bool normal_fused_ok(const glb_operator* op);
int launch_normal(glb_operator* op, void* out, const void* in, const ApplyFusion& f);
int launch_stencil2d(glb_operator* op, void* out, const void* in, const ApplyFusion& f);
// GLB_PART_* with a fused post-operation (see coarse_part_kernel)
int launch_stencil2d_part(glb_operator* op, void* out, const void* in, int part, int post = 0, const double coef[2] = nullptr,
                          const void* aux = nullptr);
int launch_stencil2d_sign(glb_operator* op, void* out, const void* in, int mode);  // 0: epsilon(x), 1: sigma_3

// ops.cu : stencil2d operator around device-resident matrices (ownership passes to the operator)
int op_adopt_stencil2d(glb_context* ctx, int X, int Y, int Yloc, int nc, cplx* d_clover, cplx* d_hopping, glb_operator** out);

// comm.cu : fills op->ghost_lo / ghost_hi from the neighbouring ranks' boundary rows of `in`
// defer != nullptr (peer-memory path): no wait kernel is launched; *defer receives the flags and the consuming kernel
// waits itself where it reads a ghost row (seq == 0 when there is nothing to wait for: one rank, NCCL transport)
int halo_exchange(glb_operator* op, const void* in, int nrows, HaloWait* defer = nullptr);
// same, boundary rows taken from explicit buffers (nrows lowest rows in send_lo, nrows highest in send_hi)
int halo_exchange_ptrs(glb_operator* op, const void* send_lo, const void* send_hi, int nrows, HaloWait* defer = nullptr);
int allreduce_device(glb_context* ctx, double* d_vals, int n);
int allreduce_sum(glb_context* ctx, double* host_vals, int n);
void comm_destroy(glb_context* ctx);
bool comm_p2p(const glb_context* ctx);
char* comm_peer(glb_context* ctx, int g);        // rank g's arena as mapped here (peer-memory path only)
long long comm_spin_budget(const glb_context* ctx);
unsigned int* comm_ticket(glb_context* ctx);     // self-resetting block counter for small push kernels
P2PRed comm_p2p_red(glb_context* ctx);
P2PRed comm_p2p_red_range(glb_context* ctx, unsigned long long count);
struct HaloTargets {
  char* dst_down_hi;
  char* dst_up_lo;
  unsigned long long* flag_down_hi;
  unsigned long long* flag_up_lo;
  HaloWait wait;
  unsigned int* ticket;
  size_t bytes;
};
int halo_p2p_begin(glb_operator* op, int nrows, HaloTargets* t);
void* comm_arena_alloc(glb_context* ctx, size_t bytes, size_t* offset, unsigned long long* seq_start);
void comm_arena_free(glb_context* ctx, size_t offset, size_t bytes, unsigned long long seq);

}  // namespace glb
