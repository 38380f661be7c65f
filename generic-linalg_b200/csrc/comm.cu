// comm.cu -- y-slab communicator: one process per GPU, ring of slabs (periodic lattice).
//
//  * halo exchange: the first / last `rows` of a slab-local vector go to the neighbouring ranks'
//    ghost buffers (2 x X x nc x 16 B per apply: 128 KiB per direction at 8192).
//  * reductions: k doubles summed over ranks.
// NCCL is bound lazily (dlopen) so that a single-GPU process never needs it; when torch is loaded
// first its bundled libnccl.so.2 is the one that gets used, otherwise the system one.
#include <dlfcn.h>

#include <cstring>

#include "runtime.hpp"

namespace glb {

// minimal NCCL surface (nccl.h: ncclUniqueId is 128 bytes; ncclFloat64 = 8; ncclSum = 0)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId_t;
enum { NCCL_SUCCESS = 0, NCCL_SUM = 0, NCCL_CHAR = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return GLB_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(GLB_ERR_COMM, std::string("cannot load libnccl: ") + dlerror());
#define BIND(field, sym)                                                      \
  *(void**)(&g_nccl.field) = dlsym(lib, sym);                                 \
  if (!g_nccl.field) return fail(GLB_ERR_COMM, std::string("libnccl lacks ") + sym);
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllReduce, "ncclAllReduce")
  BIND(Send, "ncclSend")
  BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_nccl.lib = lib;
  return GLB_OK;
}

#define GLB_NCCL(expr)                                                                                      \
  do {                                                                                                      \
    int _r = (expr);                                                                                        \
    if (_r != NCCL_SUCCESS) return fail(GLB_ERR_COMM, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

struct Comm {
  ncclComm_t nccl = nullptr;
  double* d_red = nullptr;  // device staging for host-value reductions
  double* h_red = nullptr;  // pinned
};

void comm_destroy(glb_context* ctx) {
  if (!ctx->comm) return;
  if (ctx->comm->nccl) g_nccl.CommDestroy(ctx->comm->nccl);
  cudaFree(ctx->comm->d_red);
  cudaFreeHost(ctx->comm->h_red);
  delete ctx->comm;
  ctx->comm = nullptr;
}

int halo_exchange_ptrs(glb_operator* op, const void* send_lo, const void* send_hi, int nrows) {
  glb_context* ctx = op->ctx;
  if (ctx->nranks == 1) return GLB_OK;
  if (!ctx->comm) return fail(GLB_ERR_STATE, "slab operator used before glb_comm_init");
  if (nrows < 1 || nrows > op->ghost_depth) return fail(GLB_ERR_ARG, "halo deeper than the operator's ghost rows");
  const int G = ctx->nranks, g = ctx->rank;
  const int up = (g + 1) % G, down = (g + G - 1) % G;
  const size_t rowb = (size_t)op->X * op->nc * elem_bytes(op->dtype);
  const size_t bytes = rowb * nrows;
  // my lowest rows -> `down`'s ghost_hi (its rows Yloc ..) ; my highest rows -> `up`'s ghost_lo (its rows -nrows .. -1)
  char* recv_lo = (char*)op->ghost_lo + rowb * (op->ghost_depth - nrows);
  char* recv_hi = (char*)op->ghost_hi;
  // Order matters on 2 ranks, where `up` and `down` are the same peer and NCCL pairs the k-th send
  // with the peer's k-th receive: (hi -> up) must meet the peer's (ghost_lo <- down).
  GLB_NCCL(g_nccl.GroupStart());
  GLB_NCCL(g_nccl.Send(send_hi, bytes, NCCL_CHAR, up, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.Send(send_lo, bytes, NCCL_CHAR, down, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.Recv(recv_lo, bytes, NCCL_CHAR, down, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.Recv(recv_hi, bytes, NCCL_CHAR, up, ctx->comm->nccl, ctx->stream));
  GLB_NCCL(g_nccl.GroupEnd());
  return GLB_OK;
}

int halo_exchange(glb_operator* op, const void* in, int nrows) {
  if (op->ctx->nranks == 1) return GLB_OK;
  if (nrows > op->Yloc) return fail(GLB_ERR_ARG, "slab thinner than the halo");
  const size_t rowb = (size_t)op->X * op->nc * elem_bytes(op->dtype);
  const char* base = (const char*)in;
  return halo_exchange_ptrs(op, base, base + rowb * (op->Yloc - nrows), nrows);
}

int allreduce_sum(glb_context* ctx, double* vals, int n) {
  if (ctx->nranks == 1) return GLB_OK;
  if (!ctx->comm) return fail(GLB_ERR_STATE, "reduction before glb_comm_init");
  if (n > 64) return fail(GLB_ERR_ARG, "allreduce_sum: too many values");
  std::memcpy(ctx->comm->h_red, vals, sizeof(double) * n);
  GLB_CUDA(cudaMemcpyAsync(ctx->comm->d_red, ctx->comm->h_red, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  GLB_NCCL(g_nccl.AllReduce(ctx->comm->d_red, ctx->comm->d_red, n, NCCL_FLOAT64, NCCL_SUM, ctx->comm->nccl, ctx->stream));
  GLB_CUDA(cudaMemcpyAsync(ctx->comm->h_red, ctx->comm->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(vals, ctx->comm->h_red, sizeof(double) * n);
  return GLB_OK;
}

// in-stream sum over ranks of n doubles living in device memory (device-resident CG)
int allreduce_device(glb_context* ctx, double* d_vals, int n) {
  if (ctx->nranks == 1) return GLB_OK;
  if (!ctx->comm) return fail(GLB_ERR_STATE, "reduction before glb_comm_init");
  GLB_NCCL(g_nccl.AllReduce(d_vals, d_vals, n, NCCL_FLOAT64, NCCL_SUM, ctx->comm->nccl, ctx->stream));
  return GLB_OK;
}

}  // namespace glb

using namespace glb;

extern "C" {

int glb_comm_unique_id(char id[GLB_COMM_ID_BYTES]) {
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t u;
  GLB_NCCL(g_nccl.GetUniqueId(&u));
  std::memcpy(id, u.internal, GLB_COMM_ID_BYTES);
  return GLB_OK;
}

int glb_comm_init(glb_context* ctx, int rank, int nranks, const char id[GLB_COMM_ID_BYTES]) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GLB_ERR_ARG, "glb_comm_init: bad rank/nranks");
  if (ctx->comm) return fail(GLB_ERR_STATE, "communicator already initialised");
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return GLB_OK;
  int rc = load_nccl();
  if (rc) return rc;
  GLB_CUDA(cudaSetDevice(ctx->device));
  Comm* c = new Comm();
  ncclUniqueId_t u;
  std::memcpy(u.internal, id, GLB_COMM_ID_BYTES);
  GLB_NCCL(g_nccl.CommInitRank(&c->nccl, nranks, u, rank));
  GLB_CUDA(cudaMalloc(&c->d_red, sizeof(double) * 64));
  GLB_CUDA(cudaHostAlloc((void**)&c->h_red, sizeof(double) * 64, cudaHostAllocDefault));
  ctx->comm = c;
  return GLB_OK;
}

int glb_comm_rank(glb_context* ctx) { return ctx->rank; }
int glb_comm_size(glb_context* ctx) { return ctx->nranks; }

int glb_comm_barrier(glb_context* ctx) {
  double z = 0.0;
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  return allreduce_sum(ctx, &z, 1);
}

int glb_comm_export_mailbox(glb_context* ctx, char handle[GLB_IPC_HANDLE_BYTES]) {
  (void)ctx;
  (void)handle;
  return fail(GLB_ERR_STATE, "peer-memory mailboxes are not enabled in this build");
}
int glb_comm_attach_mailboxes(glb_context* ctx, const char* all_handles) {
  (void)ctx;
  (void)all_handles;
  return fail(GLB_ERR_STATE, "peer-memory mailboxes are not enabled in this build");
}

}  // extern "C"
