// runtime.cu -- context, stream, memory and error plumbing behind include/glb200.h.
#include "runtime.hpp"

#include <cstring>
#include <mutex>

namespace glb {

static thread_local std::string t_error;
std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }
int fail(int code, const std::string& msg) {
  t_error = msg;
  return code;
}

}  // namespace glb

using namespace glb;

extern "C" {

const char* glb_last_error(void) { return t_error.c_str(); }
unsigned long long glb_kernel_launches(void) { return g_launches.load(); }

int glb_create(int device, glb_context** out) {
  if (!out) return fail(GLB_ERR_ARG, "glb_create: null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GLB_ERR_CUDA, std::string("glb_create: no CUDA device (") + cudaGetErrorString(e) +
                                  "); this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(GLB_ERR_ARG, "glb_create: bad device index");
  GLB_CUDA(cudaSetDevice(device));
  glb_context* ctx = new glb_context();
  ctx->device = device;
  cudaDeviceProp prop;
  GLB_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  GLB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  GLB_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  GLB_CUDA(cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming));
  GLB_CUDA(cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming));
  // reduction workspace: partials, ticket, device + mapped-host result slots
  const int max_red = 2 * 16 + 8;  // multi_dot of 16 complex vectors is the widest reduction
  GLB_CUDA(cudaMalloc(&ctx->red.partials, sizeof(double) * max_red * MAX_PARTIAL_BLOCKS));
  GLB_CUDA(cudaMalloc(&ctx->red.ticket, sizeof(unsigned int)));
  GLB_CUDA(cudaMemset(ctx->red.ticket, 0, sizeof(unsigned int)));
  GLB_CUDA(cudaMalloc(&ctx->red.result_dev, sizeof(double) * max_red));
  GLB_CUDA(cudaHostAlloc((void**)&ctx->result_host_ptr, sizeof(double) * max_red, cudaHostAllocMapped));
  GLB_CUDA(cudaHostGetDevicePointer((void**)&ctx->red.result_host, ctx->result_host_ptr, 0));
  ctx->table_bytes = 4096;
  GLB_CUDA(cudaHostAlloc(&ctx->h_table, ctx->table_bytes, cudaHostAllocDefault));  // pinned scratch (CG state polls)
  // keep freed vector memory in the stream-ordered pool: solvers allocate work vectors per call
  // (generic_cg.cpp:292-294) and GCR even per iteration (generic_gcr.cpp:282-283)
  cudaMemPool_t pool;
  GLB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ull;
  GLB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  *out = ctx;
  return GLB_OK;
}

int glb_destroy(glb_context* ctx) {
  if (!ctx) return GLB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  glb_prof_enable(ctx, 1);  // drops recorded events
  comm_destroy(ctx);
  cudaFree(ctx->red.partials);
  cudaFree(ctx->red.ticket);
  cudaFree(ctx->red.result_dev);
  cudaFreeHost(ctx->result_host_ptr);
  cudaFreeHost(ctx->h_table);
  cudaEventDestroy(ctx->ev_a);
  cudaEventDestroy(ctx->ev_b);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
  return GLB_OK;
}

int glb_synchronize(glb_context* ctx) {
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  return GLB_OK;
}
void* glb_stream(glb_context* ctx) { return (void*)ctx->stream; }
int glb_device(glb_context* ctx) { return ctx->device; }
int glb_sm_count(glb_context* ctx) { return ctx->sm_count; }

int glb_prof_enable(glb_context* ctx, int on) {
  if (!ctx) return fail(GLB_ERR_ARG, "glb_prof_enable: null context");
  if (on) {
    for (auto& r : ctx->prof) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
  }
  ctx->prof_on = (on != 0);
  return GLB_OK;
}
int glb_prof_read(glb_context* ctx, int cls, int cap, float* ms, int* n) {
  if (!ctx || !n) return fail(GLB_ERR_ARG, "glb_prof_read: null argument");
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  int k = 0;
  for (auto& r : ctx->prof) {
    if (r.cls != cls) continue;
    if (ms && k < cap) GLB_CUDA(cudaEventElapsedTime(&ms[k], r.a, r.b));
    k++;
  }
  *n = k;
  return GLB_OK;
}

int glb_prof_summary(glb_context* ctx, int cls, int* launches, double* ms_total, double* bytes_total) {
  if (!ctx || !launches || !ms_total || !bytes_total) return fail(GLB_ERR_ARG, "glb_prof_summary: null argument");
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  int k = 0;
  double ms = 0.0, by = 0.0;
  for (auto& r : ctx->prof) {
    if (r.cls != cls) continue;
    float t = 0.f;
    GLB_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    by += r.bytes;
    k++;
  }
  *launches = k;
  *ms_total = ms;
  *bytes_total = by;
  return GLB_OK;
}

int glb_vec_alloc(glb_context* ctx, int dtype, size_t n, void** dptr) {
  if (!dptr) return fail(GLB_ERR_ARG, "glb_vec_alloc: null output");
  size_t bytes = n * elem_bytes(dtype);
  if (bytes == 0) bytes = 32;
  GLB_CUDA(cudaMallocAsync(dptr, bytes, ctx->stream));
  return GLB_OK;
}
int glb_vec_free(glb_context* ctx, void* dptr) {
  if (!dptr) return GLB_OK;
  GLB_CUDA(cudaFreeAsync(dptr, ctx->stream));
  return GLB_OK;
}
int glb_vec_upload(glb_context* ctx, int dtype, size_t n, void* dst, const void* src) {
  GLB_CUDA(cudaMemcpyAsync(dst, src, n * elem_bytes(dtype), cudaMemcpyHostToDevice, ctx->stream));
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));  // src may be pageable / reused by the caller
  return GLB_OK;
}
int glb_vec_download(glb_context* ctx, int dtype, size_t n, void* dst, const void* src) {
  GLB_CUDA(cudaMemcpyAsync(dst, src, n * elem_bytes(dtype), cudaMemcpyDeviceToHost, ctx->stream));
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  return GLB_OK;
}
int glb_vec_zero(glb_context* ctx, int dtype, size_t n, void* d) {
  GLB_CUDA(cudaMemsetAsync(d, 0, n * elem_bytes(dtype), ctx->stream));
  return GLB_OK;
}
int glb_vec_copy(glb_context* ctx, int dtype, size_t n, void* dst, const void* src) {
  GLB_CUDA(cudaMemcpyAsync(dst, src, n * elem_bytes(dtype), cudaMemcpyDeviceToDevice, ctx->stream));
  return GLB_OK;
}
int glb_host_alloc(glb_context* ctx, size_t bytes, void** hptr) {
  (void)ctx;
  GLB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 32, cudaHostAllocDefault));
  return GLB_OK;
}
int glb_host_free(glb_context* ctx, void* hptr) {
  (void)ctx;
  if (hptr) GLB_CUDA(cudaFreeHost(hptr));
  return GLB_OK;
}

}  // extern "C"
