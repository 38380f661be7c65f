// mg.cu -- the grid-transfer operators of the adaptive multigrid preconditioner: prolong and restrict
// of multigrid/aa_mg/mg_complex.cpp:372-467 with device-resident vectors.
//
// A transfer couples a fine level (Xf x Yf sites, dof_f values per site) to the coarse level made of
// bx x by blocks with nvec values per coarse site (one per null vector).  The reference keeps the null
// vectors as nvec separate fine-sized arrays null_vectors[v][f] (mg_complex.h:164); here they are
// interleaved once at creation into null[f*nvec + v], so that the nvec numbers a fine dof needs are one
// contiguous run (128 bytes at nvec = 8) and the nvec threads of a coarse site read consecutive 16-byte
// elements in restrict.
//
//   prolong : fine[f]    = sum_v  null[v][f] * coarse[block(f)*nvec + v]          (v ascending)
//   restrict: coarse[i]  = sum_{f in block, y-major, x, dof}  conj(null[v][f]) * fine[f]
// Both run the reference's accumulation order inside one thread without FMA contraction, so the
// results are bit-identical to the CPU code (it zeroes the target and accumulates with +=).
// Block-local: on y-slabs whose height is a multiple of by no communication is needed; the handle is
// created with the LOCAL extent.
#include "runtime.hpp"

struct glb_mg_transfer {
  glb_context* ctx = nullptr;
  int Xf = 0, Yf = 0, dof_f = 1, bx = 1, by = 1, nvec = 1;
  int Xc = 0, Yc = 0;
  glb::cplx* null = nullptr;  // [fine dof][v]
};

namespace glb {

struct MgArgs {
  const cplx* null;
  int Xf, Yf, dof_f, bx, by, nvec, Xc, Yc;
};

template <int NV>
__global__ void __launch_bounds__(256) mg_prolong_kernel(const MgArgs a, cplx* __restrict__ fine,
                                                         const cplx* __restrict__ coarse) {
  const int nvec = (NV > 0) ? NV : a.nvec;
  const size_t nf = (size_t)a.Xf * a.Yf * a.dof_f;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (size_t)gridDim.x * blockDim.x) {
    const size_t site = f / a.dof_f;
    const int x = (int)(site % a.Xf), y = (int)(site / a.Xf);
    const size_t cs = (size_t)(y / a.by) * a.Xc + (x / a.bx);
    const cplx* nv = a.null + f * nvec;
    const cplx* cv = coarse + cs * nvec;
    cplx acc = mk(0.0, 0.0);
    if (NV > 0 && NV % 2 == 0) {
#pragma unroll
      for (int v = 0; v < NV; v += 2) {
        cplx n2[2], c2[2];
        ldv_nc<2>(nv + v, n2);
        ldv<2>(cv + v, c2);
        acc = fadd(acc, fmul(n2[0], c2[0]));
        acc = fadd(acc, fmul(n2[1], c2[1]));
      }
    } else {
      for (int v = 0; v < nvec; v++) acc = fadd(acc, fmul(__ldg(nv + v), cv[v]));
    }
    fine[f] = acc;
  }
}

__global__ void __launch_bounds__(256) mg_restrict_kernel(const MgArgs a, cplx* __restrict__ coarse,
                                                          const cplx* __restrict__ fine) {
  const size_t nc = (size_t)a.Xc * a.Yc * a.nvec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % a.nvec);
    const size_t cs = i / a.nvec;
    const int xc = (int)(cs % a.Xc), yc = (int)(cs / a.Xc);
    cplx acc = mk(0.0, 0.0);
    for (int y = yc * a.by; y < (yc + 1) * a.by; y++) {
      for (int x = xc * a.bx; x < (xc + 1) * a.bx; x++) {
        const size_t f0 = ((size_t)y * a.Xf + x) * a.dof_f;
        for (int d = 0; d < a.dof_f; d++)
          acc = fadd(acc, fcmul(__ldg(a.null + (f0 + d) * a.nvec + v), fine[f0 + d]));
      }
    }
    coarse[i] = acc;
  }
}

// host arrays null_vectors[v][f] -> device null[f*nvec + v]
__global__ void mg_interleave_kernel(cplx* __restrict__ dst, const cplx* __restrict__ src, size_t nf, int nvec, int v) {
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (size_t)gridDim.x * blockDim.x)
    dst[f * nvec + v] = src[f];
}

}  // namespace glb

using namespace glb;

extern "C" {

int glb_mg_transfer_create(glb_context* ctx, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                           const void* const* null_vectors, glb_mg_transfer** out) {
  if (!ctx || !null_vectors || !out) return fail(GLB_ERR_ARG, "glb_mg_transfer_create: null argument");
  if (Xf < 1 || Yf < 1 || dof_f < 1 || bx < 1 || by < 1 || nvec < 1)
    return fail(GLB_ERR_ARG, "glb_mg_transfer_create: extents must be positive");
  if (Xf % bx != 0 || Yf % by != 0)
    return fail(GLB_ERR_ARG, "glb_mg_transfer_create: the block size must divide the (local) fine lattice");
  GLB_CUDA(cudaSetDevice(ctx->device));
  glb_mg_transfer* t = new glb_mg_transfer();
  t->ctx = ctx;
  t->Xf = Xf;
  t->Yf = Yf;
  t->dof_f = dof_f;
  t->bx = bx;
  t->by = by;
  t->nvec = nvec;
  t->Xc = Xf / bx;
  t->Yc = Yf / by;
  const size_t nf = (size_t)Xf * Yf * dof_f;
  cplx* stage = nullptr;
  if (cudaMalloc(&t->null, nf * nvec * sizeof(cplx)) != cudaSuccess || cudaMalloc(&stage, nf * sizeof(cplx)) != cudaSuccess) {
    cudaFree(t->null);
    delete t;
    return fail(GLB_ERR_CUDA, "glb_mg_transfer_create: out of device memory");
  }
  const int grid = blas_grid(ctx, nf, 256, 1);
  for (int v = 0; v < nvec; v++) {
    if (!null_vectors[v]) {
      cudaFree(stage);
      cudaFree(t->null);
      delete t;
      return fail(GLB_ERR_ARG, "glb_mg_transfer_create: null vector pointer is null");
    }
    cudaMemcpyAsync(stage, null_vectors[v], nf * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream);
    mg_interleave_kernel<<<grid, 256, 0, ctx->stream>>>(t->null, stage, nf, nvec, v);
    cudaStreamSynchronize(ctx->stream);  // the host array may be pageable
  }
  cudaFree(stage);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(t->null);
    delete t;
    return fail(GLB_ERR_CUDA, std::string("glb_mg_transfer_create: ") + cudaGetErrorString(e));
  }
  *out = t;
  return GLB_OK;
}

int glb_mg_transfer_destroy(glb_mg_transfer* t) {
  if (!t) return GLB_OK;
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->null);
  delete t;
  return GLB_OK;
}

size_t glb_mg_fine_size(const glb_mg_transfer* t) { return t ? (size_t)t->Xf * t->Yf * t->dof_f : 0; }
size_t glb_mg_coarse_size(const glb_mg_transfer* t) { return t ? (size_t)t->Xc * t->Yc * t->nvec : 0; }

static MgArgs mg_args(const glb_mg_transfer* t) {
  MgArgs a;
  a.null = t->null;
  a.Xf = t->Xf;
  a.Yf = t->Yf;
  a.dof_f = t->dof_f;
  a.bx = t->bx;
  a.by = t->by;
  a.nvec = t->nvec;
  a.Xc = t->Xc;
  a.Yc = t->Yc;
  return a;
}

int glb_mg_prolong(glb_mg_transfer* t, void* d_fine, const void* d_coarse) {
  if (!t || !d_fine || !d_coarse) return fail(GLB_ERR_ARG, "glb_mg_prolong: null argument");
  glb_context* ctx = t->ctx;
  const MgArgs a = mg_args(t);
  const int grid = blas_grid(ctx, glb_mg_fine_size(t), 256, 1);
  if (t->nvec == 8)
    mg_prolong_kernel<8><<<grid, 256, 0, ctx->stream>>>(a, (cplx*)d_fine, (const cplx*)d_coarse);
  else if (t->nvec == 4)
    mg_prolong_kernel<4><<<grid, 256, 0, ctx->stream>>>(a, (cplx*)d_fine, (const cplx*)d_coarse);
  else
    mg_prolong_kernel<0><<<grid, 256, 0, ctx->stream>>>(a, (cplx*)d_fine, (const cplx*)d_coarse);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int glb_mg_restrict(glb_mg_transfer* t, void* d_coarse, const void* d_fine) {
  if (!t || !d_fine || !d_coarse) return fail(GLB_ERR_ARG, "glb_mg_restrict: null argument");
  glb_context* ctx = t->ctx;
  const MgArgs a = mg_args(t);
  const int grid = blas_grid(ctx, glb_mg_coarse_size(t), 256, 1);
  mg_restrict_kernel<<<grid, 256, 0, ctx->stream>>>(a, (cplx*)d_coarse, (const cplx*)d_fine);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

}  // extern "C"
