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
  // y-slabs: the neighbours' boundary rows of the null vectors in the same [dof][v] layout (one row each), filled by
  // glb_mg_galerkin -- the only consumer: hops across the slab edge
  glb::cplx* null_lo = nullptr;
  glb::cplx* null_hi = nullptr;
};

namespace glb {

struct MgArgs {
  const cplx* null;
  const cplx* null_lo;  // row -1 / row Yf of the slab (nullptr on a single rank: periodic wrap inside the array)
  const cplx* null_hi;
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

// ------------------------------------------------------------------------------------------------
// Set-up on the device.
//
// block_orthonormalize + block_normalize (mg_complex.cpp:191-370): one thread owns one block and runs the
// reference's loops in its order (normalise vector v-1, project v against 0..v-1, ..., final normalisation), so the
// result is bit-identical to the CPU code.  The null vectors are nvec separate device arrays (the reference layout).
struct MgNullPtrs {
  cplx* v[32];
};
__global__ void __launch_bounds__(128) mg_block_orthonormalize_kernel(const MgNullPtrs n, int Xf, int dof_f, int bx, int by,
                                                                      int nvec, int Xc, int Yc) {
  const size_t nb = (size_t)Xc * Yc;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const int x_lo = (int)(b % Xc) * bx, y_lo = (int)(b / Xc) * by;
    auto for_sites = [&](auto fn) {
      for (int y = y_lo; y < y_lo + by; y++)
        for (int x = x_lo; x < x_lo + bx; x++)
          for (int d = 0; d < dof_f; d++) fn(((size_t)y * Xf + x) * dof_f + d);
    };
    auto normalise = [&](cplx* v) {
      double nrm = 0.0;
      for_sites([&](size_t f) { nrm = xadd(nrm, fnorm(v[f])); });
      nrm = sqrt(nrm);
      for_sites([&](size_t f) { v[f] = frdiv(v[f], nrm); });
    };
    for (int c = 1; c < nvec; c++) {
      normalise(n.v[c - 1]);
      for (int m = 0; m < c; m++) {
        cplx dot = mk(0.0, 0.0);
        for_sites([&](size_t f) { dot = fadd(dot, fcmul(n.v[m][f], n.v[c][f])); });
        for_sites([&](size_t f) { n.v[c][f] = fsub(n.v[c][f], fmul(dot, n.v[m][f])); });
      }
    }
    for (int c = 0; c < nvec; c++) normalise(n.v[c]);  // block_normalize (mg_complex.cpp:191-256)
  }
}

// null_partition_staggered / null_partition_coarse (null_gen.cpp:13-160): the elements of class `which` move from
// `src_io` to `dst_out` and are zeroed in place.  Classes: BLOCK_EO (nparts = 2) -- 1 = odd sites on the top level, below
// it the elements whose index modulo colour_period lies in the upper half of the period; BLOCK_CORNER (nparts = 4) --
// top level: 1 = (x odd, y odd), 2 = (x odd, y even), 3 = (x even, y odd) (null_gen.cpp:74-88); below it the quarters
// of the period with the reference's integer bounds p/4, 2p/4, 3p/4 (:132-152).
__global__ void mg_partition_kernel(cplx* __restrict__ src_io, cplx* __restrict__ dst_out, size_t n, int X, int dof, int y0,
                                    int colour_period, int nparts, int which) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int cls;
    if (colour_period > 0) {
      // the reference takes the GLOBAL index modulo the period (null_gen.cpp:114); y0 rows precede this slab
      const int c = (int)((i + (size_t)y0 * X * dof) % colour_period), p = colour_period;
      if (nparts == 2)
        cls = (c >= p / 2) ? 1 : 0;
      else
        cls = (c >= p / 4 && c < 2 * p / 4) ? 1 : (c >= 2 * p / 4 && c < 3 * p / 4) ? 2 : (c >= 3 * p / 4) ? 3 : 0;
    } else {
      const size_t site = i / dof;
      const int xo = (int)(site % X) & 1, yo = ((int)(site / X) + y0) & 1;
      if (nparts == 2)
        cls = xo ^ yo;
      else
        cls = (xo && yo) ? 1 : (xo && !yo) ? 2 : (!xo && yo) ? 3 : 0;
    }
    if (cls == which) {
      dst_out[i] = src_io[i];
      src_io[i] = mk(0.0, 0.0);
    }
  }
}

// Galerkin coarse operator P^dag A P of a five-point fine stencil (what generate_coarse_from_fine_stencil,
// mg_complex.cpp:827-1026, assembles by probing; ignore_shifts = false: the fine shifts end up in the coarse clover).
// One thread per (coarse site, i, j): sums conj(n_i[f,a]) M[f][a,b] n_j[g,b] over the fine dofs of the block; a hop
// that stays inside the block feeds the coarse clover, one that leaves it the coarse hopping term of that direction.
struct MgFine {
  const cplx* clover;
  const cplx* hopping;
  cplx shift, eo_shift, dof_shift;
  int use_shift, use_eo, use_dof;
};
__global__ void __launch_bounds__(128) mg_galerkin_kernel(const MgArgs a, const MgFine fs, cplx* __restrict__ cl_c,
                                                          cplx* __restrict__ hp_c) {
  const int nv = a.nvec, df = a.dof_f;
  const size_t nc_sites = (size_t)a.Xc * a.Yc;
  const size_t total = nc_sites * nv * nv;
  const size_t Lf = (size_t)a.Xf * a.Yf * df;     // fine dofs
  const size_t Lc = nc_sites * nv;                // coarse dofs
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(id % nv), i = (int)((id / nv) % nv);
    const size_t cs = id / ((size_t)nv * nv);
    const int xc = (int)(cs % a.Xc), yc = (int)(cs / a.Xc);
    cplx acc_c = mk(0.0, 0.0), acc_h[4] = {mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0)};
    for (int y = yc * a.by; y < (yc + 1) * a.by; y++) {
      for (int x = xc * a.bx; x < (xc + 1) * a.bx; x++) {
        const size_t site = (size_t)y * a.Xf + x;
        const int xn[4] = {(x + 1 == a.Xf) ? 0 : x + 1, x, (x == 0) ? a.Xf - 1 : x - 1, x};
        const int yn[4] = {y, (y + 1 == a.Yf) ? 0 : y + 1, y, (y == 0) ? a.Yf - 1 : y - 1};
        const bool inside[4] = {(x + 1) % a.bx != 0, (y + 1) % a.by != 0, x % a.bx != 0, y % a.by != 0};
        for (int r = 0; r < df; r++) {
          const size_t f = site * df + r;
          const cplx ci = a.null[f * nv + i];
          // diagonal part: clover row + the shifts (coarse_stencil.cpp:153-169)
          cplx row = mk(0.0, 0.0);
          for (int c = 0; c < df; c++) row = fadd(row, fmul(fs.clover[c + df * f], a.null[(site * df + c) * nv + j]));
          const cplx self = a.null[f * nv + j];
          if (fs.use_shift) row = fadd(row, fmul(fs.shift, self));
          if (fs.use_eo) row = fadd(row, fmul(((x + y) & 1) ? fneg(fs.eo_shift) : fs.eo_shift, self));
          if (fs.use_dof) row = fadd(row, fmul(r < df / 2 ? fs.dof_shift : fneg(fs.dof_shift), self));
          acc_c = fadd(acc_c, fcmul(ci, row));
          for (int d = 0; d < 4; d++) {
            // the neighbour's null-vector entries: inside the slab, or (y-slabs) in the neighbour rank's boundary row
            const cplx* nb;
            if (a.null_lo && d == 1 && y + 1 == a.Yf)
              nb = a.null_hi + (size_t)xn[d] * df * nv;
            else if (a.null_lo && d == 3 && y == 0)
              nb = a.null_lo + (size_t)xn[d] * df * nv;
            else
              nb = a.null + ((size_t)yn[d] * a.Xf + xn[d]) * df * nv;
            cplx h = mk(0.0, 0.0);
            for (int c = 0; c < df; c++) h = fadd(h, fmul(fs.hopping[c + df * f + d * df * Lf], nb[(size_t)c * nv + j]));
            if (inside[d])
              acc_c = fadd(acc_c, fcmul(ci, h));
            else
              acc_h[d] = fadd(acc_h[d], fcmul(ci, h));
          }
        }
      }
    }
    const size_t o = (cs * nv + i) * nv + j;  // clover[c + nc*i'] with i' = cs*nc + i (coarse_stencil.h:41)
    cl_c[o] = acc_c;
#pragma unroll
    for (int d = 0; d < 4; d++) hp_c[o + (size_t)d * nv * Lc] = acc_h[d];
  }
}

// y-slabs: boundary rows of null vector v out of / into the [dof][v] layout (row = X*dof_f consecutive fine dofs)
__global__ void mg_row_gather_kernel(cplx* __restrict__ row_out, const cplx* __restrict__ null_row, size_t rowlen, int nvec, int v) {
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < rowlen; f += (size_t)gridDim.x * blockDim.x)
    row_out[f] = null_row[f * nvec + v];
}
__global__ void mg_row_scatter_kernel(cplx* __restrict__ null_row, const cplx* __restrict__ row_in, size_t rowlen, int nvec, int v) {
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < rowlen; f += (size_t)gridDim.x * blockDim.x)
    null_row[f * nvec + v] = row_in[f];
}

// host arrays null_vectors[v][f] -> device null[f*nvec + v]
__global__ void mg_interleave_kernel(cplx* __restrict__ dst, const cplx* __restrict__ src, size_t nf, int nvec, int v) {
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (size_t)gridDim.x * blockDim.x)
    dst[f * nvec + v] = src[f];
}

}  // namespace glb

using namespace glb;

static MgArgs mg_args(const glb_mg_transfer* t) {
  MgArgs a;
  a.null = t->null;
  a.null_lo = t->null_lo;
  a.null_hi = t->null_hi;
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

// the same from DEVICE-resident null vectors (after glb_mg_block_orthonormalize)
int glb_mg_transfer_create_dev(glb_context* ctx, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                               const void* const* d_null_vectors, glb_mg_transfer** out) {
  if (!ctx || !d_null_vectors || !out) return fail(GLB_ERR_ARG, "glb_mg_transfer_create_dev: null argument");
  if (Xf < 1 || Yf < 1 || dof_f < 1 || bx < 1 || by < 1 || nvec < 1 || Xf % bx != 0 || Yf % by != 0)
    return fail(GLB_ERR_ARG, "glb_mg_transfer_create_dev: bad extents");
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
  if (cudaMalloc(&t->null, nf * nvec * sizeof(cplx)) != cudaSuccess) {
    delete t;
    return fail(GLB_ERR_CUDA, "glb_mg_transfer_create_dev: out of device memory");
  }
  const int grid = blas_grid(ctx, nf, 256, 1);
  for (int v = 0; v < nvec; v++) {
    if (!d_null_vectors[v]) {
      cudaFree(t->null);
      delete t;
      return fail(GLB_ERR_ARG, "glb_mg_transfer_create_dev: null vector pointer is null");
    }
    mg_interleave_kernel<<<grid, 256, 0, ctx->stream>>>(t->null, (const cplx*)d_null_vectors[v], nf, nvec, v);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(t->null);
    delete t;
    return fail(GLB_ERR_CUDA, std::string("glb_mg_transfer_create_dev: ") + cudaGetErrorString(e));
  }
  *out = t;
  return GLB_OK;
}

int glb_mg_block_orthonormalize(glb_context* ctx, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                                void* const* d_null_vectors) {
  if (!ctx || !d_null_vectors) return fail(GLB_ERR_ARG, "glb_mg_block_orthonormalize: null argument");
  if (nvec < 1 || nvec > 32) return fail(GLB_ERR_ARG, "glb_mg_block_orthonormalize: 1..32 null vectors");
  if (Xf % bx != 0 || Yf % by != 0) return fail(GLB_ERR_ARG, "glb_mg_block_orthonormalize: the block size must divide the lattice");
  MgNullPtrs n{};
  for (int v = 0; v < nvec; v++) n.v[v] = (cplx*)d_null_vectors[v];
  const size_t nb = (size_t)(Xf / bx) * (Yf / by);
  const int grid = blas_grid(ctx, nb, 128, 1);
  mg_block_orthonormalize_kernel<<<grid, 128, 0, ctx->stream>>>(n, Xf, dof_f, bx, by, nvec, Xf / bx, Yf / by);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

static int partition_launch(glb_context* ctx, int X, int Y, int dof, int colour_period, int nparts, int which, void* d_src_io,
                            void* d_dst_out, const char* who) {
  if (!ctx || !d_src_io || !d_dst_out) return fail(GLB_ERR_ARG, std::string(who) + ": null argument");
  if (X < 1 || Y < 1 || dof < 1 || colour_period < 0 || which < 1 || which >= nparts)
    return fail(GLB_ERR_ARG, std::string(who) + ": bad extents or class");
  if (d_src_io == d_dst_out) return fail(GLB_ERR_ARG, std::string(who) + ": source and target must differ");
  int y0 = 0, Yloc = Y;  // Y is the GLOBAL extent; on y-slabs the vectors hold this rank's rows
  if (glb_slab_bounds(ctx, Y, &y0, &Yloc) != GLB_OK) return GLB_ERR_ARG;
  const size_t n = (size_t)X * Yloc * dof;
  const int grid = blas_grid(ctx, n, 256, 1);
  mg_partition_kernel<<<grid, 256, 0, ctx->stream>>>((cplx*)d_src_io, (cplx*)d_dst_out, n, X, dof, y0, colour_period, nparts, which);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int glb_mg_partition(glb_context* ctx, int X, int Y, int dof, int colour_period, void* d_even_io, void* d_odd_out) {
  return partition_launch(ctx, X, Y, dof, colour_period, 2, 1, d_even_io, d_odd_out, "glb_mg_partition");
}

int glb_mg_partition_corner(glb_context* ctx, int X, int Y, int dof, int colour_period, int which, void* d_src_io,
                            void* d_dst_out) {
  return partition_launch(ctx, X, Y, dof, colour_period, 4, which, d_src_io, d_dst_out, "glb_mg_partition_corner");
}

int glb_mg_galerkin(glb_mg_transfer* t, glb_operator* fine, int ignore_shifts, glb_operator** coarse) {
  if (!t || !fine || !coarse) return fail(GLB_ERR_ARG, "glb_mg_galerkin: null argument");
  glb_context* ctx = t->ctx;
  if (fine->kind != OPK_STENCIL || fine->has_two)
    return fail(GLB_ERR_ARG, "glb_mg_galerkin: the fine operator must be a five-point stencil2d operator");
  if (fine->X != t->Xf || fine->Yloc != t->Yf || fine->nc != t->dof_f)
    return fail(GLB_ERR_ARG, "glb_mg_galerkin: transfer and fine operator disagree (on y-slabs the transfer holds the local rows)");
  if (fine->Y % t->by != 0 || fine->y0 % t->by != 0)
    return fail(GLB_ERR_ARG, "glb_mg_galerkin: slab boundaries must coincide with block boundaries");
  const int Yc_global = fine->Y / t->by;
  if (t->Xc < 2 || Yc_global < 2 || (t->Xc & 1) || (Yc_global & 1))
    return fail(GLB_ERR_ARG, "glb_mg_galerkin: the coarse lattice needs an even number (>= 2) of sites per direction");
  if (ctx->nranks > 1) {
    // the neighbours' boundary rows of every null vector, through the fine operator's halo path (one row per side)
    const size_t rowlen = (size_t)t->Xf * t->dof_f;
    if (!t->null_lo) {
      if (cudaMalloc(&t->null_lo, rowlen * t->nvec * sizeof(cplx)) != cudaSuccess ||
          cudaMalloc(&t->null_hi, rowlen * t->nvec * sizeof(cplx)) != cudaSuccess)
        return fail(GLB_ERR_CUDA, "glb_mg_galerkin: out of device memory");
    }
    if (!fine->send_lo || !fine->send_hi) return fail(GLB_ERR_STATE, "glb_mg_galerkin: the fine operator has no halo staging");
    const int g1 = blas_grid(ctx, rowlen, 256, 1);
    const cplx* first = t->null;
    const cplx* last = t->null + (size_t)(t->Yf - 1) * rowlen * t->nvec;
    for (int v = 0; v < t->nvec; v++) {
      mg_row_gather_kernel<<<g1, 256, 0, ctx->stream>>>((cplx*)fine->send_lo, first, rowlen, t->nvec, v);
      mg_row_gather_kernel<<<g1, 256, 0, ctx->stream>>>((cplx*)fine->send_hi, last, rowlen, t->nvec, v);
      GLB_LAUNCH_CHECK();
      int rc = halo_exchange_ptrs(fine, fine->send_lo, fine->send_hi, 1);
      if (rc) return rc;
      // ghost_lo holds ghost_depth rows, the nearest one (row -1) last; ghost_hi starts with row Yloc
      const cplx* glo = (const cplx*)fine->ghost_lo + (size_t)(fine->ghost_depth - 1) * rowlen;
      mg_row_scatter_kernel<<<g1, 256, 0, ctx->stream>>>(t->null_lo, glo, rowlen, t->nvec, v);
      mg_row_scatter_kernel<<<g1, 256, 0, ctx->stream>>>(t->null_hi, (const cplx*)fine->ghost_hi, rowlen, t->nvec, v);
      GLB_LAUNCH_CHECK();
    }
  }
  const int nv = t->nvec;
  const size_t per = (size_t)t->Xc * t->Yc * nv * nv;
  cplx *cl = nullptr, *hp = nullptr;
  if (cudaMalloc(&cl, per * sizeof(cplx)) != cudaSuccess || cudaMalloc(&hp, 4 * per * sizeof(cplx)) != cudaSuccess) {
    cudaFree(cl);
    return fail(GLB_ERR_CUDA, "glb_mg_galerkin: out of device memory");
  }
  MgFine fs;
  fs.clover = fine->clover;
  fs.hopping = fine->hopping;
  fs.shift = make_double2(fine->shift[0], fine->shift[1]);
  fs.eo_shift = make_double2(fine->eo_shift[0], fine->eo_shift[1]);
  fs.dof_shift = make_double2(fine->dof_shift[0], fine->dof_shift[1]);
  fs.use_shift = !ignore_shifts && (fine->shift[0] != 0.0 || fine->shift[1] != 0.0);
  fs.use_eo = !ignore_shifts && (fine->eo_shift[0] != 0.0 || fine->eo_shift[1] != 0.0);
  fs.use_dof = !ignore_shifts && (fine->dof_shift[0] != 0.0 || fine->dof_shift[1] != 0.0);
  const int grid = blas_grid(ctx, per, 128, 1);
  mg_galerkin_kernel<<<grid, 128, 0, ctx->stream>>>(mg_args(t), fs, cl, hp);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(cl);
    cudaFree(hp);
    return fail(GLB_ERR_CUDA, std::string("glb_mg_galerkin: ") + cudaGetErrorString(e));
  }
  int rc = op_adopt_stencil2d(ctx, t->Xc, Yc_global, t->Yc, nv, cl, hp, coarse);
  if (rc) {
    cudaFree(cl);
    cudaFree(hp);
  }
  return rc;
}

int glb_mg_transfer_destroy(glb_mg_transfer* t) {
  if (!t) return GLB_OK;
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->null);
  cudaFree(t->null_lo);
  cudaFree(t->null_hi);
  delete t;
  return GLB_OK;
}

size_t glb_mg_fine_size(const glb_mg_transfer* t) { return t ? (size_t)t->Xf * t->Yf * t->dof_f : 0; }
size_t glb_mg_coarse_size(const glb_mg_transfer* t) { return t ? (size_t)t->Xc * t->Yc * t->nvec : 0; }


int glb_mg_prolong(glb_mg_transfer* t, void* d_fine, const void* d_coarse) {
  if (!t || !d_fine || !d_coarse) return fail(GLB_ERR_ARG, "glb_mg_prolong: null argument");
  glb_context* ctx = t->ctx;
  const MgArgs a = mg_args(t);
  const int grid = blas_grid(ctx, glb_mg_fine_size(t), 256, 1);
  // null vectors (nvec per fine dof) + the fine vector + the coarse vector
  ProfScope prof(ctx, PROF_MG_TRANSFER, 16.0 * ((double)glb_mg_fine_size(t) * (t->nvec + 1) + (double)glb_mg_coarse_size(t)));
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
  ProfScope prof(ctx, PROF_MG_TRANSFER, 16.0 * ((double)glb_mg_fine_size(t) * (t->nvec + 1) + (double)glb_mg_coarse_size(t)));
  mg_restrict_kernel<<<grid, 256, 0, ctx->stream>>>(a, (cplx*)d_coarse, (const cplx*)d_fine);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

}  // extern "C"
