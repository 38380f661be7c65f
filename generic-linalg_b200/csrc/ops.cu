// ops.cu -- operator objects of include/glb200.h: creation (upload + re-layout of links /
// stencil matrices, slab extraction), apply, apply with fused inner products.
#include <vector>

#include "runtime.hpp"

namespace glb {

// AoS links lattice[y*X*2 + x*2 + mu] (operators.cpp:215-224) -> two site-major planes
__global__ void split_links_kernel(const cplx* aos, cplx* Ux, cplx* Uy, size_t nsites) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsites; i += (size_t)gridDim.x * blockDim.x) {
    Ux[i] = aos[2 * i];
    Uy[i] = aos[2 * i + 1];
  }
}

static void slab_of(const glb_context* ctx, int Y, int* y0, int* Yloc) {
  const long long g = ctx->rank, G = ctx->nranks;
  const int a = (int)((long long)Y * g / G), b = (int)((long long)Y * (g + 1) / G);
  *y0 = a;
  *Yloc = b - a;
}

static int new_op(glb_context* ctx, int kind, int dtype, int X, int Y, int nc, glb_operator** out) {
  if (!ctx || !out) return fail(GLB_ERR_ARG, "null context / output");
  if (X < 1 || Y < 1 || nc < 1) return fail(GLB_ERR_ARG, "bad lattice dimensions");
  if (Y < ctx->nranks) return fail(GLB_ERR_ARG, "fewer lattice rows than ranks");
  glb_operator* op = new glb_operator();
  op->ctx = ctx;
  op->kind = kind;
  op->dtype = dtype;
  op->X = X;
  op->Y = Y;
  op->nc = nc;
  slab_of(ctx, Y, &op->y0, &op->Yloc);
  *out = op;
  return GLB_OK;
}

static int alloc_ghosts(glb_operator* op, int depth) {
  op->ghost_depth = depth;
  if (op->ctx->nranks == 1) return GLB_OK;
  if (op->Yloc < depth) return fail(GLB_ERR_ARG, "slab thinner than the stencil reach: use fewer ranks");
  const size_t bytes = (size_t)depth * op->X * op->nc * elem_bytes(op->dtype);
  GLB_CUDA(cudaMalloc(&op->send_lo, bytes));
  GLB_CUDA(cudaMalloc(&op->send_hi, bytes));
  // peer-memory path: [parity 0: lo|hi][parity 1: lo|hi][flag_lo, flag_hi] at the same offset on every rank
  size_t off = 0;
  unsigned long long seq0 = 0;
  char* area = (char*)comm_arena_alloc(op->ctx, 4 * bytes + 256, &off, &seq0);
  if (area) {
    op->ghost_p2p = true;
    op->ghost_off = off;
    op->ghost_arena_bytes = 4 * bytes + 256;
    op->halo_seq = seq0;  // a reused region: carry on counting where its previous owner stopped
    op->ghost_lo = area;
    op->ghost_hi = area + bytes;
    return GLB_OK;
  }
  GLB_CUDA(cudaMalloc(&op->ghost_lo, bytes));
  GLB_CUDA(cudaMalloc(&op->ghost_hi, bytes));
  return GLB_OK;
}

// Link planes are stored with LINK_GHOST periodic ghost rows below and above the slab, so the
// kernels address U(x, y) for y in [-LINK_GHOST, Yloc+LINK_GHOST) by plain pointer arithmetic
// (single rank: wrapped copies of the slab's own rows; slabs: the neighbours' rows, static).
// h_links is either the GLOBAL reference array lattice[y*X*2 + x*2 + mu], or (local = true) only
// this rank's rows with the same ghost rows: rows y0-2 .. y0+Yloc+1, (Yloc+4)*X*2 complex.
static int upload_links(glb_operator* op, const void* h_links, bool local = false) {
  glb_context* ctx = op->ctx;
  const int G = LINK_GHOST;
  const size_t X = op->X, rows = (size_t)op->Yloc + 2 * G, n = X * rows;
  const cplx* h = (const cplx*)h_links;
  cplx* aos = nullptr;
  GLB_CUDA(cudaMalloc(&aos, sizeof(cplx) * 2 * n));
  GLB_CUDA(cudaMalloc(&op->Ux_store, sizeof(cplx) * n));
  GLB_CUDA(cudaMalloc(&op->Uy_store, sizeof(cplx) * n));
  op->Ux = op->Ux_store + (size_t)G * X;
  op->Uy = op->Uy_store + (size_t)G * X;
  op->Uy_lo = op->Uy - X;
  if (local) {
    GLB_CUDA(cudaMemcpyAsync(aos, h, sizeof(cplx) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    // slab rows in one copy, ghost rows one by one (periodic in the global lattice)
    GLB_CUDA(cudaMemcpyAsync(aos + 2 * (size_t)G * X, h + 2 * (size_t)op->y0 * X, sizeof(cplx) * 2 * X * op->Yloc,
                             cudaMemcpyHostToDevice, ctx->stream));
    for (int g = 0; g < G; g++) {
      const int ylo = ((op->y0 - G + g) % op->Y + op->Y) % op->Y;
      const int yhi = (op->y0 + op->Yloc + g) % op->Y;
      GLB_CUDA(cudaMemcpyAsync(aos + 2 * (size_t)g * X, h + 2 * (size_t)ylo * X, sizeof(cplx) * 2 * X,
                               cudaMemcpyHostToDevice, ctx->stream));
      GLB_CUDA(cudaMemcpyAsync(aos + 2 * ((size_t)G + op->Yloc + g) * X, h + 2 * (size_t)yhi * X, sizeof(cplx) * 2 * X,
                               cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  const int grid = blas_grid(ctx, n, 256, 4);
  split_links_kernel<<<grid, 256, 0, ctx->stream>>>(aos, op->Ux_store, op->Uy_store, n);
  GLB_LAUNCH_CHECK();
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  GLB_CUDA(cudaFree(aos));
  op->has_links = true;
  return GLB_OK;
}

}  // namespace glb

using namespace glb;

// The constructors proper.  A failure part-way (upload, ghost rows, temporaries) leaves a half-built operator in
// *out; the extern "C" wrappers below destroy it and hand back a null handle.
namespace {

int create_laplace_impl(glb_context* ctx, int dtype, int X, int Y, int Nc, double diag_re, double diag_im,
                          glb_operator** out) {
  if (dtype != GLB_REAL && dtype != GLB_COMPLEX) return fail(GLB_ERR_ARG, "bad dtype");
  if (dtype == GLB_REAL && diag_im != 0.0) return fail(GLB_ERR_ARG, "real Laplacian with complex diagonal");
  int rc = new_op(ctx, OPK_LAPLACE, dtype, X, Y, Nc, out);
  if (rc) return rc;
  (*out)->diag_re = diag_re;
  (*out)->diag_im = diag_im;
  return alloc_ghosts(*out, 1);
}

// real free staggered operator of tests/multishift/multishift.cpp:677 (served by the simple
// nearest-neighbour kernel; flag 0x100 selects the staggered signs, diag carries the mass)
int create_staggered_free_real_impl(glb_context* ctx, int X, int Y, double mass, glb_operator** out) {
  int rc = new_op(ctx, OPK_LAPLACE, GLB_REAL, X, Y, 1, out);
  if (rc) return rc;
  (*out)->diag_re = mass;
  (*out)->diag_im = 0.0;
  (*out)->mass = mass;
  (*out)->flags = 0x100u;
  return alloc_ghosts(*out, 1);
}

int create_laplace_u1_impl(glb_context* ctx, const void* h_links, int X, int Y, double mass, glb_operator** out) {
  if (!h_links) return fail(GLB_ERR_ARG, "gauged Laplacian needs links");
  int rc = new_op(ctx, OPK_LAPLACE_U1, GLB_COMPLEX, X, Y, 1, out);
  if (rc) return rc;
  (*out)->mass = mass;
  rc = upload_links(*out, h_links);
  if (rc) return rc;
  return alloc_ghosts(*out, 1);
}

int create_staggered_impl(glb_context* ctx, const void* h_links, int X, int Y, double mass, unsigned flags,
                            glb_operator** out) {
  if ((flags & GLB_STAG_NORMAL) && (flags & (GLB_STAG_DAGGER | GLB_STAG_GAMMA5)))
    return fail(GLB_ERR_ARG, "NORMAL cannot be combined with DAGGER/GAMMA5");
  if ((flags & GLB_STAG_DAGGER) && (flags & GLB_STAG_GAMMA5)) return fail(GLB_ERR_ARG, "DAGGER+GAMMA5 unsupported");
  const unsigned eo = flags & (GLB_STAG_DEO | GLB_STAG_DOE | GLB_STAG_M2MDEODOE);
  if (eo && ((eo & (eo - 1)) || (flags & ~eo) || !h_links))
    return fail(GLB_ERR_ARG, "DEO / DOE / M2MDEODOE are exclusive, gauged-only flags");
  int rc = new_op(ctx, OPK_STAGGERED, GLB_COMPLEX, X, Y, 1, out);
  if (rc) return rc;
  glb_operator* op = *out;
  op->mass = mass;
  op->flags = flags;
  if (h_links) {
    rc = upload_links(op, h_links);
    if (rc) return rc;
  }
  if (flags & (GLB_STAG_NORMAL | GLB_STAG_M2MDEODOE)) GLB_CUDA(cudaMalloc(&op->tmp, sizeof(cplx) * (size_t)X * op->Yloc));
  return alloc_ghosts(op, 2);
}

int create_staggered_local_impl(glb_context* ctx, const void* h_links_local, int X, int Y, double mass,
                                    unsigned flags, glb_operator** out) {
  if (!h_links_local) return fail(GLB_ERR_ARG, "glb_op_create_staggered_local needs links");
  if ((flags & GLB_STAG_NORMAL) && (flags & (GLB_STAG_DAGGER | GLB_STAG_GAMMA5)))
    return fail(GLB_ERR_ARG, "NORMAL cannot be combined with DAGGER/GAMMA5");
  const unsigned eo = flags & (GLB_STAG_DEO | GLB_STAG_DOE | GLB_STAG_M2MDEODOE);
  if (eo && ((eo & (eo - 1)) || (flags & ~eo))) return fail(GLB_ERR_ARG, "DEO / DOE / M2MDEODOE are exclusive flags");
  int rc = new_op(ctx, OPK_STAGGERED, GLB_COMPLEX, X, Y, 1, out);
  if (rc) return rc;
  glb_operator* op = *out;
  op->mass = mass;
  op->flags = flags;
  rc = upload_links(op, h_links_local, true);
  if (rc) return rc;
  if (flags & (GLB_STAG_NORMAL | GLB_STAG_M2MDEODOE)) GLB_CUDA(cudaMalloc(&op->tmp, sizeof(cplx) * (size_t)X * op->Yloc));
  return alloc_ghosts(op, 2);
}

int create_gamma5_impl(glb_context* ctx, int X, int Y, glb_operator** out) {
  return new_op(ctx, OPK_GAMMA5, GLB_COMPLEX, X, Y, 1, out);
}

int create_stencil2d_impl(glb_context* ctx, const void* clover, const void* hopping, const void* two_link, int X,
                            int Y, int nc, const double shift[2], const double eo_shift[2],
                            const double dof_shift[2], glb_operator** out) {
  if (!clover || !hopping) return fail(GLB_ERR_ARG, "stencil2d needs clover and hopping arrays");
  int rc = new_op(ctx, OPK_STENCIL, GLB_COMPLEX, X, Y, nc, out);
  if (rc) return rc;
  glb_operator* op = *out;
  op->has_two = (two_link != nullptr);
  for (int i = 0; i < 2; i++) {
    op->shift[i] = shift ? shift[i] : 0.0;
    op->eo_shift[i] = eo_shift ? eo_shift[i] : 0.0;
    op->dof_shift[i] = dof_shift ? dof_shift[i] : 0.0;
  }
  // matrices of this slab: per direction plane, rows [y0, y0+Yloc) are contiguous in the reference layout
  const size_t per_site = (size_t)nc * nc;
  const size_t Vg = (size_t)X * Y, Vl = (size_t)X * op->Yloc, off = (size_t)X * op->y0 * per_site;
  GLB_CUDA(cudaMalloc(&op->clover, sizeof(cplx) * Vl * per_site));
  GLB_CUDA(cudaMemcpyAsync(op->clover, (const cplx*)clover + off, sizeof(cplx) * Vl * per_site,
                           cudaMemcpyHostToDevice, ctx->stream));
  GLB_CUDA(cudaMalloc(&op->hopping, sizeof(cplx) * 4 * Vl * per_site));
  for (int d = 0; d < 4; d++)
    GLB_CUDA(cudaMemcpyAsync(op->hopping + d * Vl * per_site, (const cplx*)hopping + d * Vg * per_site + off,
                             sizeof(cplx) * Vl * per_site, cudaMemcpyHostToDevice, ctx->stream));
  if (op->has_two) {
    GLB_CUDA(cudaMalloc(&op->two_link, sizeof(cplx) * 8 * Vl * per_site));
    for (int d = 0; d < 8; d++)
      GLB_CUDA(cudaMemcpyAsync(op->two_link + d * Vl * per_site, (const cplx*)two_link + d * Vg * per_site + off,
                               sizeof(cplx) * Vl * per_site, cudaMemcpyHostToDevice, ctx->stream));
  }
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  return alloc_ghosts(op, op->has_two ? 2 : 1);
}

}  // namespace

// on any error path the half-built operator is destroyed and *out set to null (a caller that checks the return
// code and then frees the handle must not be handed a leaked object)
template <typename F>
static int guarded_create(glb_operator** out, F body) {
  if (out) *out = nullptr;
  const int rc = body();
  if (rc != GLB_OK && out && *out) {
    glb_op_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

extern "C" {

int glb_op_create_laplace(glb_context* ctx, int dtype, int X, int Y, int Nc, double diag_re, double diag_im,
                          glb_operator** out) {
  return guarded_create(out, [&] { return create_laplace_impl(ctx, dtype, X, Y, Nc, diag_re, diag_im, out); });
}
int glb_op_create_staggered_free_real(glb_context* ctx, int X, int Y, double mass, glb_operator** out) {
  return guarded_create(out, [&] { return create_staggered_free_real_impl(ctx, X, Y, mass, out); });
}
int glb_op_create_laplace_u1(glb_context* ctx, const void* h_links, int X, int Y, double mass, glb_operator** out) {
  return guarded_create(out, [&] { return create_laplace_u1_impl(ctx, h_links, X, Y, mass, out); });
}
int glb_op_create_staggered(glb_context* ctx, const void* h_links, int X, int Y, double mass, unsigned flags,
                            glb_operator** out) {
  return guarded_create(out, [&] { return create_staggered_impl(ctx, h_links, X, Y, mass, flags, out); });
}
int glb_op_create_staggered_local(glb_context* ctx, const void* h_links_local, int X, int Y, double mass,
                                  unsigned flags, glb_operator** out) {
  return guarded_create(out, [&] { return create_staggered_local_impl(ctx, h_links_local, X, Y, mass, flags, out); });
}
int glb_op_create_gamma5(glb_context* ctx, int X, int Y, glb_operator** out) {
  return guarded_create(out, [&] { return create_gamma5_impl(ctx, X, Y, out); });
}
int glb_op_create_stencil2d(glb_context* ctx, const void* clover, const void* hopping, const void* two_link, int X,
                            int Y, int nc, const double shift[2], const double eo_shift[2],
                            const double dof_shift[2], glb_operator** out) {
  return guarded_create(out, [&] {
    return create_stencil2d_impl(ctx, clover, hopping, two_link, X, Y, nc, shift, eo_shift, dof_shift, out);
  });
}

}  // extern "C"

namespace glb {
// a five-point stencil2d operator around matrices that already live on the device (ownership passes to the
// operator): used by the Galerkin set-up in mg.cu.  Y is the GLOBAL extent; the arrays hold this rank's Yloc rows
// in the slab-local plane layout of glb_op_create_stencil2d.
int op_adopt_stencil2d(glb_context* ctx, int X, int Y, int Yloc, int nc, cplx* d_clover, cplx* d_hopping, glb_operator** out) {
  int rc = new_op(ctx, OPK_STENCIL, GLB_COMPLEX, X, Y, nc, out);
  if (rc) return rc;
  glb_operator* op = *out;
  if (op->Yloc != Yloc) {
    delete op;
    *out = nullptr;
    return fail(GLB_ERR_ARG, "device-side stencil set-up: the coarse slab does not match this rank's share of the coarse lattice");
  }
  op->has_two = false;
  rc = alloc_ghosts(op, 1);
  if (rc) {
    glb_op_destroy(op);   // clover / hopping are still null: the caller keeps its arrays
    *out = nullptr;
    return rc;
  }
  op->clover = d_clover;
  op->hopping = d_hopping;
  return GLB_OK;
}
}  // namespace glb

extern "C" {

int glb_slab_bounds(glb_context* ctx, int Y, int* y0, int* Yloc) {
  if (!ctx || !y0 || !Yloc) return fail(GLB_ERR_ARG, "glb_slab_bounds: null argument");
  slab_of(ctx, Y, y0, Yloc);
  return GLB_OK;
}

int glb_op_destroy(glb_operator* op) {
  if (!op) return GLB_OK;
  cudaStreamSynchronize(op->ctx->stream);
  cudaFree(op->Ux_store);
  cudaFree(op->Uy_store);
  cudaFree(op->tmp);
  if (!op->ghost_p2p) {
    cudaFree(op->ghost_lo);
    cudaFree(op->ghost_hi);
  } else {
    comm_arena_free(op->ctx, op->ghost_off, op->ghost_arena_bytes, op->halo_seq);  // every rank in the same order
  }
  if (op->cs_ready) comm_arena_free(op->ctx, op->cs_off, 4 * ((size_t)3 * 2 * op->X * sizeof(cplx)) + 256, op->cs_seq);
  cudaFree(op->send_lo);
  cudaFree(op->send_hi);
  cudaFree(op->tmp2);
  if (op->base) {  // a view: the matrices are the base's
    if (op->owns_base) glb_op_destroy(op->base);
  } else {
    cudaFree(op->clover);
    cudaFree(op->hopping);
    cudaFree(op->two_link);
  }
  delete op;
  return GLB_OK;
}

int glb_op_set_mass(glb_operator* op, double mass) {
  op->mass = mass;
  if (op->kind == OPK_LAPLACE && (op->flags & 0x100u)) op->diag_re = mass;
  return GLB_OK;
}
int glb_op_set_shifts(glb_operator* op, const double shift[2], const double eo_shift[2], const double dof_shift[2]) {
  if (!op || op->kind != OPK_STENCIL) return fail(GLB_ERR_ARG, "glb_op_set_shifts: not a stencil2d operator");
  if (op->base) return glb_op_set_shifts(op->base, shift, eo_shift, dof_shift);
  for (int i = 0; i < 2; i++) {
    if (shift) op->shift[i] = shift[i];
    if (eo_shift) op->eo_shift[i] = eo_shift[i];
    if (dof_shift) op->dof_shift[i] = dof_shift[i];
  }
  return GLB_OK;
}
int glb_op_get_shifts(const glb_operator* op, double shift[2], double eo_shift[2], double dof_shift[2]) {
  if (!op || op->kind != OPK_STENCIL) return fail(GLB_ERR_ARG, "glb_op_get_shifts: not a stencil2d operator");
  if (op->base) return glb_op_get_shifts(op->base, shift, eo_shift, dof_shift);
  for (int i = 0; i < 2; i++) {
    if (shift) shift[i] = op->shift[i];
    if (eo_shift) eo_shift[i] = op->eo_shift[i];
    if (dof_shift) dof_shift[i] = op->dof_shift[i];
  }
  return GLB_OK;
}
int glb_op_stencil_download(glb_operator* op, void* h_clover, void* h_hopping) {
  if (!op || op->kind != OPK_STENCIL) return fail(GLB_ERR_ARG, "glb_op_stencil_download: not a stencil2d operator");
  const glb_operator* src = op->base ? op->base : op;  // a view shows its base's matrices
  GLB_CUDA(cudaSetDevice(op->ctx->device));
  const size_t per = (size_t)op->X * op->Yloc * op->nc * op->nc * sizeof(cplx);  // this rank's rows of every plane
  if (h_clover) GLB_CUDA(cudaMemcpyAsync(h_clover, src->clover, per, cudaMemcpyDeviceToHost, op->ctx->stream));
  if (h_hopping) GLB_CUDA(cudaMemcpyAsync(h_hopping, src->hopping, 4 * per, cudaMemcpyDeviceToHost, op->ctx->stream));
  GLB_CUDA(cudaStreamSynchronize(op->ctx->stream));
  return GLB_OK;
}
int glb_op_dtype(const glb_operator* op) { return op->dtype; }
size_t glb_op_local_size(const glb_operator* op) { return (size_t)op->X * op->Yloc * op->nc; }
size_t glb_op_global_size(const glb_operator* op) { return (size_t)op->X * op->Y * op->nc; }
glb_context* glb_op_context(const glb_operator* op) { return op->ctx; }

double glb_op_bytes_per_apply(const glb_operator* op) {
  // distinct elements read once / written once
  const double V = (double)op->X * op->Yloc;
  const double e = (double)elem_bytes(op->dtype);
  switch (op->kind) {
    case OPK_LAPLACE: return V * op->nc * 2 * e;
    case OPK_LAPLACE_U1: return V * 64.0;
    case OPK_STAGGERED: {
      const double one = op->has_links ? 64.0 : 32.0;
      if (op->flags & GLB_STAG_M2MDEODOE) return V * (2 * one + 16.0);  // two hopping passes + the m^2 in term
      return V * ((op->flags & GLB_STAG_NORMAL) ? 2 * one : one);
    }
    case OPK_GAMMA5: return V * 32.0;
    case OPK_STENCIL: {
      const double nc = op->nc;
      return V * ((op->has_two ? 13.0 : 5.0) * nc * nc + 2 * nc) * 16.0;
    }
  }
  return 0.0;
}

}  // extern "C"

namespace glb {

// one operator application with optional fused reductions; halo rows exchanged first on slabs
// A view shares the base's matrices and shifts; they are read at apply time so that glb_op_set_shifts on the base is seen.
static void sync_view(glb_operator* v) {
  const glb_operator* b = v->base;
  v->clover = b->clover;
  v->hopping = b->hopping;
  v->two_link = b->two_link;
  v->has_two = b->has_two;
  for (int i = 0; i < 2; i++) {
    v->shift[i] = b->shift[i];
    v->eo_shift[i] = b->eo_shift[i];
    v->dof_shift[i] = b->dof_shift[i];
  }
}

static int part_apply(glb_operator* op, void* out, const void* in, int part, int post = 0, const double coef[2] = nullptr,
                      const void* aux = nullptr) {
  int rc = halo_exchange(op, in, op->has_two ? 2 : 1);
  if (rc) return rc;
  return launch_stencil2d_part(op, out, in, part, post, coef, aux);
}

// the composite stencil operators of operators_stencil.cpp:196-214 and mg_complex.cpp:1228-1372
static int apply_composite(glb_operator* op, void* out, const void* in) {
  sync_view(op);
  // shift*shift as std::complex multiplies it (operators_stencil.cpp:209: stenc->shift*stenc->shift*rhs[i])
  const double s2[2] = {op->shift[0] * op->shift[0] - op->shift[1] * op->shift[1],
                        op->shift[0] * op->shift[1] + op->shift[1] * op->shift[0]};
  int rc;
  switch (op->composite) {
    case GLB_SV_M2MDEODOE:  // tmp = D_oe in ; out = m^2 in - D_eo tmp on even sites, 0 on odd
      if ((rc = part_apply(op, op->tmp, in, GLB_PART_OE))) return rc;
      return part_apply(op, out, op->tmp, GLB_PART_EO, 1, s2, in);
    case GLB_SV_M2MDTBDBT:
      if ((rc = part_apply(op, op->tmp, in, GLB_PART_BT))) return rc;
      return part_apply(op, out, op->tmp, GLB_PART_TB, 1, s2, in);
    case GLB_SV_NORMAL_EO:  // out = m^2 in - D_oe D_eo in - D_eo D_oe in
      if ((rc = part_apply(op, op->tmp, in, GLB_PART_EO))) return rc;
      if ((rc = part_apply(op, out, op->tmp, GLB_PART_OE, 3, s2, in))) return rc;
      if ((rc = part_apply(op, op->tmp, in, GLB_PART_OE))) return rc;
      return part_apply(op, out, op->tmp, GLB_PART_EO, 4);
    case GLB_SV_NORMAL_TB:
      if ((rc = part_apply(op, op->tmp, in, GLB_PART_TB))) return rc;
      if ((rc = part_apply(op, out, op->tmp, GLB_PART_BT, 3, s2, in))) return rc;
      if ((rc = part_apply(op, op->tmp, in, GLB_PART_BT))) return rc;
      return part_apply(op, out, op->tmp, GLB_PART_TB, 4);
    case GLB_SV_DAGGER_EO:  // epsilon D epsilon
    case GLB_SV_DAGGER_TB: {  // sigma_3 D sigma_3
      const int mode = (op->composite == GLB_SV_DAGGER_EO) ? 0 : 1;
      ApplyFusion none;
      if ((rc = launch_stencil2d_sign(op, op->tmp, in, mode))) return rc;
      if ((rc = halo_exchange(op, op->tmp, op->has_two ? 2 : 1))) return rc;
      if ((rc = launch_stencil2d(op, op->tmp2, op->tmp, none))) return rc;
      return launch_stencil2d_sign(op, out, op->tmp2, mode);
    }
  }
  return fail(GLB_ERR_ARG, "unknown composite stencil operator");
}

static int apply_impl(glb_operator* op, void* out, const void* in, const ApplyFusion& f) {
  glb_context* ctx = op->ctx;
  if (out == in) return fail(GLB_ERR_ARG, "apply: output must not alias input");
  int rc;
  switch (op->kind) {
    case OPK_LAPLACE:
      if ((rc = halo_exchange(op, in, 1))) return rc;
      return launch_laplace(op, out, in, f);
    case OPK_GAMMA5:
      if (f.w || f.w_is_input) return fail(GLB_ERR_ARG, "gamma5 has no fused reductions");
      return launch_gamma5(op, out, in);
    case OPK_STENCIL:
      if (op->composite) {
        if (f.w || f.w_is_input || f.want_norm || f.p_new || f.cg_role)
          return fail(GLB_ERR_ARG, "composite stencil operators have no fused reductions");
        return apply_composite(op, out, in);
      }
      if ((rc = halo_exchange(op, in, op->has_two ? 2 : 1))) return rc;
      return launch_stencil2d(op, out, in, f);
    case OPK_LAPLACE_U1:
    case OPK_STAGGERED: {
      if (op->flags & GLB_STAG_NORMAL) {  // operators.cpp:444-453 : tmp = D in ; out = D^dag tmp
        if (normal_fused_ok(op)) {        // one pass, tmp stays on the SM; slabs exchange two rows once
          ApplyFusion g1 = f;
          if ((rc = halo_exchange(op, in, 2, &g1.wait))) return rc;  // the kernel waits for the flags in its prologue
          return launch_normal(op, out, in, g1);
        }
        ApplyFusion none;
        if ((rc = halo_exchange(op, in, 1, &none.wait))) return rc;
        if ((rc = launch_staggered(op, op->tmp, in, false, none))) return rc;
        ApplyFusion g = f;
        if ((rc = halo_exchange(op, op->tmp, 1, &g.wait))) return rc;
        if (g.w_is_input) {  // the dot partner is the ORIGINAL input, not tmp
          g.w_is_input = false;
          g.w = in;
        }
        return launch_staggered(op, out, op->tmp, true, g);
      }
      if (op->flags & GLB_STAG_M2MDEODOE) {  // operators.cpp:549-571 : tmp = D_oe in ; out = m^2 in - D_eo tmp | 0
        ApplyFusion none;
        if ((rc = halo_exchange(op, in, 1))) return rc;
        if ((rc = launch_staggered_eo(op, op->tmp, in, 1, 0, 0.0, nullptr, none))) return rc;
        if ((rc = halo_exchange(op, op->tmp, 1))) return rc;
        ApplyFusion g = f;
        if (g.w_is_input) {
          g.w_is_input = false;
          g.w = in;
        }
        return launch_staggered_eo(op, out, op->tmp, 0, 1, op->mass * op->mass, in, g);
      }
      ApplyFusion g2 = f;
      if ((rc = halo_exchange(op, in, 1, &g2.wait))) return rc;  // only the boundary row blocks wait, inside the kernel
      if (op->flags & GLB_STAG_DEO) return launch_staggered_eo(op, out, in, 0, 0, 0.0, nullptr, g2);
      if (op->flags & GLB_STAG_DOE) return launch_staggered_eo(op, out, in, 1, 0, 0.0, nullptr, g2);
      return launch_staggered(op, out, in, (op->flags & GLB_STAG_DAGGER) != 0, g2);
    }
  }
  (void)ctx;
  return fail(GLB_ERR_ARG, "unknown operator kind");
}

int op_apply_fused(glb_operator* op, void* out, const void* in, const ApplyFusion& f) { return apply_impl(op, out, in, f); }

}  // namespace glb

extern "C" {

int glb_op_apply(glb_operator* op, void* d_out, const void* d_in) {
  ApplyFusion none;
  return apply_impl(op, d_out, d_in, none);
}

int glb_op_apply_part(glb_operator* op, void* d_out, const void* d_in, int part) {
  if (!op || op->kind != OPK_STENCIL) return fail(GLB_ERR_ARG, "glb_op_apply_part needs a stencil2d operator");
  if (d_out == d_in) return fail(GLB_ERR_ARG, "apply: output must not alias input");
  if (op->base) sync_view(op);
  return part_apply(op, d_out, d_in, part);
}

int glb_op_create_stencil_view(glb_operator* base, int kind, int adopt_base, glb_operator** out) {
  if (!base || !out || base->kind != OPK_STENCIL || base->base)
    return fail(GLB_ERR_ARG, "glb_op_create_stencil_view: the base must be a plain stencil2d operator");
  if (kind < GLB_SV_M2MDEODOE || kind > GLB_SV_DAGGER_TB) return fail(GLB_ERR_ARG, "glb_op_create_stencil_view: unknown kind");
  if ((kind == GLB_SV_M2MDTBDBT || kind == GLB_SV_NORMAL_TB) && base->nc % 2 != 0)
    return fail(GLB_ERR_ARG, "glb_op_create_stencil_view: top/bottom operators need an even number of colours");
  glb_context* ctx = base->ctx;
  GLB_CUDA(cudaSetDevice(ctx->device));
  int rc = new_op(ctx, OPK_STENCIL, GLB_COMPLEX, base->X, base->Y, base->nc, out);
  if (rc) return rc;
  glb_operator* v = *out;
  v->composite = kind;
  v->base = base;
  v->owns_base = false;  // set last: a failed creation must not take the base with it
  sync_view(v);
  const size_t bytes = sizeof(cplx) * (size_t)v->X * v->Yloc * v->nc;
  if (cudaMalloc(&v->tmp, bytes) != cudaSuccess ||
      ((kind == GLB_SV_DAGGER_EO || kind == GLB_SV_DAGGER_TB) && cudaMalloc(&v->tmp2, bytes) != cudaSuccess)) {
    glb_op_destroy(v);
    *out = nullptr;
    return fail(GLB_ERR_CUDA, "glb_op_create_stencil_view: out of device memory");
  }
  rc = alloc_ghosts(v, v->has_two ? 2 : 1);
  if (rc) {
    glb_op_destroy(v);
    *out = nullptr;
    return rc;
  }
  v->owns_base = adopt_base != 0;
  return GLB_OK;
}

int glb_stencil_prec_prepare(glb_operator* op, int top_bottom, void* d_rhs_part, const void* d_rhs_orig) {
  if (!op || op->kind != OPK_STENCIL || op->composite) return fail(GLB_ERR_ARG, "glb_stencil_prec_prepare needs a plain stencil2d operator");
  if (d_rhs_part == d_rhs_orig) return fail(GLB_ERR_ARG, "glb_stencil_prec_prepare: output must not alias input");
  if (top_bottom && op->nc % 2 != 0) return fail(GLB_ERR_ARG, "glb_stencil_prec_prepare: top/bottom needs an even number of colours");
  return part_apply(op, d_rhs_part, d_rhs_orig, top_bottom ? GLB_PART_TB : GLB_PART_EO, 1, op->shift, d_rhs_orig);
}

int glb_stencil_prec_reconstruct(glb_operator* op, int top_bottom, void* d_lhs_full, const void* d_lhs_part,
                                 const void* d_rhs_other) {
  if (!op || op->kind != OPK_STENCIL || op->composite) return fail(GLB_ERR_ARG, "glb_stencil_prec_reconstruct needs a plain stencil2d operator");
  if (d_lhs_full == d_lhs_part || d_lhs_full == d_rhs_other)
    return fail(GLB_ERR_ARG, "glb_stencil_prec_reconstruct: output must not alias an input");
  if (top_bottom && op->nc % 2 != 0) return fail(GLB_ERR_ARG, "glb_stencil_prec_reconstruct: top/bottom needs an even number of colours");
  const double inv[2] = {1.0 / op->shift[0], 0.0};  // operators_stencil.cpp:222: inv_mass = 1.0/real(stenc->shift)
  return part_apply(op, d_lhs_full, d_lhs_part, top_bottom ? GLB_PART_BT : GLB_PART_OE, 2, inv, d_rhs_other);
}

int glb_stag_eoprec_prepare(glb_operator* op, void* d_rhs_e, const void* d_rhs_orig) {
  if (!op || op->kind != OPK_STAGGERED || !op->has_links) return fail(GLB_ERR_ARG, "eoprec_prepare needs a gauged staggered operator");
  if (d_rhs_e == d_rhs_orig) return fail(GLB_ERR_ARG, "eoprec_prepare: output must not alias input");
  ApplyFusion none;
  int rc = halo_exchange(op, d_rhs_orig, 1);
  if (rc) return rc;
  return launch_staggered_eo(op, d_rhs_e, d_rhs_orig, 0, 1, op->mass, d_rhs_orig, none);
}

int glb_stag_eoprec_reconstruct(glb_operator* op, void* d_lhs_full, const void* d_lhs_e, const void* d_rhs_o) {
  if (!op || op->kind != OPK_STAGGERED || !op->has_links) return fail(GLB_ERR_ARG, "eoprec_reconstruct needs a gauged staggered operator");
  if (d_lhs_full == d_lhs_e || d_lhs_full == d_rhs_o) return fail(GLB_ERR_ARG, "eoprec_reconstruct: output must not alias an input");
  ApplyFusion none;
  int rc = halo_exchange(op, d_lhs_e, 1);
  if (rc) return rc;
  return launch_staggered_eo(op, d_lhs_full, d_lhs_e, 1, 2, 1.0 / op->mass, d_rhs_o, none);
}

int glb_op_apply_dot(glb_operator* op, void* d_out, const void* d_in, const void* d_w, int want_norm, double dots[3]) {
  glb_context* ctx = op->ctx;
  if (op->composite || op->kind == OPK_GAMMA5) {  // no fused epilogue: the reductions follow as their own passes
    int rc0 = op->composite ? apply_composite(op, d_out, d_in) : launch_gamma5(op, d_out, d_in);
    if (rc0) return rc0;
    const size_t n = glb_op_local_size(op);
    rc0 = glb_dot(ctx, op->dtype, n, d_w ? d_w : d_in, d_out, dots);
    if (rc0) return rc0;
    dots[2] = 0.0;
    return want_norm ? glb_norm2sq(ctx, op->dtype, n, d_out, &dots[2]) : GLB_OK;
  }
  ApplyFusion f;
  f.w = d_w ? d_w : d_in;
  f.w_is_input = (d_w == nullptr || d_w == d_in);
  f.want_norm = want_norm != 0;
  f.to_host = true;
  int rc = apply_impl(op, d_out, d_in, f);
  if (rc) return rc;
  GLB_CUDA(cudaStreamSynchronize(ctx->stream));
  const bool cx = (op->dtype == GLB_COMPLEX);
  double tmp[3] = {0, 0, 0};
  const int k = (cx ? 2 : 1) + (want_norm ? 1 : 0);
  for (int i = 0; i < k; i++) tmp[i] = ctx->result_host_ptr[i];
  if (ctx->nranks > 1) {
    rc = allreduce_sum(ctx, tmp, k);
    if (rc) return rc;
  }
  if (cx) {
    dots[0] = tmp[0], dots[1] = tmp[1], dots[2] = want_norm ? tmp[2] : 0.0;
  } else {
    dots[0] = tmp[0], dots[1] = 0.0, dots[2] = want_norm ? tmp[1] : 0.0;
  }
  return GLB_OK;
}

}  // extern "C"
