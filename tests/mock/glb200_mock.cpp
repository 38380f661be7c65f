// glb200_mock.cpp -- TEST INFRASTRUCTURE: a host-memory implementation of the C ABI
// (include/glb200.h) so that the C++ solver shells (generic-linalg_b200/host/*.cpp) can be
// exercised WITHOUT a GPU in `pytest -m "not gpu"`.  "Device" vectors are malloc'ed host arrays,
// reductions are the serial sums of generic_vector.h and operators are the CPU oracle's
// (oracle/port_oracle.cpp, included below) -- so the shells running on this mock must reproduce
// the reference's solvers BIT FOR BIT, which pins every piece of host logic (scalar recurrences,
// stopping tests, counting, permutations).  Never shipped, never loaded by the product.
#include "../../oracle/port_oracle.cpp"   // anonymous-namespace helpers: PortOp, apply(), v_dot, ...

#include <cstdlib>

#include "../../include/glb200.h"

struct glb_context {
  int dummy;
};
struct glb_operator {
  glb_context* ctx;
  PortOp* op;
  int dtype;
  std::vector<cplx> links;
  int composite;          // GLB_SV_* view of `base` (op is then the base's PortOp)
  glb_operator* base;
  bool owns_base;
  glb_operator() : ctx(0), op(0), dtype(0), composite(0), base(0), owns_base(false) {}
};

static std::string g_err;
static unsigned long long g_calls = 0;
static glb_context g_ctx;

template <typename T>
static T cf(const double a[2]);
template <>
double cf<double>(const double a[2]) { return a[0]; }
template <>
cplx cf<cplx>(const double a[2]) { return cplx(a[0], a[1]); }

#define BOTH(dtype, BODY)                      \
  do {                                         \
    if ((dtype) == GLB_COMPLEX) {              \
      typedef cplx T;                          \
      BODY                                     \
    } else {                                   \
      typedef double T;                        \
      BODY                                     \
    }                                          \
  } while (0)

static void pk(double a, double o[2]) { o[0] = a; o[1] = 0.0; }
static void pk(const cplx& a, double o[2]) { o[0] = a.real(); o[1] = a.imag(); }
static double cj(double a) { return a; }
static cplx cj(const cplx& a) { return std::conj(a); }

template <typename T>
static int mock_krylov(glb_operator* op, int alg, T* x, const T* b, int max_iter, double eps, glb_cg_report* rep,
                       double* hist, int hist_cap) {
  glb_context* c = op->ctx;
  const int dt = op->dtype;
  const size_t n = op->op->size;
  std::vector<T> v[6];
  for (int i = 0; i < 6; i++) v[i].assign(n, T(0));
  double bsq = 0.0;
  glb_norm2sq(c, dt, n, b, &bsq);
  const double bnorm = std::sqrt(bsq);
  int iter = 0, hit_max = 0;
  double rsq = 0.0;
  auto stop = [&](int k) {  // the epilogue of the x / r update kernel
    iter = k + 1;
    if (hist && k < hist_cap) hist[k] = rsq;
    const bool conv = std::sqrt(rsq) < eps * bnorm, last = (k == max_iter - 1);
    if (conv || last) hit_max = last ? 1 : 0;
    return conv || last;
  };
  if (alg == GLB_KRYLOV_BICGSTAB) {
    T *r = v[0].data(), *r0 = v[1].data(), *p = v[2].data(), *Ap = v[3].data(), *s = v[4].data(), *As = v[5].data();
    double d[3], o3[3], ca[2], co[2], cb[2];
    glb_op_apply(op, Ap, x);
    glb_sub(c, dt, n, b, Ap, r);
    glb_vec_copy(c, dt, n, r0, r);
    glb_vec_copy(c, dt, n, p, r);
    glb_dot(c, dt, n, r0, r, d);
    T rho = cf<T>(d);
    glb_op_apply_dot(op, Ap, p, r0, 0, d);
    T r0Ap = cf<T>(d);
    for (int k = 0; k < max_iter; k++) {
      const T alpha = rho / r0Ap;                       // prologue of the s kernel
      const T nalpha = -alpha;
      pk(nalpha, ca);
      glb_axpyz(c, dt, n, ca, Ap, r, s);                // s = r - alpha Ap
      glb_op_apply_dot(op, As, s, s, 1, d);             // As = A s ; <s,As>, |As|^2
      const T omega = cj(cf<T>(d)) / T(d[2]);           // prologue of the x / r update
      pk(alpha, ca);
      pk(omega, co);
      glb_bicgstab_update(c, dt, n, ca, p, co, s, As, r0, x, r, o3);
      rsq = o3[0];
      if (stop(k)) break;
      const T rhoNew = cf<T>(o3 + 1);                   // its epilogue
      const T beta = rhoNew / rho * (alpha / omega);
      rho = rhoNew;
      pk(beta, cb);
      glb_bicgstab_pupdate(c, dt, n, r, cb, co, Ap, p);
      glb_op_apply_dot(op, Ap, p, r0, 0, d);            // Ap = A p ; <r0,Ap>
      r0Ap = cf<T>(d);
    }
    rep->ops = 2 + iter + (iter > 0 ? iter - 1 : 0);
  } else {
    T *r = v[0].data(), *Ar = v[1].data(), *p = v[2].data(), *Ap = v[3].data();
    double d[3], ca[2], cna[2], cb[2], apsq = 0.0;
    glb_op_apply(op, p, x);
    glb_sub(c, dt, n, b, p, r);
    glb_vec_copy(c, dt, n, p, r);
    glb_op_apply(op, Ap, p);
    glb_vec_copy(c, dt, n, Ar, Ap);
    glb_norm2sq(c, dt, n, Ap, &apsq);
    glb_dot(c, dt, n, Ap, r, d);
    T Apr = cf<T>(d);
    for (int k = 0; k < max_iter; k++) {
      const T alpha = Apr / apsq;                       // prologue of the x / r update
      const T nalpha = -alpha;
      pk(alpha, ca);
      pk(nalpha, cna);
      glb_update_xr_norm(c, dt, n, ca, p, x, cna, Ap, r, &rsq);
      if (stop(k)) break;
      glb_op_apply_dot(op, Ar, r, Ap, 0, d);            // Ar = A r ; <Ap,Ar>
      const T beta = -cf<T>(d) / apsq;                  // prologue of the p / Ap update
      pk(beta, cb);
      glb_update_p_ap_norm(c, dt, n, r, Ar, cb, p, Ap, &apsq);
      glb_dot(c, dt, n, Ap, r, d);                      // ... which also sums <Ap,r> for the next iteration
      Apr = cf<T>(d);
    }
    rep->ops = 2 + (iter > 0 ? iter - 1 : 0);
  }
  rep->iterations = iter;
  rep->hit_max_iter = hit_max;
  rep->rsq = rsq;
  rep->bnorm = bnorm;
  return GLB_OK;
}

template <typename T>
static int mock_cg(glb_operator* op, T* x, const T* b, int max_iter, double eps, glb_cg_report* rep, double* hist,
                   int hist_cap) {
  glb_context* c = op->ctx;
  const int dt = op->dtype;
  const size_t n = op->op->size;
  std::vector<T> r(n), p(n), Ap(n);
  double bsq = 0.0, rsq = 0.0, rsqNew = 0.0, d[3], ca[2], cna[2], cb[2];
  glb_norm2sq(c, dt, n, b, &bsq);
  const double bnorm = std::sqrt(bsq);
  glb_op_apply(op, p.data(), x);
  glb_sub(c, dt, n, b, p.data(), r.data());
  glb_vec_copy(c, dt, n, p.data(), r.data());
  glb_op_apply_dot(op, Ap.data(), p.data(), p.data(), 0, d);
  T pAp = cf<T>(d);
  glb_norm2sq(c, dt, n, r.data(), &rsq);
  int iter = 0, hit_max = 0;
  for (int k = 0; k < max_iter; k++) {
    const T alpha = rsq / pAp;
    const T nalpha = -alpha;
    pk(alpha, ca);
    pk(nalpha, cna);
    glb_update_xr_norm(c, dt, n, ca, p.data(), x, cna, Ap.data(), r.data(), &rsqNew);
    iter = k + 1;
    if (hist && k < hist_cap) hist[k] = rsqNew;
    const bool conv = std::sqrt(rsqNew) < eps * bnorm, last = (k == max_iter - 1);
    if (conv || last) {
      hit_max = last ? 1 : 0;
      break;
    }
    const T beta = T(rsqNew / rsq);
    rsq = rsqNew;
    pk(beta, cb);
    glb_xpay(c, dt, n, r.data(), cb, p.data());
    glb_op_apply_dot(op, Ap.data(), p.data(), p.data(), 0, d);
    pAp = cf<T>(d);
  }
  rep->iterations = iter;
  rep->ops = 2 + (iter > 0 ? iter - 1 : 0);
  rep->hit_max_iter = hit_max;
  rep->rsq = rsqNew;
  rep->bnorm = bnorm;
  return GLB_OK;
}

extern "C" {
int glb_create(int, glb_context** ctx) { *ctx = &g_ctx; return GLB_OK; }
int glb_destroy(glb_context*) { return GLB_OK; }
const char* glb_last_error(void) { return g_err.c_str(); }
int glb_synchronize(glb_context*) { return GLB_OK; }
void* glb_stream(glb_context*) { return 0; }
int glb_device(glb_context*) { return -1; }
int glb_sm_count(glb_context*) { return 0; }
unsigned long long glb_kernel_launches(void) { return g_calls; }
int glb_prof_enable(glb_context*, int) { return GLB_OK; }
int glb_prof_read(glb_context*, int, int, float*, int* n) { if (n) *n = 0; return GLB_OK; }
int glb_prof_summary(glb_context*, int, int* launches, double* ms, double* bytes) {
  if (launches) *launches = 0;
  if (ms) *ms = 0.0;
  if (bytes) *bytes = 0.0;
  return GLB_OK;
}
int glb_comm_unique_id(char*) { return GLB_ERR_COMM; }
int glb_comm_init(glb_context*, int, int, const char*) { return GLB_ERR_COMM; }
int glb_comm_rank(glb_context*) { return 0; }
int glb_comm_size(glb_context*) { return 1; }
int glb_comm_p2p_enabled(glb_context*) { return 0; }
int glb_comm_barrier(glb_context*) { return GLB_OK; }
int glb_slab_bounds(glb_context*, int Y, int* y0, int* Yloc) { *y0 = 0; *Yloc = Y; return GLB_OK; }

int glb_vec_alloc(glb_context*, int dtype, size_t n, void** p) {
  *p = std::calloc(n ? n : 1, dtype == GLB_COMPLEX ? 16 : 8);
  return *p ? GLB_OK : GLB_ERR_CUDA;
}
int glb_vec_free(glb_context*, void* p) { std::free(p); return GLB_OK; }
static size_t eb(int dtype) { return dtype == GLB_COMPLEX ? 16 : 8; }
int glb_vec_upload(glb_context*, int dt, size_t n, void* d, const void* s) { std::memcpy(d, s, n * eb(dt)); return GLB_OK; }
int glb_vec_download(glb_context*, int dt, size_t n, void* d, const void* s) { std::memcpy(d, s, n * eb(dt)); return GLB_OK; }
int glb_vec_zero(glb_context*, int dt, size_t n, void* d) { std::memset(d, 0, n * eb(dt)); return GLB_OK; }
int glb_vec_copy(glb_context*, int dt, size_t n, void* d, const void* s) { std::memmove(d, s, n * eb(dt)); return GLB_OK; }
int glb_host_alloc(glb_context*, size_t b, void** p) { *p = std::malloc(b ? b : 1); return GLB_OK; }
int glb_host_free(glb_context*, void* p) { std::free(p); return GLB_OK; }

static glb_operator* wrap(int kind, int X, int Y, int Nc, double mass, const void* links, int dtype) {
  glb_operator* o = new glb_operator();
  o->ctx = &g_ctx;
  o->dtype = dtype;
  orc_op_desc d;
  std::memset(&d, 0, sizeof d);
  d.kind = kind;
  d.X = X;
  d.Y = Y;
  d.Nc = Nc;
  d.mass = mass;
  if (links) {
    o->links.assign((const cplx*)links, (const cplx*)links + 2 * (size_t)X * Y);
    d.links = (const double*)o->links.data();
  }
  o->op = (PortOp*)port_op_prepare(&d);
  return o;
}
int glb_op_create_laplace(glb_context*, int dtype, int X, int Y, int Nc, double dre, double dim, glb_operator** op) {
  // mass is recovered from the diagonal exactly as the callers built it (4 + mass)
  if (dtype == GLB_REAL)
    *op = wrap(Nc == 1 && X == Y ? ORC_OP_LAPLACE_REAL : ORC_OP_LAPLACE_REAL_NC, X, Y, Nc, dre - 4, 0, dtype);
  else if (dim != 0.0)
    *op = wrap(ORC_OP_LAPLACE_IMAG, X, Y, 1, dre - 4.0, 0, dtype);
  else
    *op = wrap(ORC_OP_LAPLACE_NC, X, Y, Nc, dre - 4, 0, dtype);
  return GLB_OK;
}
int glb_op_create_staggered_free_real(glb_context*, int X, int Y, double m, glb_operator** op) {
  *op = wrap(ORC_OP_STAG_FREE_REAL, X, Y, 1, m, 0, GLB_REAL);
  return GLB_OK;
}
int glb_op_create_laplace_u1(glb_context*, const void* l, int X, int Y, double m, glb_operator** op) {
  *op = wrap(ORC_OP_LAPLACE_U1, X, Y, 1, m, l, GLB_COMPLEX);
  return GLB_OK;
}
int glb_op_create_staggered(glb_context*, const void* l, int X, int Y, double m, unsigned flags, glb_operator** op) {
  int kind = l ? ORC_OP_STAG_U1 : ORC_OP_STAG_FREE;
  if (flags & GLB_STAG_DAGGER) kind = ORC_OP_STAG_DAGGER_U1;
  if (flags & GLB_STAG_GAMMA5) kind = l ? ORC_OP_STAG_GAMMA5_U1 : ORC_OP_STAG_GAMMA5_FREE;
  if (flags & GLB_STAG_NORMAL) kind = ORC_OP_STAG_NORMAL_U1;
  if (flags & GLB_STAG_DEO) kind = ORC_OP_STAG_DEO_U1;
  if (flags & GLB_STAG_DOE) kind = ORC_OP_STAG_DOE_U1;
  if (flags & GLB_STAG_M2MDEODOE) kind = ORC_OP_STAG_M2MDEODOE_U1;
  *op = wrap(kind, X, Y, 1, m, l, GLB_COMPLEX);
  return GLB_OK;
}
int glb_op_create_staggered_local(glb_context*, const void*, int, int, double, unsigned, glb_operator**) { return GLB_ERR_ARG; }
int glb_op_create_gamma5(glb_context*, int X, int Y, glb_operator** op) {
  *op = wrap(ORC_OP_GAMMA5, X, Y, 1, 0, 0, GLB_COMPLEX);
  return GLB_OK;
}
int glb_op_create_stencil2d(glb_context*, const void* cl, const void* hp, const void* tl, int X, int Y, int nc,
                            const double sh[2], const double eo[2], const double df[2], glb_operator** op) {
  glb_operator* o = new glb_operator();
  o->ctx = &g_ctx;
  o->dtype = GLB_COMPLEX;
  orc_op_desc d;
  std::memset(&d, 0, sizeof d);
  d.kind = ORC_OP_STENCIL;
  d.X = X; d.Y = Y; d.Nc = nc;
  d.clover = (const double*)cl; d.hopping = (const double*)hp; d.two_link = (const double*)tl; d.has_two = tl != 0;
  for (int i = 0; i < 2; i++) { d.shift[i] = sh[i]; d.eo_shift[i] = eo[i]; d.dof_shift[i] = df[i]; }
  o->op = (PortOp*)port_op_prepare(&d);
  *op = o;
  return GLB_OK;
}
int glb_op_destroy(glb_operator* o) {
  if (!o) return GLB_OK;
  if (o->base) {
    if (o->owns_base) glb_op_destroy(o->base);
  } else {
    port_op_free(o->op);
  }
  delete o;
  return GLB_OK;
}
int glb_op_create_stencil_view(glb_operator* base, int kind, int adopt, glb_operator** out) {
  if (!base || base->base || kind < GLB_SV_M2MDEODOE || kind > GLB_SV_DAGGER_TB) return GLB_ERR_ARG;
  glb_operator* v = new glb_operator();
  v->ctx = base->ctx; v->op = base->op; v->dtype = base->dtype;
  v->composite = kind; v->base = base; v->owns_base = adopt != 0;
  *out = v;
  return GLB_OK;
}
// operators_stencil.cpp:179-193 / mg_complex.cpp:1211-1225
int glb_stencil_prec_prepare(glb_operator* o, int tb, void* rhs_part, const void* rhs_orig) {
  g_calls++;
  PortOp* st = o->op;
  const int n = st->size, X = st->d.X, nc = st->nc;
  cplx* out = (cplx*)rhs_part; const cplx* orig = (const cplx*)rhs_orig;
  port_stencil_apply_part(st, tb ? GLB_PART_TB : GLB_PART_EO, (double*)out, (const double*)orig);
  for (int i = 0; i < n; i++) {
    const int site = i / nc;
    const bool sel = tb ? (i % nc) < nc / 2 : ((site % X + site / X) % 2 == 0);
    if (sel) out[i] = st->shift * orig[i] - out[i];
  }
  return GLB_OK;
}
// operators_stencil.cpp:217-236 / mg_complex.cpp:1252-1272
int glb_stencil_prec_reconstruct(glb_operator* o, int tb, void* lhs_full, const void* lhs_part, const void* rhs_other) {
  g_calls++;
  PortOp* st = o->op;
  const int n = st->size, X = st->d.X, nc = st->nc;
  cplx* out = (cplx*)lhs_full; const cplx* part = (const cplx*)lhs_part; const cplx* other = (const cplx*)rhs_other;
  const double inv_mass = 1.0 / real(st->shift);
  port_stencil_apply_part(st, tb ? GLB_PART_BT : GLB_PART_OE, (double*)out, (const double*)part);
  for (int i = 0; i < n; i++) {
    const int site = i / nc;
    const bool sel = tb ? (i % nc) < nc / 2 : ((site % X + site / X) % 2 == 0);
    if (!sel) out[i] = inv_mass * (other[i] - out[i]);
    else out[i] = part[i];
  }
  return GLB_OK;
}
int glb_op_set_mass(glb_operator* o, double m) { o->op->d.mass = m; return GLB_OK; }
int glb_op_dtype(const glb_operator* o) { return o->dtype; }
size_t glb_op_local_size(const glb_operator* o) { return o->op->size; }
size_t glb_op_global_size(const glb_operator* o) { return o->op->size; }
glb_context* glb_op_context(const glb_operator* o) { return o->ctx; }
double glb_op_bytes_per_apply(const glb_operator*) { return 0; }
int glb_op_apply_part(glb_operator* o, void* out, const void* in, int part) {
  g_calls++;
  port_stencil_apply_part(o->op, part, (double*)out, (const double*)in);
  return GLB_OK;
}
int glb_stag_eoprec_prepare(glb_operator* o, void* rhs_e, const void* rhs_orig) {
  g_calls++;
  port_eoprec_prepare(o->op, (double*)rhs_e, (const double*)rhs_orig);
  return GLB_OK;
}
int glb_stag_eoprec_reconstruct(glb_operator* o, void* lhs_full, const void* lhs_e, const void* rhs_o) {
  g_calls++;
  port_eoprec_reconstruct(o->op, (double*)lhs_full, (const double*)lhs_e, (const double*)rhs_o);
  return GLB_OK;
}
// the composite stencil operators in the reference's statement order (operators_stencil.cpp:196-214,
// mg_complex.cpp:1228-1372)
static void mock_composite(glb_operator* o, cplx* lhs, const cplx* rhs) {
  PortOp* st = o->op;
  const int n = st->size, X = st->d.X, nc = st->nc;
  std::vector<cplx> tmp(n), tmp2(n);
  const int kind = o->composite;
  const bool tb = (kind == GLB_SV_M2MDTBDBT || kind == GLB_SV_NORMAL_TB || kind == GLB_SV_DAGGER_TB);
  const int A = tb ? GLB_PART_TB : GLB_PART_EO, B = tb ? GLB_PART_BT : GLB_PART_OE;
  if (kind == GLB_SV_M2MDEODOE || kind == GLB_SV_M2MDTBDBT) {
    port_stencil_apply_part(st, B, (double*)tmp.data(), (const double*)rhs);
    port_stencil_apply_part(st, A, (double*)lhs, (const double*)tmp.data());
    for (int i = 0; i < n; i++) {
      const int site = i / nc;
      const bool sel = tb ? (i % nc) < nc / 2 : ((site % X + site / X) % 2 == 0);
      if (sel) lhs[i] = st->shift * st->shift * rhs[i] - lhs[i];
    }
  } else if (kind == GLB_SV_NORMAL_EO || kind == GLB_SV_NORMAL_TB) {
    port_stencil_apply_part(st, A, (double*)tmp2.data(), (const double*)rhs);
    port_stencil_apply_part(st, B, (double*)lhs, (const double*)tmp2.data());
    port_stencil_apply_part(st, B, (double*)tmp2.data(), (const double*)rhs);
    port_stencil_apply_part(st, A, (double*)tmp.data(), (const double*)tmp2.data());
    for (int i = 0; i < n; i++) lhs[i] = st->shift * st->shift * rhs[i] - lhs[i] - tmp[i];
  } else {
    auto sign = [&](cplx* out, const cplx* in) {
      for (int i = 0; i < n; i++) {
        const int site = i / nc;
        const bool flip = tb ? (nc % 2 == 0 && (i % nc) >= nc / 2) : ((site % X + site / X) % 2 != 0);
        out[i] = flip ? -in[i] : in[i];
      }
    };
    sign(lhs, rhs);
    port_op_apply(st, (double*)tmp.data(), (const double*)lhs);
    sign(lhs, tmp.data());
  }
}
int glb_op_apply(glb_operator* o, void* out, const void* in) {
  g_calls++;
  if (o->composite)
    mock_composite(o, (cplx*)out, (const cplx*)in);
  else
    port_op_apply(o->op, (double*)out, (const double*)in);
  return GLB_OK;
}
int glb_dot(glb_context*, int dt, size_t n, const void* x, const void* y, double out[2]) {
  port_dot(dt == GLB_COMPLEX, (const double*)x, (const double*)y, (int)n, out);
  return GLB_OK;
}
int glb_norm2sq(glb_context*, int dt, size_t n, const void* x, double* out) {
  *out = port_norm2sq(dt == GLB_COMPLEX, (const double*)x, (int)n);
  return GLB_OK;
}
int glb_diffnorm2sq(glb_context*, int dt, size_t n, const void* x, const void* y, double* out) {
  *out = port_diffnorm2sq(dt == GLB_COMPLEX, (const double*)x, (const double*)y, (int)n);
  return GLB_OK;
}
int glb_op_apply_dot(glb_operator* o, void* out, const void* in, const void* w, int want_norm, double dots[3]) {
  glb_op_apply(o, out, in);
  glb_dot(o->ctx, o->dtype, o->op->size, w ? w : in, out, dots);
  dots[2] = 0.0;
  if (want_norm) glb_norm2sq(o->ctx, o->dtype, o->op->size, out, &dots[2]);
  return GLB_OK;
}
int glb_dot_norm(glb_context* c, int dt, size_t n, const void* x, const void* y, double out[3]) {
  glb_dot(c, dt, n, x, y, out);
  return glb_norm2sq(c, dt, n, x, &out[2]);
}
int glb_multi_dot(glb_context* c, int dt, size_t n, int k, const void* const* X, const void* y, double* out) {
  for (int j = 0; j < k; j++) glb_dot(c, dt, n, X[j], y, out + 2 * j);
  return GLB_OK;
}
int glb_sub(glb_context*, int dt, size_t n, const void* a, const void* b, void* o) {
  BOTH(dt, { for (size_t i = 0; i < n; i++) ((T*)o)[i] = ((const T*)a)[i] - ((const T*)b)[i]; });
  return GLB_OK;
}
int glb_add(glb_context*, int dt, size_t n, const void* a, const void* b, void* o) {
  BOTH(dt, { for (size_t i = 0; i < n; i++) ((T*)o)[i] = ((const T*)a)[i] + ((const T*)b)[i]; });
  return GLB_OK;
}
int glb_axpy(glb_context*, int dt, size_t n, const double a[2], const void* x, void* y) {
  BOTH(dt, { const T c = cf<T>(a); for (size_t i = 0; i < n; i++) ((T*)y)[i] = ((T*)y)[i] + c * ((const T*)x)[i]; });
  return GLB_OK;
}
int glb_xpay(glb_context*, int dt, size_t n, const void* x, const double a[2], void* y) {
  BOTH(dt, { const T c = cf<T>(a); for (size_t i = 0; i < n; i++) ((T*)y)[i] = ((const T*)x)[i] + c * ((T*)y)[i]; });
  return GLB_OK;
}
int glb_axpyz(glb_context*, int dt, size_t n, const double a[2], const void* x, const void* y, void* z) {
  BOTH(dt, { const T c = cf<T>(a); for (size_t i = 0; i < n; i++) ((T*)z)[i] = ((const T*)y)[i] + c * ((const T*)x)[i]; });
  return GLB_OK;
}
int glb_rdiv(glb_context*, int dt, size_t n, const void* x, double d, void* o) {
  BOTH(dt, { for (size_t i = 0; i < n; i++) ((T*)o)[i] = ((const T*)x)[i] / d; });
  return GLB_OK;
}
int glb_axpy_norm(glb_context* c, int dt, size_t n, const double a[2], const void* x, void* y, double* nrm) {
  glb_axpy(c, dt, n, a, x, y);
  return glb_norm2sq(c, dt, n, y, nrm);
}
int glb_lincomb(glb_context*, int dt, size_t n, int k, const double* coef, const void* const* X, const void* init,
                void* out) {
  BOTH(dt, {
    for (size_t i = 0; i < n; i++) {
      T v = init ? ((const T*)init)[i] : T(0.0);
      for (int j = 0; j < k; j++) v = v + cf<T>(coef + 2 * j) * ((const T*)X[j])[i];
      ((T*)out)[i] = v;
    }
  });
  return GLB_OK;
}
int glb_update_xr_norm(glb_context* c, int dt, size_t n, const double a[2], const void* p, void* x, const double b[2],
                       const void* q, void* r, double* rsq) {
  glb_axpy(c, dt, n, a, p, x);
  glb_axpy(c, dt, n, b, q, r);
  return glb_norm2sq(c, dt, n, r, rsq);
}
int glb_update_p_ap_norm(glb_context* c, int dt, size_t n, const void* r, const void* Ar, const double beta[2], void* p,
                         void* Ap, double* apsq) {
  glb_xpay(c, dt, n, r, beta, p);
  glb_xpay(c, dt, n, Ar, beta, Ap);
  return glb_norm2sq(c, dt, n, Ap, apsq);
}
int glb_bicgstab_update(glb_context* c, int dt, size_t n, const double alpha[2], const void* p, const double omega[2],
                        const void* s, const void* As, const void* r0, void* x, void* r, double out[3]) {
  BOTH(dt, {
    const T al = cf<T>(alpha); const T om = cf<T>(omega);
    for (size_t i = 0; i < n; i++) ((T*)x)[i] = ((T*)x)[i] + al * ((const T*)p)[i] + om * ((const T*)s)[i];
    for (size_t i = 0; i < n; i++) ((T*)r)[i] = ((const T*)s)[i] - om * ((const T*)As)[i];
  });
  glb_norm2sq(c, dt, n, r, &out[0]);
  out[2] = 0.0;
  return glb_dot(c, dt, n, r0, r, out + 1);
}
int glb_bicgstab_pupdate(glb_context*, int dt, size_t n, const void* r, const double beta[2], const double omega[2],
                         const void* Ap, void* p) {
  BOTH(dt, {
    const T be = cf<T>(beta); const T om = cf<T>(omega);
    for (size_t i = 0; i < n; i++) ((T*)p)[i] = ((const T*)r)[i] + be * (((T*)p)[i] - om * ((const T*)Ap)[i]);
  });
  return GLB_OK;
}
int glb_conj(glb_context*, int dt, size_t n, const void* x, void* out) {
  if (dt == GLB_COMPLEX) {
    for (size_t i = 0; i < n; i++) ((cplx*)out)[i] = std::conj(((const cplx*)x)[i]);
  } else {
    for (size_t i = 0; i < n; i++) ((double*)out)[i] = ((const double*)x)[i];
  }
  return GLB_OK;
}
int glb_bicgstabm_update_s(glb_context*, int dt, size_t n, const double c[10], const void* r, const void* w,
                           const void* rp, void* sn) {
  BOTH(dt, {
    const T c0 = cf<T>(c); const T c1 = cf<T>(c + 2); const T c2 = cf<T>(c + 4); const T c3 = cf<T>(c + 6); const T c4 = cf<T>(c + 8);
    for (size_t i = 0; i < n; i++)
      ((T*)sn)[i] = c0 * ((const T*)r)[i] + c1 * (((T*)sn)[i] - c2 * (c3 * ((const T*)w)[i] - c4 * ((const T*)rp)[i]));
  });
  return GLB_OK;
}
int glb_cgm_update_x(glb_context*, int dt, size_t n, int ns, const double* beta_s, const void* const* p_s, void* const* x) {
  BOTH(dt, {
    for (int s = 0; s < ns; s++) {
      const T c = cf<T>(beta_s + 2 * s);
      for (size_t i = 0; i < n; i++) ((T*)x[s])[i] = ((T*)x[s])[i] - c * ((const T*)p_s[s])[i];
    }
  });
  return GLB_OK;
}
int glb_cgm_update_p(glb_context*, int dt, size_t n, int ns, const double* zeta, const double* alpha_s, const void* r,
                     void* const* p_s) {
  BOTH(dt, {
    for (int s = 0; s < ns; s++) {
      const T z = cf<T>(zeta + 2 * s); const T a = cf<T>(alpha_s + 2 * s);
      for (size_t i = 0; i < n; i++) ((T*)p_s[s])[i] = z * ((const T*)r)[i] + a * ((T*)p_s[s])[i];
    }
  });
  return GLB_OK;
}
// the device-resident CG is a CUDA-only entry point: the shells fall back to the host-scalar loop
// when forced (glb200_force_host_scalars), which is what the mock tests do.
static bool mock_krylov_on() {
  const char* e = std::getenv("GLB200_MOCK_NO_KRYLOV");
  return !(e && *e && std::atoi(e) != 0);
}
int glb_cg_solve_supported(const glb_operator* o) { return (o && !o->composite && mock_krylov_on()) ? 1 : 0; }
// glb_krylov_solve (csrc/krylov.cu) restated on the host: the SAME sequence of vector operations and the SAME scalar
// formulas the device loop evaluates in its kernel prologues / epilogues (alpha = rho / <r0,Ap>, omega =
// conj(<s,As>) / T(|As|^2), beta = rhoNew/rho * (alpha/omega); CR: alpha = <Ap,r> / |Ap|^2, beta = -<Ap,Ar> / |Ap|^2,
// <Ap,r> of the next iteration taken while p and Ap are updated), the stopping test, the iteration / operator counts
// and the residual history.  With it the device-loop branches of bicgstab_dev / cr_dev (host/dev_solvers.cpp: history
// replay for VERB_DETAIL, success flags, ops counts) run on the CPU and are compared with the reference line by line
// (tests/test_krylov_mock_cpu.py, tests/test_reference_programs_cpu.py).  GLB200_MOCK_NO_KRYLOV=1 switches it off.
int glb_krylov_solve_supported(const glb_operator* o, int alg) {
  if (!o || (alg != GLB_KRYLOV_BICGSTAB && alg != GLB_KRYLOV_CR)) return 0;
  if (o->composite) return 0;  // as the CUDA library: no fused reductions on the composite views
  return mock_krylov_on() ? 1 : 0;
}
static unsigned long long g_krylov_calls = 0;
unsigned long long glb200_mock_krylov_calls(void) { return g_krylov_calls; }  // test aid: was the device-loop branch taken?
int glb_krylov_solve(glb_operator* op, int alg, void* x, const void* b, int max_iter, double eps, glb_cg_report* rep,
                     double* hist, int hist_cap) {
  g_krylov_calls++;
  if (!op || !x || !b || !rep || max_iter < 1 || !glb_krylov_solve_supported(op, alg)) {
    g_err = "glb_krylov_solve: bad argument / not supported";
    return GLB_ERR_ARG;
  }
  if (op->dtype == GLB_COMPLEX) return mock_krylov<cplx>(op, alg, (cplx*)x, (const cplx*)b, max_iter, eps, rep, hist, hist_cap);
  return mock_krylov<double>(op, alg, (double*)x, (const double*)b, max_iter, eps, rep, hist, hist_cap);
}
int glb_krylov_graph_mode(int) { return 0; }
int glb_krylov_last_used_graph(void) { return 0; }
double glb_cg_last_pred_err(void) { return 0.0; }
int glb_cg_step_mode(int, int) { return 0; }
int glb_dbg_p2p_bench(glb_context*, int, int, double, float*) { return GLB_ERR_STATE; }
int glb_dbg_p2p_pingpong(glb_context*, int, int, float*) { return GLB_ERR_STATE; }
// glb_cg_solve restated on the host: the two-kernel device loop of csrc/cg.cu (= the reference's CG recurrence,
// generic_cg.cpp:304-354), so that the device-loop branch of cg_dev (history replay, counts, success) runs on the CPU.
// (The single-kernel iteration of cgstep.cu predicts beta's numerator one step ahead; that arithmetic is CUDA-only
// and checked on the GPU.)
int glb_cg_solve(glb_operator* op, void* x, const void* b, int max_iter, double eps, glb_cg_report* rep, double* hist,
                 int hist_cap) {
  g_krylov_calls++;
  if (!op || !x || !b || !rep || max_iter < 1 || !glb_cg_solve_supported(op)) {
    g_err = "glb_cg_solve: bad argument / not supported";
    return GLB_ERR_ARG;
  }
  if (op->dtype == GLB_COMPLEX) return mock_cg<cplx>(op, (cplx*)x, (const cplx*)b, max_iter, eps, rep, hist, hist_cap);
  return mock_cg<double>(op, (double*)x, (const double*)b, max_iter, eps, rep, hist, hist_cap);
}

// multigrid grid transfers: host loops in the reference's accumulation order (mg_complex.cpp:372-467)
struct glb_mg_transfer {
  int Xf, Yf, dof_f, bx, by, nvec, Xc, Yc;
  std::vector<std::vector<cplx> > null;
};
int glb_mg_transfer_create(glb_context*, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                           const void* const* null_vectors, glb_mg_transfer** out) {
  if (Xf % bx != 0 || Yf % by != 0) return GLB_ERR_ARG;
  glb_mg_transfer* t = new glb_mg_transfer();
  t->Xf = Xf; t->Yf = Yf; t->dof_f = dof_f; t->bx = bx; t->by = by; t->nvec = nvec;
  t->Xc = Xf / bx; t->Yc = Yf / by;
  const size_t nf = (size_t)Xf * Yf * dof_f;
  for (int v = 0; v < nvec; v++) t->null.push_back(std::vector<cplx>((const cplx*)null_vectors[v], (const cplx*)null_vectors[v] + nf));
  *out = t;
  return GLB_OK;
}
int glb_mg_transfer_destroy(glb_mg_transfer* t) { delete t; return GLB_OK; }
size_t glb_mg_fine_size(const glb_mg_transfer* t) { return (size_t)t->Xf * t->Yf * t->dof_f; }
size_t glb_mg_coarse_size(const glb_mg_transfer* t) { return (size_t)t->Xc * t->Yc * t->nvec; }
int glb_mg_prolong(glb_mg_transfer* t, void* d_fine, const void* d_coarse) {
  g_calls++;
  cplx* fine = (cplx*)d_fine; const cplx* coarse = (const cplx*)d_coarse;
  const size_t nf = glb_mg_fine_size(t);
  for (size_t f = 0; f < nf; f++) {
    const size_t site = f / t->dof_f;
    const int x = (int)(site % t->Xf), y = (int)(site / t->Xf);
    const size_t cs = (size_t)(y / t->by) * t->Xc + x / t->bx;
    cplx acc = 0.0;
    for (int v = 0; v < t->nvec; v++) acc += t->null[v][f] * coarse[cs * t->nvec + v];
    fine[f] = acc;
  }
  return GLB_OK;
}
int glb_mg_restrict(glb_mg_transfer* t, void* d_coarse, const void* d_fine) {
  g_calls++;
  cplx* coarse = (cplx*)d_coarse; const cplx* fine = (const cplx*)d_fine;
  const size_t nc = glb_mg_coarse_size(t);
  for (size_t i = 0; i < nc; i++) {
    const int v = (int)(i % t->nvec);
    const size_t cs = i / t->nvec;
    const int xc = (int)(cs % t->Xc), yc = (int)(cs / t->Xc);
    cplx acc = 0.0;
    for (int y = yc * t->by; y < (yc + 1) * t->by; y++)
      for (int x = xc * t->bx; x < (xc + 1) * t->bx; x++)
        for (int d = 0; d < t->dof_f; d++) {
          const size_t f = ((size_t)y * t->Xf + x) * t->dof_f + d;
          acc += std::conj(t->null[v][f]) * fine[f];
        }
    coarse[i] = acc;
  }
  return GLB_OK;
}
// ---- multigrid set-up: host loops in the reference's order
int glb_rscale(glb_context*, int dt, size_t n, const void* x, double sc, void* out) {
  BOTH(dt, { for (size_t i = 0; i < n; i++) { T v = ((const T*)x)[i]; v *= sc; ((T*)out)[i] = v; } });
  return GLB_OK;
}
int glb_op_set_shifts(glb_operator* o, const double sh[2], const double eo[2], const double df[2]) {
  if (sh) o->op->shift = cplx(sh[0], sh[1]);
  if (eo) o->op->eo_shift = cplx(eo[0], eo[1]);
  if (df) o->op->dof_shift = cplx(df[0], df[1]);
  return GLB_OK;
}
int glb_op_get_shifts(const glb_operator* o, double sh[2], double eo[2], double df[2]) {
  if (sh) { sh[0] = o->op->shift.real(); sh[1] = o->op->shift.imag(); }
  if (eo) { eo[0] = o->op->eo_shift.real(); eo[1] = o->op->eo_shift.imag(); }
  if (df) { df[0] = o->op->dof_shift.real(); df[1] = o->op->dof_shift.imag(); }
  return GLB_OK;
}
int glb_op_stencil_download(glb_operator* o, void* cl, void* hp) {
  if (cl) std::memcpy(cl, o->op->clover.data(), o->op->clover.size() * sizeof(cplx));
  if (hp) std::memcpy(hp, o->op->hopping.data(), o->op->hopping.size() * sizeof(cplx));
  return GLB_OK;
}
int glb_mg_transfer_create_dev(glb_context* c, int Xf, int Yf, int dof_f, int bx, int by, int nvec,
                               const void* const* d_null_vectors, glb_mg_transfer** out) {
  return glb_mg_transfer_create(c, Xf, Yf, dof_f, bx, by, nvec, d_null_vectors, out);
}
// mg_complex.cpp:259-370 block_orthonormalize followed by block_normalize (:191-256)
int glb_mg_block_orthonormalize(glb_context*, int Xf, int Yf, int dof_f, int bx, int by, int nvec, void* const* nulls) {
  g_calls++;
  cplx** nv = (cplx**)nulls;
  const int Xc = Xf / bx, Yc = Yf / by;
  for (int b = 0; b < Xc * Yc; b++) {
    const int x0 = (b % Xc) * bx, y0 = (b / Xc) * by;
    std::vector<size_t> idx;
    for (int y = y0; y < y0 + by; y++)
      for (int x = x0; x < x0 + bx; x++)
        for (int d = 0; d < dof_f; d++) idx.push_back(((size_t)y * Xf + x) * dof_f + d);
    for (int c = 1; c < nvec; c++) {
      double norm = 0.0;
      for (size_t k = 0; k < idx.size(); k++) norm += real(conj(nv[c - 1][idx[k]]) * nv[c - 1][idx[k]]);
      norm = sqrt(norm);
      for (size_t k = 0; k < idx.size(); k++) nv[c - 1][idx[k]] /= norm;
      for (int m = 0; m < c; m++) {
        cplx dot_prod = 0.0;
        for (size_t k = 0; k < idx.size(); k++) dot_prod += conj(nv[m][idx[k]]) * nv[c][idx[k]];
        for (size_t k = 0; k < idx.size(); k++) nv[c][idx[k]] -= dot_prod * nv[m][idx[k]];
      }
    }
    for (int c = 0; c < nvec; c++) {
      double norm = 0.0;
      for (size_t k = 0; k < idx.size(); k++) norm += real(conj(nv[c][idx[k]]) * nv[c][idx[k]]);
      norm = sqrt(norm);
      for (size_t k = 0; k < idx.size(); k++) nv[c][idx[k]] /= norm;
    }
  }
  return GLB_OK;
}
// null_gen.cpp:26-35 (by_colour = 0) / :109-126 (by_colour = 1)
int glb_mg_partition(glb_context*, int X, int Y, int dof, int colour_period, void* even_io, void* odd_out) {
  g_calls++;
  cplx* e = (cplx*)even_io; cplx* o = (cplx*)odd_out;
  const size_t n = (size_t)X * Y * dof;
  for (size_t i = 0; i < n; i++) {
    bool odd;
    if (colour_period > 0) {
      odd = (int)(i % colour_period) >= colour_period / 2;
    } else {
      const size_t site = i / dof;
      odd = ((site % X + site / X) % 2) != 0;
    }
    if (odd) { o[i] = e[i]; e[i] = 0.0; }
  }
  return GLB_OK;
}
// null_gen.cpp:74-88 (colour_period = 0) / :132-152 (colour_period = n_vectors[curr_level])
int glb_mg_partition_corner(glb_context*, int X, int Y, int dof, int colour_period, int which, void* src_io, void* dst_out) {
  g_calls++;
  cplx* e = (cplx*)src_io; cplx* o = (cplx*)dst_out;
  const size_t n = (size_t)X * Y * dof;
  for (size_t i = 0; i < n; i++) {
    int cls = 0;
    if (colour_period > 0) {
      const int c = (int)(i % colour_period), p = colour_period;
      if (c >= p / 4 && c < 2 * p / 4) cls = 1;
      else if (c >= 2 * p / 4 && c < 3 * p / 4) cls = 2;
      else if (c >= 3 * p / 4) cls = 3;
    } else {
      const size_t site = i / dof;
      const int x = (int)(site % X), y = (int)(site / X);
      if (x % 2 == 1 && y % 2 == 0) cls = 2;
      else if (x % 2 == 0 && y % 2 == 1) cls = 3;
      else if (x % 2 == 1 && y % 2 == 1) cls = 1;
    }
    if (cls == which) { o[i] = e[i]; e[i] = 0.0; }
  }
  return GLB_OK;
}
// P^dag A P summed directly (the device kernel's formulation of mg_complex.cpp:827-1026)
int glb_mg_galerkin(glb_mg_transfer* t, glb_operator* fine, int ignore_shifts, glb_operator** coarse) {
  g_calls++;
  const PortOp& f = *fine->op;
  const int nv = t->nvec, df = t->dof_f, Xf = t->Xf, Yf = t->Yf, Xc = t->Xc, Yc = t->Yc;
  if (f.has_two || (Xc & 1) || (Yc & 1) || Xc < 2 || Yc < 2) return GLB_ERR_ARG;
  const size_t Lf = (size_t)Xf * Yf * df, per = (size_t)Xc * Yc * nv * nv;
  std::vector<cplx> cl(per, 0.0), hp(4 * per, 0.0);
  const bool us = !ignore_shifts && std::abs(f.shift) != 0.0, ue = !ignore_shifts && std::abs(f.eo_shift) != 0.0,
             ud = !ignore_shifts && std::abs(f.dof_shift) != 0.0;
  for (size_t cs = 0; cs < (size_t)Xc * Yc; cs++) {
    const int xc = (int)(cs % Xc), yc = (int)(cs / Xc);
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) {
        cplx acc_c = 0.0, acc_h[4] = {0.0, 0.0, 0.0, 0.0};
        for (int y = yc * t->by; y < (yc + 1) * t->by; y++)
          for (int x = xc * t->bx; x < (xc + 1) * t->bx; x++) {
            const size_t site = (size_t)y * Xf + x;
            const int xn[4] = {(x + 1) % Xf, x, (x + Xf - 1) % Xf, x};
            const int yn[4] = {y, (y + 1) % Yf, y, (y + Yf - 1) % Yf};
            const bool inside[4] = {(x + 1) % t->bx != 0, (y + 1) % t->by != 0, x % t->bx != 0, y % t->by != 0};
            for (int r = 0; r < df; r++) {
              const size_t fi = site * df + r;
              const cplx ci = t->null[i][fi];
              cplx row = 0.0;
              for (int c = 0; c < df; c++) row += f.clover[c + df * fi] * t->null[j][site * df + c];
              const cplx self = t->null[j][fi];
              if (us) row += f.shift * self;
              if (ue) row += (((x + y) & 1) ? -f.eo_shift : f.eo_shift) * self;
              if (ud) row += (r < df / 2 ? f.dof_shift : -f.dof_shift) * self;
              acc_c += conj(ci) * row;
              for (int d = 0; d < 4; d++) {
                const size_t g = ((size_t)yn[d] * Xf + xn[d]) * df;
                cplx h = 0.0;
                for (int c = 0; c < df; c++) h += f.hopping[c + df * fi + d * df * Lf] * t->null[j][g + c];
                if (inside[d]) acc_c += conj(ci) * h; else acc_h[d] += conj(ci) * h;
              }
            }
          }
        const size_t o = (cs * nv + i) * nv + j;
        cl[o] = acc_c;
        for (int d = 0; d < 4; d++) hp[o + d * per] = acc_h[d];
      }
  }
  const double z[2] = {0.0, 0.0};
  return glb_op_create_stencil2d(fine->ctx, cl.data(), hp.data(), 0, Xc, Yc, nv, z, z, z, coarse);
}
}  // extern "C"
