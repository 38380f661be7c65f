/* c_abi_cgne.c -- the C ABI (include/glb200.h) from plain C: build a 2-D U(1) staggered operator from a gauge field in the
 * reference layout, run one D apply and a device-resident CG solve of D^dag D x = D^dag b, read the report.  This is the
 * whole surface a cgo / JNI / ctypes binding needs for the headline path.
 *
 *   gcc -std=c99 -I include examples/c_abi_cgne.c -L generic-linalg_b200 -lglb200 -lm \
 *       -Wl,-rpath,$PWD/generic-linalg_b200 -o c_abi_cgne && ./c_abi_cgne [L]
 *
 * Without a CUDA device glb_create fails and the program says so (there is no CPU path). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "glb200.h"

#define CHECK(call)                                                     \
  do {                                                                  \
    int rc_ = (call);                                                   \
    if (rc_ != GLB_OK) {                                                \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, glb_last_error()); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 256;
  const size_t V = (size_t)L * L;
  glb_context* ctx = NULL;
  CHECK(glb_create(0, &ctx));

  /* synthetic inputs: links exp(i theta) in the layout lattice[y*L*2 + x*2 + mu], a right-hand side (re, im) */
  double* links = (double*)malloc(sizeof(double) * 4 * V);
  double* b = (double*)malloc(sizeof(double) * 2 * V);
  double* x = (double*)malloc(sizeof(double) * 2 * V);
  unsigned s = 1337u;
  for (size_t i = 0; i < 2 * V; i++) {
    s = s * 1664525u + 1013904223u;
    const double theta = 0.4 * ((double)(s >> 8) / 16777216.0 - 0.5);
    links[2 * i] = cos(theta);
    links[2 * i + 1] = sin(theta);
  }
  for (size_t i = 0; i < 2 * V; i++) {
    s = s * 1664525u + 1013904223u;
    b[i] = (double)(s >> 8) / 16777216.0 - 0.5;
  }

  glb_operator *D = NULL, *Ddag = NULL, *DdagD = NULL;
  CHECK(glb_op_create_staggered(ctx, links, L, L, 0.1, 0u, &D));
  CHECK(glb_op_create_staggered(ctx, links, L, L, 0.1, GLB_STAG_DAGGER, &Ddag));
  CHECK(glb_op_create_staggered(ctx, links, L, L, 0.1, GLB_STAG_NORMAL, &DdagD));

  void *d_b = NULL, *d_bp = NULL, *d_x = NULL, *d_chk = NULL;
  CHECK(glb_vec_alloc(ctx, GLB_COMPLEX, V, &d_b));
  CHECK(glb_vec_alloc(ctx, GLB_COMPLEX, V, &d_bp));
  CHECK(glb_vec_alloc(ctx, GLB_COMPLEX, V, &d_x));
  CHECK(glb_vec_alloc(ctx, GLB_COMPLEX, V, &d_chk));
  CHECK(glb_vec_upload(ctx, GLB_COMPLEX, V, d_b, b));
  CHECK(glb_op_apply(Ddag, d_bp, d_b));            /* b' = D^dag b */
  CHECK(glb_vec_zero(ctx, GLB_COMPLEX, V, d_x));   /* zero initial guess */

  glb_cg_report rep;
  CHECK(glb_cg_solve(DdagD, d_x, d_bp, 100000, 1e-10, &rep, NULL, 0));

  /* true residual of the original system, |D x - b| / |b| */
  double nb = 0.0, nr = 0.0;
  CHECK(glb_op_apply(D, d_chk, d_x));
  CHECK(glb_diffnorm2sq(ctx, GLB_COMPLEX, V, d_chk, d_b, &nr));
  CHECK(glb_norm2sq(ctx, GLB_COMPLEX, V, d_b, &nb));
  CHECK(glb_vec_download(ctx, GLB_COMPLEX, V, x, d_x));
  printf("CGNE on %d x %d: %d iterations, %d operator applies, |Dx-b|/|b| = %.3e, x[0] = (%.6e, %.6e), %llu kernel launches\n",
         L, L, rep.iterations, rep.ops, sqrt(nr / nb), x[0], x[1], glb_kernel_launches());

  glb_vec_free(ctx, d_b);
  glb_vec_free(ctx, d_bp);
  glb_vec_free(ctx, d_x);
  glb_vec_free(ctx, d_chk);
  glb_op_destroy(D);
  glb_op_destroy(Ddag);
  glb_op_destroy(DdagD);
  glb_destroy(ctx);
  free(links);
  free(b);
  free(x);
  return 0;
}
