// TEST INFRASTRUCTURE.  The 19 operator identities the reference checks in tests/staggered_pieces/staggered_pieces.cpp
// (:256-741; that program no longer compiles against the reference's own headers), written afresh
// against the reference's public interface: functions of operators.h, stencils of operators_stencil.h /
// coarse_stencil.h (full and e/o, o/e, t/b, b/t partial applies), the even/odd- and top/bottom-preconditioned solves,
// and the "hypercube into internal degrees of freedom" rotation the reference builds out of its multigrid interface
// (null_generate_free with BLOCK_CORNER, block_orthonormalize, generate_coarse_from_fine_stencil, restrict, prolong).
// tests/test_reference_programs_cpu.py builds this file against the reference's sources and against
// generic-linalg_b200/host and compares the two outputs: every identity has to hold, and what each side computed
// (the checksum of the left-hand side, solver iteration counts) has to agree.
//
//   pieces_driver [L] [mass]      prints one line per identity:  T<n> <squared difference> <checksum re> <checksum im> [iters]
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

using namespace std;

#include "coarse_stencil.h"
#include "generic_bicgstab_l.h"
#include "generic_cg.h"
#include "generic_gcr.h"
#include "generic_gmres.h"
#include "generic_vector.h"
#include "lattice.h"
#ifdef PIECES_DECLARE_LATTICE_FUNCTIONS  // reference build: lattice_functions.h DEFINES its two functions non-inline and
void lattice_epsilon(complex<double>* out, complex<double>* in, Lattice* latt);  // mg_complex.cpp includes it too -- a
void lattice_sigma3(complex<double>* out, complex<double>* in, Lattice* latt);   // second inclusion cannot link
#else
#include "lattice_functions.h"
#endif
#include "mg.h"
#include "mg_complex.h"
#include "null_gen.h"
#include "operators.h"
#include "operators_stencil.h"
#include "u1_utils.h"
#include "verbosity.h"

typedef complex<double> zc;
typedef vector<zc> zvec;

static int g_n;

static void report(int id, const zvec& a, const zvec& b, int it1 = -1, int it2 = -1) {
  zc s = 0.0;
  for (int i = 0; i < g_n; i++) s += a[i] * (double)(1 + i % 5);
  double d = 0.0, nrm = 0.0;
  for (int i = 0; i < g_n; i++) {
    d += norm(a[i] - b[i]);
    nrm += norm(a[i]);
  }
  printf("T%d %.17g %.17g %.17g", id, d / nrm, s.real(), s.imag());
  if (it1 >= 0) printf(" %d %d", it1, it2);
  printf("\n");
}

int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 16;
  const double mass = argc > 2 ? atof(argv[2]) : 0.1;
  const int n = L * L;
  const double tol = 1e-10;
  g_n = n;
  mt19937 gen(1337);

  zvec U(2 * n), b(n), x(n), y(n), t(n), t2(n), bp(n), xi(n), yi(n), bi(n);
  gauss_gauge_u1(U.data(), L, L, gen, 6.0);
  gaussian<double>(b.data(), n, gen);

  staggered_u1_op D;
  D.lattice = U.data();
  D.mass = mass;
  D.x_fine = L;
  D.y_fine = L;
  D.Nc = 1;
  void* Dv = (void*)&D;

  inversion_verbose_struct verb;
  verb.verbosity = VERB_NONE;
  verb.verb_prefix = "";
  verb.precond_verbosity = VERB_NONE;
  verb.precond_verb_prefix = "";

  // ---- functions of operators.h (staggered_pieces.cpp:256-346)
  square_staggered_gamma5_u1(x.data(), b.data(), Dv);                       // 1: g5 D  ==  g5 (D .)
  square_staggered_u1(t.data(), b.data(), Dv);
  gamma_5(y.data(), t.data(), Dv);
  report(1, x, y);

  square_staggered_dagger_u1(x.data(), b.data(), Dv);                       // 2: D^dag  ==  g5 D g5
  gamma_5(t.data(), b.data(), Dv);
  square_staggered_u1(t2.data(), t.data(), Dv);
  gamma_5(y.data(), t2.data(), Dv);
  report(2, x, y);

  square_staggered_u1(x.data(), b.data(), Dv);                              // 3: D  ==  D_eo + D_oe + m
  square_staggered_deo_u1(y.data(), b.data(), Dv);
  square_staggered_doe_u1(t.data(), b.data(), Dv);
  for (int i = 0; i < n; i++) y[i] = y[i] + t[i] + mass * b[i];
  report(3, x, y);

  square_staggered_normal_u1(x.data(), b.data(), Dv);                       // 4: D^dag D  ==  m^2 - D_eo D_oe - D_oe D_eo
  square_staggered_doe_u1(t.data(), b.data(), Dv);
  square_staggered_deo_u1(y.data(), t.data(), Dv);
  square_staggered_deo_u1(t.data(), b.data(), Dv);
  square_staggered_doe_u1(t2.data(), t.data(), Dv);
  for (int i = 0; i < n; i++) y[i] = mass * mass * b[i] - y[i] - t2[i];
  report(4, x, y);

  {                                                                         // 5: even/odd preconditioned solve == direct solve
    zero<double>(x.data(), n);
    inversion_info a = minv_vector_bicgstab_l(x.data(), b.data(), n, 100000, tol, 4, square_staggered_u1, Dv, &verb);
    square_staggered_eoprec_prepare(bp.data(), b.data(), Dv);
    zero<double>(t.data(), n);
    inversion_info c = minv_vector_cg(t.data(), bp.data(), n, 100000, tol, square_staggered_m2mdeodoe_u1, Dv, &verb);
    square_staggered_eoprec_reconstruct(y.data(), t.data(), b.data(), Dv);
    report(5, x, y, a.iter, c.iter);
  }

  // ---- the same through stencils (:351-508)
  int dims[2] = {L, L};
  Lattice lat0(2, dims, 1), lat0g(2, dims, 1), lat0d(2, dims, 1);
  stencil_2d S(&lat0, get_stencil_size(STAGGERED)), Sg(&lat0g, get_stencil_size(STAGGERED)), Sd(&lat0d, get_stencil_size(STAGGERED));
  get_square_staggered_u1_stencil(&S, &D);
  get_square_staggered_gamma5_u1_stencil(&Sg, &D);
  get_square_staggered_dagger_u1_stencil(&Sd, &D);

  square_staggered_u1(x.data(), b.data(), Dv);                              // 6
  apply_stencil_2d(y.data(), b.data(), &S);
  report(6, x, y);
  square_staggered_gamma5_u1(x.data(), b.data(), Dv);                       // 7
  apply_stencil_2d(y.data(), b.data(), &Sg);
  report(7, x, y);
  square_staggered_dagger_u1(x.data(), b.data(), Dv);                       // 8
  apply_stencil_2d(y.data(), b.data(), &Sd);
  report(8, x, y);

  apply_stencil_2d(x.data(), b.data(), &Sg);                                // 9: g5 D stencil == epsilon (D stencil)
  apply_stencil_2d(t.data(), b.data(), &S);
  lattice_epsilon(y.data(), t.data(), S.lat);
  report(9, x, y);

  apply_stencil_2d(x.data(), b.data(), &Sd);                                // 10: D^dag stencil == epsilon D epsilon
  lattice_epsilon(t.data(), b.data(), S.lat);
  apply_stencil_2d(t2.data(), t.data(), &S);
  lattice_epsilon(y.data(), t2.data(), S.lat);
  report(10, x, y);

  apply_stencil_2d(x.data(), b.data(), &S);                                 // 11: partial applies add up
  apply_stencil_2d_eo(y.data(), b.data(), &S);
  apply_stencil_2d_oe(t.data(), b.data(), &S);
  for (int i = 0; i < n; i++) y[i] = y[i] + t[i] + S.shift * b[i];
  report(11, x, y);

  apply_stencil_2d(t.data(), b.data(), &S);                                 // 12: D^dag D from the partial applies
  apply_stencil_2d(x.data(), t.data(), &Sd);
  apply_stencil_2d_eo(t.data(), b.data(), &S);
  apply_stencil_2d_oe(y.data(), t.data(), &S);
  apply_stencil_2d_oe(t.data(), b.data(), &S);
  apply_stencil_2d_eo(t2.data(), t.data(), &S);
  for (int i = 0; i < n; i++) y[i] = S.shift * S.shift * b[i] - y[i] - t2[i];
  report(12, x, y);

  {                                                                         // 13: even/odd preconditioned stencil solve
    zero<double>(x.data(), n);
    inversion_info a = minv_vector_bicgstab_l(x.data(), b.data(), n, 100000, tol, 4, apply_stencil_2d, &S, &verb);
    apply_square_staggered_eoprec_prepare_stencil(bp.data(), b.data(), &S);
    zero<double>(t.data(), n);
    inversion_info c = minv_vector_cg(t.data(), bp.data(), n, 100000, tol, apply_square_staggered_m2mdeodoe_stencil, &S, &verb);
    apply_square_staggered_eoprec_reconstruct_stencil(y.data(), t.data(), b.data(), &S);
    report(13, x, y, a.iter, c.iter);
  }

  // ---- 2x2 hypercubes rotated into 4 internal degrees of freedom (:510-741): the multigrid interface used as a
  // unitary change of basis -- one corner-partitioned constant vector, block-normalised
  int dims1[2] = {L / 2, L / 2};
  Lattice lat1(2, dims1, 4);
  mg_operator_struct_complex mg;
  mg.x_fine = L;
  mg.y_fine = L;
  mg.Nc = 1;
  mg.n_refine = 1;
  int two = 2, four = 4;
  mg.blocksize_x = &two;
  mg.blocksize_y = &two;
  mg.n_vectors = &four;
  Lattice* lats[2] = {&lat0, &lat1};
  mg.latt = lats;
  stencil_2d S1(&lat1, get_stencil_size(STAGGERED));
  stencil_2d* stens[2] = {&S, &S1};
  mg.stencils = stens;
  mg.have_dagger_stencil = false;
  mg.dagger_stencils = 0;
  vector<zvec> corner(4, zvec(n, 0.0));
  zc* cptr[4] = {corner[0].data(), corner[1].data(), corner[2].data(), corner[3].data()};
  zc** levels[1] = {cptr};
  mg.null_vectors = levels;
  mg.matrix_vector = square_staggered_u1;
  mg.matrix_vector_dagger = 0;
  mg.matrix_extra_data = Dv;
  mg.dslash_count = new dslash_tracker(1);
  mg.curr_level = 0;
  mg.curr_dof_fine = 1;
  mg.curr_x_fine = L;
  mg.curr_y_fine = L;
  mg.curr_fine_size = n;
  mg.curr_dof_coarse = 4;
  mg.curr_x_coarse = L / 2;
  mg.curr_y_coarse = L / 2;
  mg.curr_coarse_size = n;

  null_vector_params nv;
  nv.opt_null = STAGGERED;
  nv.n_null_vectors.push_back(1);
  nv.null_partitions = 4;
  nv.bstrat = BLOCK_CORNER;
  nv.do_global_ortho_conj = false;
  nv.do_ortho_eo = false;
  null_generate_free(&mg, &nv, false, 0);
  block_orthonormalize(&mg);
  generate_coarse_from_fine_stencil(&S1, &S, &mg, true);
  S1.shift = S.shift;

  apply_stencil_2d(x.data(), b.data(), &S);                                 // 14: D == P D_internal R
  restrict(bi.data(), b.data(), &mg);
  apply_stencil_2d(xi.data(), bi.data(), &S1);
  prolong(y.data(), xi.data(), &mg);
  report(14, x, y);

  apply_stencil_2d(x.data(), b.data(), &Sg);                                // 15: g5 D == P sigma3 D_internal R
  apply_stencil_2d(xi.data(), bi.data(), &S1);
  lattice_sigma3(t.data(), xi.data(), &lat1);
  prolong(y.data(), t.data(), &mg);
  report(15, x, y);

  apply_stencil_2d(x.data(), b.data(), &Sd);                                // 16: D^dag == P sigma3 D_internal sigma3 R
  lattice_sigma3(xi.data(), bi.data(), &lat1);
  apply_stencil_2d(t.data(), xi.data(), &S1);
  lattice_sigma3(xi.data(), t.data(), &lat1);
  prolong(y.data(), xi.data(), &mg);
  report(16, x, y);

  apply_stencil_2d(x.data(), b.data(), &S);                                 // 17: top/bottom pieces add up
  apply_stencil_2d_tb(xi.data(), bi.data(), &S1);
  apply_stencil_2d_bt(t.data(), bi.data(), &S1);
  for (int i = 0; i < n; i++) xi[i] = xi[i] + t[i] + S1.shift * bi[i];
  prolong(y.data(), xi.data(), &mg);
  report(17, x, y);

  apply_stencil_2d(t.data(), b.data(), &S);                                 // 18: D^dag D from the top/bottom pieces
  apply_stencil_2d(x.data(), t.data(), &Sd);
  apply_stencil_2d_bt(t.data(), bi.data(), &S1);
  apply_stencil_2d_tb(xi.data(), t.data(), &S1);
  apply_stencil_2d_tb(t.data(), bi.data(), &S1);
  apply_stencil_2d_bt(t2.data(), t.data(), &S1);
  for (int i = 0; i < n; i++) xi[i] = S1.shift * S1.shift * bi[i] - xi[i] - t2[i];
  prolong(y.data(), xi.data(), &mg);
  report(18, x, y);

  {                                                                         // 19: top/bottom preconditioned solve
    zero<double>(xi.data(), n);
    inversion_info a = minv_vector_bicgstab_l(xi.data(), bi.data(), n, 100000, tol, 4, apply_stencil_2d, &S1, &verb);
    apply_square_staggered_tbprec_prepare_stencil(bp.data(), bi.data(), &S1);
    zero<double>(t.data(), n);
    inversion_info c = minv_vector_cg(t.data(), bp.data(), n, 100000, tol, apply_square_staggered_m2mdtbdbt_stencil, &S1, &verb);
    apply_square_staggered_tbprec_reconstruct_stencil(yi.data(), t.data(), bi.data(), &S1);
    prolong(x.data(), xi.data(), &mg);
    prolong(y.data(), yi.data(), &mg);
    report(19, x, y, a.iter, c.iter);
  }

  // ---- tests/staggered_gcr_cgne_equiv/gcr_cgne_equiv.cpp:243-289 (stale in the reference tree as well): a source
  // projected on the even sites, solved by GCR(inf), GMRES(inf), CG on D^dag D with D^dag b and with b itself.  The
  // second column of these lines is |x_solver - x_GCR|^2 / |x_GCR|^2 (not an identity: 23 solves another system).
  {
    gamma_5(t.data(), b.data(), Dv);
    for (int i = 0; i < n; i++) bp[i] = 0.5 * (b[i] + t[i]);
    zero<double>(x.data(), n);
    inversion_info a = minv_vector_gcr(x.data(), bp.data(), n, 100000, tol, square_staggered_u1, Dv, &verb);
    report(20, x, x, a.iter, a.ops_count);
    zero<double>(y.data(), n);
    a = minv_vector_gmres(y.data(), bp.data(), n, 100000, tol, square_staggered_u1, Dv, &verb);
    report(21, y, x, a.iter, a.ops_count);
    gamma_5(t.data(), bp.data(), Dv);
    square_staggered_u1(t2.data(), t.data(), Dv);
    gamma_5(t.data(), t2.data(), Dv);
    zero<double>(y.data(), n);
    a = minv_vector_cg(y.data(), t.data(), n, 100000, tol, square_staggered_normal_u1, Dv, &verb);
    report(22, y, x, a.iter, a.ops_count);
    zero<double>(y.data(), n);
    a = minv_vector_cg(y.data(), bp.data(), n, 100000, tol, square_staggered_normal_u1, Dv, &verb);
    report(23, y, x, a.iter, a.ops_count);
  }

  // ---- one direction of a stencil at a time (stencil_2d::sdir; multigrid/aa_mg/tests.cpp:520-622 "piece" test): the
  // pieces add up to the full apply -- for the full and the four partial applies, on the 4-colour one-link stencil
  // (24.x) and on a two-link stencil of D^dag D generated by probing (25.x: 13 directions)
  {
    typedef void (*apply_fn)(zc*, zc*, void*);
    const apply_fn variants[5] = {apply_stencil_2d, apply_stencil_2d_eo, apply_stencil_2d_oe, apply_stencil_2d_tb, apply_stencil_2d_bt};
    stencil_2d S2(&lat0, get_stencil_size(STAGGERED_NORMAL));
    generate_stencil_2d(&S2, square_staggered_normal_u1, Dv);
    for (int which = 0; which < 2; which++) {
      stencil_2d* st = which == 0 ? &S1 : &S2;
      const zvec& src = which == 0 ? bi : b;
      const int last = which == 0 ? (int)DIR_YM1 : (int)DIR_XP1YM1;
      for (int v = 0; v < (which == 0 ? 5 : 3); v++) {
        st->sdir = DIR_ALL;
        variants[v](x.data(), const_cast<zc*>(src.data()), st);
        zero<double>(y.data(), n);
        zc piece_sum = 0.0;
        for (int d = (int)DIR_0; d <= last; d++) {
          st->sdir = (stencil_dir)d;
          variants[v](t.data(), const_cast<zc*>(src.data()), st);
          for (int i = 0; i < n; i++) {
            y[i] += t[i];
            piece_sum += t[i] * (double)(d + i % 3);
          }
        }
        st->sdir = DIR_ALL;
        printf("T%d.%d %.17g %.17g %.17g %.17g %.17g\n", 24 + which, v, diffnorm2sq<double>(x.data(), y.data(), n) / norm2sq<double>(b.data(), n),
               piece_sum.real(), piece_sum.imag(), x[1].real(), x[n - 2].imag());
      }
    }
  }

  // ---- block_normalize on its own (mg_complex.cpp:191; the driver's PDAGP_TEST, aa_mg_square_staggered_u1.cpp:1388-1470):
  // four gaussian vectors scaled to unit norm on every 2x2 block, then restrict -> prolong through them
  {
    for (int v = 0; v < 4; v++) gaussian<double>(corner[v].data(), n, gen);
    block_normalize(&mg);
    double worst = 0.0;
    for (int v = 0; v < 4; v++)
      for (int by = 0; by < L; by += 2)
        for (int bx = 0; bx < L; bx += 2) {
          const double nb = norm(corner[v][by * L + bx]) + norm(corner[v][by * L + bx + 1]) + norm(corner[v][(by + 1) * L + bx]) +
                            norm(corner[v][(by + 1) * L + bx + 1]);
          if (fabs(nb - 1.0) > worst) worst = fabs(nb - 1.0);
        }
    zc cs = 0.0;
    for (int v = 0; v < 4; v++)
      for (int i = 0; i < n; i++) cs += corner[v][i] * (double)(1 + v + i % 7);
    restrict(xi.data(), b.data(), &mg);
    prolong(y.data(), xi.data(), &mg);
    printf("T26 %.3g %.17g %.17g %.17g %.17g\n", worst, cs.real(), cs.imag(), y[0].real(), y[n - 1].imag());
  }
  return 0;
}
