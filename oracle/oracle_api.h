/* oracle_api.h -- C interface shared by the two CPU checkers in oracle/.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries, and only as the checker
 * or the reported CPU baseline -- never as the thing measured or shipped.
 *
 * Two libraries implement this same interface:
 *   oracle/_ref/libref_oracle.so   (prefix ref_)  : thin shim (oracle/ref_shim.cpp)
 *        compiled against the UNMODIFIED reference sources where they lie in
 *        /root/reference (recipe: oracle/Makefile).  kind = "reference".
 *   oracle/libport_oracle.so       (prefix port_) : our own CPU restatement
 *        (oracle/port_oracle.cpp) of the same algorithms.   kind = "port".
 * The CPU tests require both to agree bit-for-bit on every fixture, and both
 * to reproduce the golden numbers in tests/golden/.
 */
#ifndef GLB200_ORACLE_API_H
#define GLB200_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ORC_PREFIX
#error "define ORC_PREFIX to ref_ or port_ before including oracle_api.h"
#endif
#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC(name) ORC_CAT(ORC_PREFIX, name)

/* Operator kinds.  Each cites the reference function it stands for. */
enum orc_op_kind {
  ORC_OP_LAPLACE_REAL = 0,      /* square_laplace.cpp:182 / unit_test.cpp:573  (double; diag 4+m2)          */
  ORC_OP_LAPLACE_IMAG = 1,      /* imag_laplace.cpp:126   (complex; diag 4+m2+i)                           */
  ORC_OP_LAPLACE_NC = 2,        /* operator_utils/operators.cpp:28  square_laplace (complex, Nc colours)   */
  ORC_OP_LAPLACE_U1 = 3,        /* operators.cpp:73   square_laplace_u1                                     */
  ORC_OP_STAG_FREE = 4,         /* operators.cpp:127  square_staggered                                      */
  ORC_OP_STAG_U1 = 5,           /* operators.cpp:184  square_staggered_u1                                   */
  ORC_OP_STAG_GAMMA5_U1 = 6,    /* operators.cpp:316  square_staggered_gamma5_u1                            */
  ORC_OP_STAG_DAGGER_U1 = 7,    /* operators.cpp:372  square_staggered_dagger_u1                            */
  ORC_OP_STAG_NORMAL_U1 = 8,    /* operators.cpp:444  square_staggered_normal_u1 (D^dag D through a tmp)    */
  ORC_OP_GAMMA5 = 9,            /* operators.cpp:242  gamma_5                                               */
  ORC_OP_STENCIL = 10,          /* stencil_2d/coarse_stencil.cpp:12 apply_stencil_2d, DIR_ALL path          */
  ORC_OP_STENCIL_FROM_STAG = 11,/* operators_stencil.cpp:14 get_square_staggered_u1_stencil + apply         */
  ORC_OP_STAG_GAMMA5_FREE = 12, /* operators.cpp:262  square_staggered_gamma5                               */
  ORC_OP_LAPLACE_REAL_NC = 13,  /* tests/multishift/multishift.cpp:634 square_laplace (double, Nc colours)  */
  ORC_OP_STAG_FREE_REAL = 14,   /* tests/multishift/multishift.cpp:677 square_staggered (double)            */
  ORC_OP_STAG_DEO_U1 = 15,      /* operators.cpp:456  square_staggered_deo_u1   (hop term, even sites)      */
  ORC_OP_STAG_DOE_U1 = 16,      /* operators.cpp:494  square_staggered_doe_u1   (hop term, odd sites)       */
  ORC_OP_STAG_M2MDEODOE_U1 = 17,/* operators.cpp:549  square_staggered_m2mdeodoe_u1 (m^2 - D_eo D_oe)       */
  ORC_OP_SYMMSHIFT_X = 18,      /* operators.cpp:688  staggered_symmshift_x                 (reference only) */
  ORC_OP_SYMMSHIFT_Y = 19,      /* operators.cpp:728  staggered_symmshift_y                 (reference only) */
  ORC_OP_STAG_2LINK_U1 = 20,    /* operators.cpp:625  square_staggered_2linklaplace_u1      (reference only) */
  ORC_OP_STAG_INDEX = 21        /* operators.cpp:782  staggered_index_operator              (reference only) */
};

/* Description of one operator.  Arrays are host pointers owned by the caller
 * and must outlive the prepared handle. Complex data is interleaved (re,im). */
typedef struct orc_op_desc {
  int kind;
  int X, Y, Nc;          /* lattice extent; Nc colours per site (Laplace-NC, stencil)                   */
  double mass;           /* mass, or m^2 for the plain Laplacians                                       */
  const double* links;   /* U(1) links, reference layout lattice[y*X*2 + x*2 + mu], complex             */
  const double* clover;  /* stencil: nc*nc*V complex                                                     */
  const double* hopping; /* stencil: 4 direction-major planes of nc*nc*V complex                         */
  const double* two_link;/* stencil: 8 planes, only if has_two                                           */
  int has_two;
  double shift[2], eo_shift[2], dof_shift[2];
  int view;              /* stencil kinds only: 0 = apply_stencil_2d, 1..6 = the composite operator built on the
                            stencil (numbering of include/glb200.h GLB_SV_*): 1 m2mdeodoe (operators_stencil.cpp:196),
                            2 m2mdtbdbt, 3 normal_eo, 4 normal_tb, 5 dagger_eo, 6 dagger_tb (mg_complex.cpp:1228-1372).
                            Reference library only.                                                       */
  double wilson_coeff;   /* staggered_u1_op::wilson_coeff (ORC_OP_STAG_2LINK_U1)                             */
} orc_op_desc;

typedef struct orc_result {
  double resSq;
  int iter;
  int success;
  int ops_count;
  int n_rhs;
  double resSqmrhs[32];
  char name[64];
} orc_result;

enum orc_solver {
  ORC_CG = 0, ORC_CG_RESTART = 1,
  ORC_CR = 2, ORC_CR_RESTART = 3,
  ORC_GCR = 4, ORC_GCR_RESTART = 5,
  ORC_BICGSTAB = 6, ORC_BICGSTAB_RESTART = 7,
  ORC_BICGSTAB_L = 8, ORC_BICGSTAB_L_RESTART = 9,
  ORC_GMRES = 10, ORC_GMRES_RESTART = 11
};

/* "reference" or "port" */
const char* ORC(kind)(void);

/* std::mt19937 stream shared by gauge and rhs generation (u1_utils.cpp:92,
 * generic_vector.h:35,48). */
void* ORC(rng_new)(unsigned seed);
void ORC(rng_free)(void* rng);
void ORC(gauss_gauge_u1)(void* rng, double* links, int X, int Y, double beta);
void ORC(unit_gauge_u1)(double* links, int X, int Y);
void ORC(gaussian_real)(void* rng, double* v, int n);
void ORC(gaussian_complex)(void* rng, double* v, int n);
/* u1_utils.cpp:17 read_gauge_u1 (text file of phases); returns 0 on success */
int ORC(read_gauge_u1)(double* links, int X, int Y, const char* path);
void ORC(plaquette_u1)(const double* links, int X, int Y, double out[2]);

/* BLAS-1 of generic_vector.h (serial left-to-right sums). is_complex selects the overload. */
void ORC(dot)(int is_complex, const double* a, const double* b, int n, double out[2]);
double ORC(norm2sq)(int is_complex, const double* a, int n);
double ORC(diffnorm2sq)(int is_complex, const double* a, const double* b, int n);

/* Operators. */
void* ORC(op_prepare)(const orc_op_desc* d);
void ORC(op_free)(void* op);
int ORC(op_is_complex)(void* op);
int ORC(op_size)(void* op);
void ORC(op_apply)(void* op, double* lhs, const double* rhs);

/* Even/odd preconditioning of the staggered operator described by `op` (any gauged staggered kind):
 * operators.cpp:528 square_staggered_eoprec_prepare, :574 square_staggered_eoprec_reconstruct. */
void ORC(eoprec_prepare)(void* op, double* rhs_e, const double* rhs_orig);
void ORC(eoprec_reconstruct)(void* op, double* lhs_full, const double* lhs_e, const double* rhs_o);

/* Partial stencil applies on an ORC_OP_STENCIL(_FROM_STAG) operator: part 1 apply_stencil_2d_eo, 2 _oe, 3 _tb,
 * 4 _bt (coarse_stencil.cpp:395, 560, 725, 1120; DIR_ALL path). */
void ORC(stencil_apply_part)(void* op, int part, double* lhs, const double* rhs);

/* Solvers: phi is in/out (initial guess -> solution), phi0 the rhs.  verbosity:
 * 0 none .. 3 detail (verbosity.h:9-16), printed to stdout exactly as the reference does. */
int ORC(solve)(int solver, void* op, double* phi, const double* phi0, int max_iter, double eps,
               int restart_freq, int l, int verbosity, orc_result* out);
/* generic_cg_m.cpp:23,312.  phi: n_shift host pointers. shifts is permuted and restored by the solver. */
int ORC(solve_cg_m)(void* op, double** phi, const double* phi0, int n_shift, int resid_freq_check,
                    int max_iter, double eps, double* shifts, int worst_first, int verbosity,
                    orc_result* out);

#ifdef __cplusplus
}
#endif
#endif
