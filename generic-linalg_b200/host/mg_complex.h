// mg_complex.h -- the adaptive-multigrid preconditioner on DEVICE-resident vectors: counterpart of
// multigrid/aa_mg/mg_complex.h.
//
// What is kept: the reference's cycle (mg_preconditioner, mg_complex.cpp:514-822) with its parameters --
// pre-smooth, residual, restrict, coarse solve or recursion (MLEVEL_SMOOTH = V cycle, MLEVEL_RECURSIVE =
// K cycle through VPGCR), prolong, post-smooth -- its printed progress lines and its dslash counters.
// What changes: every level's operator is a device operator (include/glb200.h: the fine staggered
// stencil or a coarse stencil_2d uploaded once with glb_op_create_stencil2d), the grid transfers are the
// device kernels of glb_mg_prolong / glb_mg_restrict, and all work vectors live in HBM.  The hierarchy is
// either handed over as arrays (null vectors + per-level stencils built elsewhere) or set up on the device:
// null_generate_random_smooth_dev (null_gen.h), block_orthonormalize_dev and
// generate_coarse_from_fine_stencil_dev below.
#ifndef GLB200_MG_COMPLEX_H
#define GLB200_MG_COMPLEX_H

#include <complex>

#include "generic_inverters.h"
#include "glb200.h"
#include "mg.h"
#include "verbosity.h"

// mg_complex.h:17-21
enum mg_multilevel_type { MLEVEL_SMOOTH = 0, MLEVEL_RECURSIVE = 1 };

// mg_complex.h:104-136
struct dslash_tracker {
  int n_refine;
  int* krylov;
  int* presmooth;
  int* postsmooth;
  int* residual;
  int* nullvectors;
  explicit dslash_tracker(int refine);
  ~dslash_tracker();
};

// Device counterpart of mg_operator_struct_complex (mg_complex.h:139-182): one operator per level
// (stencils[0] fine ... stencils[n_refine] coarsest), one transfer per refinement, the current level.
struct mg_operator_struct_complex_dev {
  int n_refine;
  glb_operator** stencils;
  glb_mg_transfer** transfers;
  int curr_level;
  dslash_tracker* dslash_count;
  // ---- set-up state (needed by the *_dev set-up routines only; names of mg_complex.h:139-182)
  int x_fine, y_fine;                      // top-level lattice; Nc = 1
  int* blocksize_x;                        // [n_refine]
  int* blocksize_y;                        // [n_refine]
  int* n_vectors;                          // [n_refine] null vectors per refinement = coarse dofs per site
  std::complex<double>*** null_vectors;    // DEVICE arrays null_vectors[level][v], level-l lattice size each
  // BLOCK_TOPO only (null_gen.cpp:36-71 builds the taste-singlet projectors from the symmetric shifts of the gauge
  // field in matrix_extra_data): the two shift operators on the top level, e.g. from
  // glb200_operator_from_callback(staggered_symmshift_x / _y, &stagif)
  glb_operator* symmshift_x;
  glb_operator* symmshift_y;
  // ---- normal-equation variants of the cycle only (mg_precond_struct_complex_dev::normal_eqn_smooth / _mg): D^dag of
  // every level they touch -- level 0: e.g. glb200_operator_from_callback(square_staggered_dagger_u1, &stagif) (the
  // reference applies matrix_vector_dagger there, mg_complex.cpp:119-123), below it the dagger stencils (:97-100)
  glb_operator** dagger_stencils;
};

// lattice of level l (mg_complex.h: latt[l]): sites and dofs per site
void mg_level_dims(const mg_operator_struct_complex_dev* mgstruct, int level, int* X, int* Y, int* dof);

// block_orthonormalize + block_normalize (mg_complex.cpp:191-370) of null_vectors[curr_level], in place on the device.
void block_orthonormalize_dev(mg_operator_struct_complex_dev* mgstruct);

// generate_coarse_from_fine_stencil (mg_complex.cpp:827-1026) at curr_level: builds transfers[curr_level] from
// null_vectors[curr_level] (replacing an existing one) and stencils[curr_level+1] = P^dag stencils[curr_level] P
// (replacing an existing one) as a device stencil2d operator.  ignore_shifts as in the reference: true leaves the
// fine shifts out of the coarse clover (the caller copies the shift down, aa_mg_square_staggered_u1.cpp:1080-1093).
void generate_coarse_from_fine_stencil_dev(mg_operator_struct_complex_dev* mgstruct, bool ignore_shifts);

// mg_precond_struct_complex (mg_complex.h:185-236) without the function pointers: the operators are
// the device operators of the hierarchy.  normal_eqn_smooth (the smoother runs on D^dag D z = D^dag r, CGNR) and
// normal_eqn_mg (the whole cycle runs on D^dag D: fine, coarse and smoothing operator) need mgstruct->dagger_stencils.
struct mg_precond_struct_complex_dev {
  minv_inverter in_smooth_type;
  double omega_smooth;
  int* n_pre_smooth;
  int* n_post_smooth;
  bool normal_eqn_mg;
  bool normal_eqn_smooth;
  mg_multilevel_type mlevel_type;
  inner_solver in_solve_type;
  int n_max;
  int n_restart;
  double* rel_res;
  mg_operator_struct_complex_dev* mgstruct;
  bool quiet;  // true: do not print the reference's "[MG]: ..." / "[L2]: ..." progress lines
};

// mg_complex.h:53-57
void level_down(mg_operator_struct_complex_dev* mgstruct);
void level_up(mg_operator_struct_complex_dev* mgstruct);

// d_lhs = M^-1 d_rhs: one multigrid cycle from the current level down (mg_complex.cpp:514).  Has the
// device variant of the reference's preconditioner signature, so it plugs into
// minv_vector_gcr_var_precond(_restart)_dev; extra_data is a mg_precond_struct_complex_dev*.
void mg_preconditioner_dev(std::complex<double>* d_lhs, std::complex<double>* d_rhs, int size, void* extra_data,
                           inversion_verbose_struct* verb = 0);

// =====================================================================================================================
// The reference's HOST-pointer multigrid interface (multigrid/aa_mg/mg_complex.h:22-236), kept so that a driver written
// against it compiles and runs: the structs with the reference's members, and the routines on plain host arrays.  Each
// routine ships its operands to the device, runs the device kernel / cycle above, and ships the result back, so values
// agree with the _dev forms; inside a drop-in solve (minv_vector_gcr_var_precond(_restart) with fine_square_staggered
// and mg_preconditioner as its callbacks) the hierarchy is uploaded ONCE and the whole solve stays on the device.
// Every level needs a generated stencil (the reference's explicit-projection fall-back for stencil-less levels is not
// carried over).
// =====================================================================================================================
#include "coarse_stencil.h"

// mg_complex.h:139-182
struct mg_operator_struct_complex {
  int x_fine;
  int y_fine;
  int n_refine;       // 1 = two levels, 2 = three levels, ...
  int* blocksize_x;   // per refinement
  int* blocksize_y;
  unsigned int Nc;    // colours on the top level (square_laplace only)
  Lattice** latt;     // one per level (n_refine + 1)
  stencil_2d** stencils;
  bool have_dagger_stencil;
  stencil_2d** dagger_stencils;
  int* n_vectors;                         // null vectors per refinement = dofs per site of the next level
  std::complex<double>*** null_vectors;   // [refinement][vector][fine dof of that level], host arrays
  void (*matrix_vector)(std::complex<double>*, std::complex<double>*, void*);         // top-level operator
  void (*matrix_vector_dagger)(std::complex<double>*, std::complex<double>*, void*);
  void* matrix_extra_data;
  int curr_level;
  int curr_dof_fine, curr_x_fine, curr_y_fine, curr_fine_size;
  int curr_dof_coarse, curr_x_coarse, curr_y_coarse, curr_coarse_size;
  dslash_tracker* dslash_count;
};

// mg_complex.h:185-236
struct mg_precond_struct_complex {
  minv_inverter in_smooth_type;
  double omega_smooth;
  int* n_pre_smooth;
  int* n_post_smooth;
  bool normal_eqn_mg;
  bool normal_eqn_smooth;
  mg_multilevel_type mlevel_type;
  inner_solver in_solve_type;
  int n_max;
  int n_restart;
  double* rel_res;
  mg_operator_struct_complex* mgstruct;
  void (*fine_matrix_vector)(std::complex<double>*, std::complex<double>*, void*);
  void (*coarse_matrix_vector)(std::complex<double>*, std::complex<double>*, void*);
  void (*fine_matrix_vector_dagger)(std::complex<double>*, std::complex<double>*, void*);
  void (*coarse_matrix_vector_dagger)(std::complex<double>*, std::complex<double>*, void*);
  void (*fine_matrix_vector_normal)(std::complex<double>*, std::complex<double>*, void*);
  void (*coarse_matrix_vector_normal)(std::complex<double>*, std::complex<double>*, void*);
  void* matrix_extra_data;
};

// mg_complex.h:24-44: the operator of the current level / of the level below it (extra_data: mg_operator_struct_complex*);
// the daggered forms use dagger_stencils when present and epsilon / sigma_3 conjugation of the stencil otherwise
void coarse_square_staggered(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void fine_square_staggered(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void coarse_square_staggered_dagger(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void fine_square_staggered_dagger(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void coarse_square_staggered_normal(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void fine_square_staggered_normal(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
// mg_complex.h:50-68
void block_normalize(mg_operator_struct_complex* mgstruct);       // mg_complex.cpp:191
void block_orthonormalize(mg_operator_struct_complex* mgstruct);  // mg_complex.cpp:259, ends with block_normalize
void prolong(std::complex<double>* x_fine, std::complex<double>* x_coarse, mg_operator_struct_complex* mgstruct);
void restrict(std::complex<double>* x_coarse, std::complex<double>* x_fine, mg_operator_struct_complex* mgstruct);
void level_down(mg_operator_struct_complex* mgstruct);
void level_up(mg_operator_struct_complex* mgstruct);
void generate_coarse_from_fine_stencil(stencil_2d* stenc_coarse, stencil_2d* stenc_fine, mg_operator_struct_complex* mgstruct,
                                       bool ignore_shifts);
void generate_coarse_from_fine_stencil(stencil_2d* stenc_coarse, stencil_2d* stenc_fine, mg_operator_struct_complex* mgstruct);
// mg_complex.h:238: one cycle, lhs = M^-1 rhs from the current level down; extra_data: mg_precond_struct_complex*
void mg_preconditioner(std::complex<double>* lhs, std::complex<double>* rhs, int size, void* extra_data,
                       inversion_verbose_struct* verb = 0);

// the device image of a host hierarchy (used by mg_preconditioner above and by the drop-in solvers, which keep one for
// the duration of a solve)
#include <vector>
namespace glb200_mg_host {
struct Hierarchy {
  glb_context* ctx;
  std::vector<glb_operator*> ops;
  std::vector<glb_operator*> dag;  // D^dag per level when the host struct has them (normal-equation variants), else empty
  std::vector<glb_mg_transfer*> trs;
  mg_operator_struct_complex_dev mg;
  mg_precond_struct_complex_dev pc;
  // with_dagger: also D^dag of every level (the normal-equation variants of the cycle need them)
  explicit Hierarchy(mg_operator_struct_complex* host, bool with_dagger = false);
  void set_precond(const mg_precond_struct_complex* p);
  void release();
  ~Hierarchy();

 private:
  Hierarchy(const Hierarchy&);
  Hierarchy& operator=(const Hierarchy&);
};
glb_operator* upload_stencil(glb_context* ctx, stencil_2d* st);
glb_operator* upload_adjoint_stencil(glb_context* ctx, stencil_2d* st);  // D^dag of a generated stencil
}  // namespace glb200_mg_host

#endif
