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
// the device operators of the hierarchy.  normal_eqn_smooth / normal_eqn_mg must be false.
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

#endif
