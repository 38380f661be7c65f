// null_gen.h -- null-vector generation of the adaptive multigrid on DEVICE-resident vectors: counterpart of
// multigrid/aa_mg/null_gen.h.  Same enums, same parameter struct, same routines with a _dev suffix;
// the vectors of mgstruct->null_vectors are device arrays, the smoothing solve is the device solver family, and
// the partition / normalise / orthogonalise steps are C-ABI calls.  The gaussian sources are drawn on the host from
// the caller's std::mt19937 exactly as generic_vector.h:48-60 draws them (so the reference and this code see the
// same random numbers) and uploaded.
#ifndef GLB200_NULL_GEN_H
#define GLB200_NULL_GEN_H

#include <random>
#include <vector>

#include "generic_inverters.h"
#include "inverter_struct.h"
#include "mg_complex.h"
#include "operators.h"

// null_gen.h:14-21
enum blocking_strategy {
  BLOCK_NONE = 0,    // block fully
  BLOCK_EO = 1,      // even/odd
  BLOCK_CORNER = 2,  // corners
  BLOCK_TOPO = 3     // taste singlet: (1 +- Gamma_5)/2 from the symmetric shifts (needs mgstruct->symmshift_x/_y)
};

// null_gen.h:24-29
enum null_precond_strategy {
  NULL_PRECOND_NONE = 0,
  NULL_PRECOND_EO = 1,      // not on the accelerated path
  NULL_PRECOND_NORMAL = 2,  // not on the accelerated path
};

// How the null vectors of every refinement are made (null_gen.h:31-64).  Filled by the driver from its command line in
// the reference; here any caller of the _dev set-up routines fills it.  Per-level entries are indexed by the
// refinement (0 = top).  `opt_null` of the reference is dropped: the operator the vectors are smoothed with is
// whatever stencils[curr_level] holds when the routine is called.
struct null_vector_params {
  op_type opt_null;                     // kept for source compatibility (null_gen.h:33); the set-up smooths with stencils[curr_level]
  std::vector<int> n_null_vectors;      // per refinement: smoothed vectors BEFORE the partition multiplies them
  minv_inverter null_gen;               // solver of the smoothing solve A x = -A x0 (default BiCGStab)
  null_precond_strategy null_prec;      // plain, even/odd (top/bottom below the top level) or normal equations
  std::vector<double> null_precisions;  // per refinement: relative tolerance of the smoothing solve (driver: 5e-5)
  std::vector<int> null_max_iters;      // per refinement: its iteration cap (driver: 500)
  bool null_restart;                    // restarted solver, every null_restart_freq iterations
  int null_restart_freq;
  int null_bicgstab_l;                  // l of BiCGStab-l
  double null_relaxation;               // omega of SOR / MinRes
  double null_mass;                     // mass (stencil shift) used while generating, driver default 1e-2
  int null_partitions;                  // 1 (BLOCK_NONE), 2 (BLOCK_EO, BLOCK_TOPO) or 4 (BLOCK_CORNER)
  blocking_strategy bstrat;
  bool do_global_ortho_conj;            // also orthogonalise against the complex conjugates of earlier vectors
  bool do_ortho_eo;                     // partition every vector right after its solve instead of at the end
  bool quiet;  // true: skip the reference's "[L*_NULLVEC]: Pre-orthog cosines ..." lines (and their three reductions each)

  null_vector_params() {
    opt_null = STAGGERED;
    null_gen = MINV_BICGSTAB;
    null_prec = NULL_PRECOND_NONE;
    null_restart = false;
    null_restart_freq = -1;
    null_bicgstab_l = -1;
    null_relaxation = 1.0;
    null_mass = 0.0;
    null_partitions = 0;
    bstrat = BLOCK_NONE;
    do_global_ortho_conj = false;
    do_ortho_eo = false;
    quiet = false;
  }
};

// null_gen.cpp:13-103: partition null vector `num_null_vec` of the top level (BLOCK_EO: its odd sites move to vector
// num_null_vec + n_vectors[0]/2; BLOCK_CORNER: the three odd corners move to num_null_vec + k*n_vectors[0]/4).
// BLOCK_NONE: nothing.  BLOCK_TOPO: v and Gamma_5 v combined into the two chiral projections (null_gen.cpp:36-71).
void null_partition_staggered_dev(mg_operator_struct_complex_dev* mgstruct, int num_null_vec, blocking_strategy bstrat);

// null_gen.cpp:106-160: the same below the top level (BLOCK_EO: the upper half of the colour index moves).
void null_partition_coarse_dev(mg_operator_struct_complex_dev* mgstruct, int num_null_vec, blocking_strategy bstrat);

// null_gen.cpp:162-191: free-field null vectors -- the constant vector (times the HOST array gauge_trans on the top
// level when do_gauge_transform), partitioned and normalised part by part.
void null_generate_free_dev(mg_operator_struct_complex_dev* mgstruct, null_vector_params* nvec_params,
                            bool do_gauge_transform = false, std::complex<double>* gauge_trans = 0);

// null_gen.cpp:193-400: for each of n_vectors[curr_level]/null_partitions vectors draw a gaussian x0, orthogonalise it
// against the vectors found so far, solve A x = -A x0 from a zero guess with `null_gen` on stencils[curr_level]
// (tolerance / iteration cap of this level), keep x + x0, partition, normalise, orthogonalise, normalise.
void null_generate_random_smooth_dev(mg_operator_struct_complex_dev* mgstruct, null_vector_params* nvec_params,
                                     inversion_verbose_struct* verb, std::mt19937* generator);

// ---- the reference's HOST-pointer forms (null_gen.h:66-80) on mg_operator_struct_complex: the vectors of the current
// level go to the device, the _dev routine above runs, the vectors come back
void null_partition_staggered(mg_operator_struct_complex* mgstruct, int num_null_vec, blocking_strategy bstrat, Lattice* Lat);
void null_partition_coarse(mg_operator_struct_complex* mgstruct, int num_null_vec, blocking_strategy bstrat);
void null_generate_free(mg_operator_struct_complex* mgstruct, null_vector_params* nvec_params, bool do_gauge_transform = false,
                        std::complex<double>* gauge_trans = 0);
void null_generate_random_smooth(mg_operator_struct_complex* mgstruct, null_vector_params* nvec_params,
                                 inversion_verbose_struct* verb, std::mt19937* generator);

#endif
