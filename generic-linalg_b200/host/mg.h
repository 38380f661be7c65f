// mg.h -- solver selectors of the multigrid preconditioner; drop-in for multigrid/aa_mg/mg.h:4-20.
#ifndef GLB200_MG_H
#define GLB200_MG_H

enum inner_solver {
  NONE = 0,
  MINRES = 1,  // not on the accelerated path
  CG = 2,
  GCR = 3,
  BICGSTAB = 4,
  CR = 5,
  BICGSTAB_L = 6,
};

enum outer_solver {
  OUTER_GCR = 0,       // VPGCR
  OUTER_CG = 1,        // FPCG (not on the accelerated path)
  OUTER_BICGSTAB = 2   // PBiCGStab (not on the accelerated path)
};

#endif
