// operators_stencil.h -- fill a stencil_2d from a gauge field (host, once per field):
// drop-in for operator_utils/operators_stencil.h (builders at operators_stencil.cpp:14,65,120).
#ifndef GLB200_OPERATORS_STENCIL_H
#define GLB200_OPERATORS_STENCIL_H
#include "coarse_stencil.h"
#include "operators.h"

void get_square_staggered_u1_stencil(stencil_2d* stenc, staggered_u1_op* stagif);
void get_square_staggered_gamma5_u1_stencil(stencil_2d* stenc, staggered_u1_op* stagif);
void get_square_staggered_dagger_u1_stencil(stencil_2d* stenc, staggered_u1_op* stagif);

// ---- even/odd preconditioned solves through a stencil (operators_stencil.h:31-38, operators_stencil.cpp:179-236):
// prepare the even right-hand side, apply m^2 - D_eo D_oe, reconstruct the odd sites.  Host vectors; the m2mdeodoe
// callback is recognised by every minv_* drop-in (the solve then runs on the device).
void apply_square_staggered_eoprec_prepare_stencil(std::complex<double>* rhs_e, std::complex<double>* rhs_orig,
                                                   stencil_2d* stenc);
void apply_square_staggered_m2mdeodoe_stencil(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void apply_square_staggered_eoprec_reconstruct_stencil(std::complex<double>* lhs_full, std::complex<double>* lhs_e,
                                                       std::complex<double>* rhs_o, stencil_2d* stenc);

// ---- the same with the colour index split into a top and a bottom half (coarse levels), the non-Galerkin normal
// operators and the daggered operators (multigrid/aa_mg/mg_complex.h:76-101, mg_complex.cpp:1211-1372)
void apply_square_staggered_tbprec_prepare_stencil(std::complex<double>* rhs_t, std::complex<double>* rhs_orig,
                                                   stencil_2d* stenc);
void apply_square_staggered_m2mdtbdbt_stencil(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void apply_square_staggered_tbprec_reconstruct_stencil(std::complex<double>* lhs_full, std::complex<double>* lhs_t,
                                                       std::complex<double>* rhs_b, stencil_2d* stenc);
void apply_square_staggered_normal_eo_stencil(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void apply_square_staggered_normal_tb_stencil(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void apply_square_staggered_dagger_eo_stencil(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);
void apply_square_staggered_dagger_tb_stencil(std::complex<double>* lhs, std::complex<double>* rhs, void* extra_data);

#endif
