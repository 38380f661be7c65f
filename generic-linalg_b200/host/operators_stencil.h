// operators_stencil.h -- fill a stencil_2d from a gauge field (host, once per field):
// drop-in for operator_utils/operators_stencil.h (builders at operators_stencil.cpp:14,65,120).
#ifndef GLB200_OPERATORS_STENCIL_H
#define GLB200_OPERATORS_STENCIL_H
#include "coarse_stencil.h"
#include "operators.h"

void get_square_staggered_u1_stencil(stencil_2d* stenc, staggered_u1_op* stagif);
void get_square_staggered_gamma5_u1_stencil(stencil_2d* stenc, staggered_u1_op* stagif);
void get_square_staggered_dagger_u1_stencil(stencil_2d* stenc, staggered_u1_op* stagif);

#endif
