// generic_cg_precond.h -- kept so that `#include "generic_cg_precond.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters_precond.h.
#ifndef GLB200_FWD_generic_cg_precond_H
#define GLB200_FWD_generic_cg_precond_H
#include "generic_inverters_precond.h"
#endif
