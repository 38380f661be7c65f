// generic_bicgstab_precond.h -- kept so that `#include "generic_bicgstab_precond.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters_precond.h.
#ifndef GLB200_FWD_generic_bicgstab_precond_H
#define GLB200_FWD_generic_bicgstab_precond_H
#include "generic_inverters_precond.h"
#endif
