// generic_bicgstab_l.h -- kept so that `#include "generic_bicgstab_l.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_bicgstab_l_H
#define GLB200_FWD_generic_bicgstab_l_H
#include "generic_inverters.h"
#endif
