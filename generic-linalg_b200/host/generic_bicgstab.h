// generic_bicgstab.h -- kept so that `#include "generic_bicgstab.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_bicgstab_H
#define GLB200_FWD_generic_bicgstab_H
#include "generic_inverters.h"
#endif
