// generic_cg.h -- kept so that `#include "generic_cg.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_cg_H
#define GLB200_FWD_generic_cg_H
#include "generic_inverters.h"
#endif
