// generic_minres.h -- kept so that `#include "generic_minres.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_minres_H
#define GLB200_FWD_generic_minres_H
#include "generic_inverters.h"
#endif
