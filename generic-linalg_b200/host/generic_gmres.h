// generic_gmres.h -- kept so that `#include "generic_gmres.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_gmres_H
#define GLB200_FWD_generic_gmres_H
#include "generic_inverters.h"
#endif
