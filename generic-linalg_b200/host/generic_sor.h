// generic_sor.h -- kept so that `#include "generic_sor.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_sor_H
#define GLB200_FWD_generic_sor_H
#include "generic_inverters.h"
#endif
