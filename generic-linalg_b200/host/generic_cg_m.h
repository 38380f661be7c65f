// generic_cg_m.h -- kept so that `#include "generic_cg_m.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters.h.
#ifndef GLB200_FWD_generic_cg_m_H
#define GLB200_FWD_generic_cg_m_H
#include "generic_inverters.h"
#endif
