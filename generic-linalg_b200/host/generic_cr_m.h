// generic_cr_m.h -- kept so that `#include "generic_cr_m.h"` in code written against the reference still
// compiles; every prototype lives in generic_inverters_precond.h.
#ifndef GLB200_FWD_generic_cr_m_H
#define GLB200_FWD_generic_cr_m_H
#include "generic_inverters_precond.h"
#endif
