// verbosity.h -- reporting controls and printers.  Drop-in for the reference's verbosity.h:9-133:
// same enum values, struct members and printed line formats (tests/bicgstab_l/run_test.sh parses
// them by field position with awk).
#ifndef GLB200_VERBOSITY_H
#define GLB200_VERBOSITY_H

#include <iostream>
#include <string>

enum inversion_verbose_level {
  VERB_PASS_THROUGH = -1,  // preconditioner inherits the outer level
  VERB_NONE = 0,
  VERB_SUMMARY = 1,         // one line at the end
  VERB_RESTART_DETAIL = 2,  // + one line per restart
  VERB_DETAIL = 3           // + one line per iteration
};

struct inversion_verbose_struct {
  inversion_verbose_level verbosity;
  std::string verb_prefix;
  inversion_verbose_level precond_verbosity;
  std::string precond_verb_prefix;
};

// verbosity of the inner solver of a restarted run (verbosity.h:27-44)
inline void shuffle_verbosity_restart(inversion_verbose_struct* inner, inversion_verbose_struct* outer) {
  if (outer == 0) {
    inner->verbosity = VERB_NONE;
    inner->verb_prefix = "";
    inner->precond_verbosity = VERB_NONE;
    inner->precond_verb_prefix = "";
    return;
  }
  const bool quiet = (outer->verbosity == VERB_RESTART_DETAIL || outer->verbosity == VERB_SUMMARY);
  inner->verbosity = quiet ? VERB_NONE : outer->verbosity;
  inner->verb_prefix = outer->verb_prefix;
  inner->precond_verbosity = outer->precond_verbosity;
  inner->precond_verb_prefix = outer->precond_verb_prefix;
}

// verbosity handed to a preconditioner (verbosity.h:46-70)
inline void shuffle_verbosity_precond(inversion_verbose_struct* inner, inversion_verbose_struct* outer) {
  inner->precond_verbosity = VERB_NONE;
  inner->precond_verb_prefix = "";
  if (outer == 0) {
    inner->verbosity = VERB_NONE;
    inner->verb_prefix = "";
    return;
  }
  inner->verbosity = (outer->precond_verbosity == VERB_PASS_THROUGH) ? outer->verbosity : outer->precond_verbosity;
  inner->verb_prefix = outer->precond_verb_prefix;
}

inline void print_verbosity_resid(inversion_verbose_struct* verb, std::string alg, int iter, int ops_count,
                                  double relres) {
  if (verb != 0 && verb->verbosity == VERB_DETAIL)
    std::cout << verb->verb_prefix << alg << " Iter " << iter << " Ops " << ops_count << " RelRes " << relres << "\n";
}

inline void print_verbosity_summary(inversion_verbose_struct* verb, std::string alg, bool success, int iter,
                                    int ops_count, double relres) {
  if (verb != 0 && verb->verbosity >= VERB_SUMMARY)
    std::cout << verb->verb_prefix << alg << " Success " << (success ? "Y" : "N") << " Iter " << iter << " Ops "
              << ops_count << " RelRes " << relres << "\n";
}

// the algorithm label is hard-wired to "CG-M " in the reference (verbosity.h:107); kept
inline void print_verbosity_summary_multi(inversion_verbose_struct* verb, std::string alg, bool success, int iter,
                                          int ops_count, double* relres, int n_res) {
  (void)alg;
  if (verb != 0 && verb->verbosity >= VERB_SUMMARY) {
    std::cout << verb->verb_prefix << "CG-M " << " Success " << (success ? "Y" : "N") << " Iter " << iter << " Ops "
              << ops_count << " RelRes ";
    for (int n = 0; n < n_res; n++) std::cout << relres[n] << " ";
    std::cout << "\n";
  }
}

inline void print_verbosity_restart(inversion_verbose_struct* verb, std::string alg, int iter, int ops_count,
                                    double relres) {
  if (verb != 0 && verb->verbosity >= VERB_RESTART_DETAIL)
    std::cout << verb->verb_prefix << alg << " Iter " << iter << " Ops " << ops_count << " RelRes " << relres << "\n";
}

#endif
