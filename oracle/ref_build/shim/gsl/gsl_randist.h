// Stand-in for <gsl/gsl_randist.h>; gsl_ran_gamma(r, shape a, scale b) as used at
// MCnucl.cpp:1285,1298. Distribution-equivalent, not stream-equivalent, to GSL.
#ifndef SMC_SHIM_GSL_RANDIST_H
#define SMC_SHIM_GSL_RANDIST_H
#include "gsl_rng.h"
static inline double gsl_ran_gamma(gsl_rng* r, double a, double b) {
  return std::gamma_distribution<double>(a, b)(r->g);
}
#endif
