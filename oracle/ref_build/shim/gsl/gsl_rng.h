// Stand-in for <gsl/gsl_rng.h>: GSL is not installed in this image.
// Only the entry points the reference calls (MCnucl.cpp:74-81,198) are provided.
// TEST INFRASTRUCTURE ONLY (used to build oracle/_ref from /root/reference sources).
#ifndef SMC_SHIM_GSL_RNG_H
#define SMC_SHIM_GSL_RNG_H
#include <random>
struct gsl_rng { std::mt19937 g; };
typedef int gsl_rng_type;
static const gsl_rng_type* gsl_rng_default = 0;
static inline void gsl_rng_env_setup() {}
static inline gsl_rng* gsl_rng_alloc(const gsl_rng_type*) { return new gsl_rng; }
static inline void gsl_rng_set(gsl_rng* r, unsigned long s) { r->g.seed((unsigned)s); }
static inline void gsl_rng_free(gsl_rng* r) { delete r; }
static inline double gsl_rng_uniform(gsl_rng* r) {
  return std::uniform_real_distribution<double>(0.0, 1.0)(r->g);
}
#endif
