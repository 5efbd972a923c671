// Stand-in for <gsl/gsl_spline.h>: natural cubic spline (gsl_interp_cspline semantics),
// used by rcBKfunc / Largex only (their table files are absent upstream).
#ifndef SMC_SHIM_GSL_SPLINE_H
#define SMC_SHIM_GSL_SPLINE_H
#include <vector>
#include <cstddef>
struct gsl_interp_accel { size_t cache; };
typedef int gsl_interp_type;
static const gsl_interp_type* gsl_interp_cspline = 0;
struct gsl_spline { std::vector<double> x, y, y2; size_t n; };
static inline gsl_interp_accel* gsl_interp_accel_alloc() { gsl_interp_accel* a = new gsl_interp_accel; a->cache = 0; return a; }
static inline void gsl_interp_accel_free(gsl_interp_accel* a) { delete a; }
static inline gsl_spline* gsl_spline_alloc(const gsl_interp_type*, size_t n) { gsl_spline* s = new gsl_spline; s->n = n; return s; }
static inline void gsl_spline_free(gsl_spline* s) { delete s; }
static inline int gsl_spline_init(gsl_spline* s, const double* xa, const double* ya, size_t n) {
  s->n = n; s->x.assign(xa, xa + n); s->y.assign(ya, ya + n); s->y2.assign(n, 0.0);
  std::vector<double> u(n, 0.0);
  for (size_t i = 1; i + 1 < n; i++) {
    double sig = (s->x[i] - s->x[i-1]) / (s->x[i+1] - s->x[i-1]);
    double p = sig * s->y2[i-1] + 2.0;
    s->y2[i] = (sig - 1.0) / p;
    u[i] = (s->y[i+1] - s->y[i]) / (s->x[i+1] - s->x[i]) - (s->y[i] - s->y[i-1]) / (s->x[i] - s->x[i-1]);
    u[i] = (6.0 * u[i] / (s->x[i+1] - s->x[i-1]) - sig * u[i-1]) / p;
  }
  for (size_t k = n - 1; k-- > 0;) s->y2[k] = s->y2[k] * s->y2[k+1] + u[k];
  return 0;
}
static inline double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel*) {
  size_t lo = 0, hi = s->n - 1;
  while (hi - lo > 1) { size_t k = (hi + lo) >> 1; if (s->x[k] > x) hi = k; else lo = k; }
  double h = s->x[hi] - s->x[lo], a = (s->x[hi] - x) / h, b = (x - s->x[lo]) / h;
  return a * s->y[lo] + b * s->y[hi] + ((a*a*a - a) * s->y2[lo] + (b*b*b - b) * s->y2[hi]) * (h*h) / 6.0;
}
#endif
