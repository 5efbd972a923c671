// smc_host_math.h -- run constants computed once on the host at context creation.
// Product code (not the oracle): what the reference computes in the MCnucl, GaussianNucleonsCal and
// Nucleus constructors (src/MCnucl.cpp:57-64,94-96,120; src/GaussianNucleonsCal.cpp:24-55,130-163;
// src/Nucleus.cpp:65-148; src/Regge96.cpp:27-50).
#pragma once
#include <cmath>
#include <cstring>

namespace smc_host {

// sigma_inel(pp) = sigma_tot - sigma_el from the PDG-1996 Regge fit (src/Regge96.cpp:27-50, channel 0)
inline double sigma_inel(double ecm) {
  const double s = ecm * ecm;
  const double tot = 22.0 * std::pow(s, 0.079) + 56.1 * std::pow(s, -0.46);
  const double bel = 2.0 * 2.3 + 2.0 * 2.3 + 4.0 * std::pow(s, 0.0808) - 4.2;
  return tot - 0.0511 * tot * tot / bel;
}

// Gauss-Legendre 38-point rule, positive half tabulated to 13 digits as in src/Nucleus.cpp:697-751
inline void gauss38_unit(double* xn, double* wn) {
  static const double xp[19] = {4.078514790458e-2, 1.220840253379e-1, 2.025704538921e-1, 2.817088097902e-1,
    3.589724404794e-1, 4.338471694324e-1, 5.058347179279e-1, 5.744560210478e-1, 6.392544158297e-1, 6.997986803792e-1,
    7.556859037540e-1, 8.065441676053e-1, 8.520350219324e-1, 8.918557390046e-1, 9.257413320486e-1, 9.534663309335e-1,
    9.748463285902e-1, 9.897394542664e-1, 9.980499305357e-1};
  static const double wp[19] = {8.152502928039e-2, 8.098249377060e-2, 7.990103324353e-2, 7.828784465821e-2,
    7.615366354845e-2, 7.351269258474e-2, 7.038250706690e-2, 6.678393797914e-2, 6.274093339213e-2, 5.828039914700e-2,
    5.343201991033e-2, 4.822806186076e-2, 4.270315850467e-2, 3.689408159400e-2, 3.083950054518e-2, 2.457973973823e-2,
    1.815657770961e-2, 1.161344471647e-2, 5.002880749632e-3};
  for (int i = 0; i < 38; i++) {
    const double x = i < 19 ? -xp[18 - i] : xp[i - 19], w = i < 19 ? wp[18 - i] : wp[i - 19];
    xn[i] = (1.0 - 0.0) * x / 2.0 + (0.0 + 1.0) / 2.0;      // mapped to [0,1]
    wn[i] = (1.0 - 0.0) * w / 2.0;
  }
}

// sigma_gg solving sigma_in = Int d^2b [1 - exp(-sigma_gg Tpp(b))] (src/GaussianNucleonsCal.cpp:130-163)
inline double sigma_gg_newton(double siginNN, double width) {
  double xs[38], ws[38];
  gauss38_unit(xs, ws);
  const double target = siginNN * 0.1, bmax = 5.0 * width;
  double sg = 10.0, prev;
  do {
    prev = sg;
    double f = 0.0, df = 0.0;
    for (int k = 0; k < 38; k++) {
      const double b = xs[k] * bmax, db = ws[k] * bmax;
      const double tpp = std::exp(-b * b / (4. * width * width)) / (M_PI * (4. * width * width));
      f += 2 * M_PI * b * db * (1.0 - std::exp(-sg * tpp));
      df += 2 * M_PI * b * db * tpp * std::exp(-sg * tpp);
    }
    sg -= (f - target) / df;
  } while (std::fabs(sg - prev) > 1e-4);
  return sg;
}

// Integral of exp(-t)/t over [lo, hi] by nested Simpson refinement: level L uses 2^L panels whose midpoints are the only new
// nodes, so every integrand value is computed once.  The width of shape_of_nucleons 3 inherits the rounding of the
// reference's integrator (src/arsenal.cpp:531-571, called with tolerance 1e-10): the node formula lo + h (k + 1/2), the order
// of the three partial sums and the stopping rule (refined - previous <= tol, at most 51 refinements) are the same, which is
// what makes the constants bit-equal (tests/test_oracle_golden.py, pbpb5020_lambda_width).
inline double simpson_exp_over_t(double lo, double hi, double tol) {
  const auto integrand = [](double t) { return 1. / t * std::exp(-t); };
  const double span = hi - lo, ends = integrand(lo) + integrand(hi);
  double interior = 0.0;      // integrand summed over the interior panel boundaries of the current level
  double h = span;            // panel width of the current level
  double estimate = 0.0;
  for (int level = 0;; level++) {
    const long panels = 1L << level;
    double mid = 0.0;
    for (long k = 0; k < panels; k++) mid += integrand(lo + h * (k + 0.5));
    const double refined = (span / 6 / panels) * (ends + interior * 2. + mid * 4.);
    if (level > 0 && (!(std::fabs(refined - estimate) > tol) || level > 50)) return refined;
    estimate = refined; interior += mid; h /= 2.0;
  }
}

// GaussianNucleonsCal constructor (src/GaussianNucleonsCal.cpp:24-55); false for an unknown shape
inline bool gaussian_nucleon(int shape, double siginNN, double user_width, double gaussian_lambda, double* width, double* sigma_gg) {
  if (shape == 3) {                 // energy-dependent width from sigma_in / sigma_gg = (gamma_E + E1(lambda) + ln lambda) / lambda
    const double lam = gaussian_lambda;
    const double ratio = (0.5772156649 + simpson_exp_over_t(lam, lam + 100., 1e-10) + std::log(lam)) / lam;
    *width = std::sqrt(siginNN * 0.1 / (4 * M_PI * lam * ratio));
    *sigma_gg = siginNN * 0.1 / ratio;
    return true;
  }
  if (shape == 1) *width = std::sqrt(0.1 * siginNN / (M_PI)) / 2.0;
  else if (shape == 2) *width = std::sqrt(0.1 * siginNN / M_PI) / std::sqrt(8);
  else if (shape == 4) *width = user_width;
  else return false;
  *sigma_gg = sigma_gg_newton(siginNN, *width);
  return true;
}

struct WoodsSaxon { double rad, dr, rmaxCut, rwMax, beta2, beta4; };
inline WoodsSaxon woods_saxon(int A, int deformed) {
  WoodsSaxon w; std::memset(&w, 0, sizeof w);
  if (A <= 1) return w;
  const double a = (double)A;
  w.rad = 1.12 * std::pow(a, 0.333333) - 0.86 / std::pow(a, 0.333333); w.dr = 0.54;
  switch (A) {
    case 197: w.rad = 6.42; w.dr = 0.45; break;
    case 63: w.rad = 4.28; w.dr = 0.5; break;
    case 238: w.rad = 6.86; w.dr = 0.44; break;
    case 208: w.rad = 6.67; w.dr = 0.44; break;
    case 129: w.rad = 5.36; w.dr = 0.590; break;
  }
  w.rmaxCut = w.rad + 2.5;
  w.rwMax = 1.0 / (1.0 + std::exp(-w.rad / w.dr));
  if (deformed) {
    switch (A) {
      case 197: w.beta2 = -0.13; w.beta4 = -0.03; break;
      case 63: w.beta2 = 0.162; w.beta4 = 0.006; break;
      case 129: w.beta2 = 0.162; w.beta4 = -0.003; break;
      case 238: w.beta2 = 0.28; w.beta4 = 0.093; break;
    }
  }
  return w;
}

// largest double t with fl(sqrt(t)) <= c : turns the reference's `sqrt(dc) > 5w -> 0` into `dc <= t`
inline double sqrt_threshold(double c) {
  double t = c * c;
  while (std::sqrt(t) > c) t = std::nextafter(t, 0.0);
  while (std::sqrt(std::nextafter(t, INFINITY)) <= c) t = std::nextafter(t, INFINITY);
  return t;
}

}  // namespace smc_host
