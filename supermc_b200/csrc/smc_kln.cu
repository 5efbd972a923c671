// smc_kln.cu -- K5: the MC-KLN dN/dy(TA,TB) look-up table.
//
// Replaces MCnucl::makeTable (reference src/MCnucl.cpp:911-960) and, per entry, KLNModel::getdNdy ->
// ktF_MCintegral -> func (src/KLNModel.cpp:97-125,177-277,360-399; KLNfunc.h:14-17).  The reference
// integrates the 3-d (pT, kT, phi) kT-factorisation integrand with BASES/VEGAS Monte Carlo re-seeded
// to 12345 per entry (0.1 % target accuracy, 17 ms per entry on one CPU core).  Here the integrand is
// restated exactly and integrated with a deterministic Gauss-Legendre(pT) x Gauss-Legendre(kT) x
// midpoint(phi) product rule, one table entry per CTA: everything that depends on pT only (x1, x2,
// Qs^2, (1-x)^4, alpha_s(Qs^2)) is hoisted out of the (kT, phi) loops.
#include "smc_common.cuh"

namespace smc {

struct KlnCfg {
  double ecm, lambda, y, dT; int tmax; int pt_order;
  int npt, nkt, nphi;
  const double *xp, *wp, *xk, *wk, *cphi;    // device node tables
  // rcBK tabulated uGD (model 100 / 101): kt, N_A and natural-spline second derivatives [maxQ0][maxY][maxKt]
  int model, maxQ0, maxY, maxKt; double dQ0, siginNN200; const double *rkt, *rna, *ry2;
};

__device__ __forceinline__ double kln_alpha_s(double q2) {      // KLNModel.h:90-95 (alphaS=0.5, Lambda=0.2, Nf=3)
  const double lq2 = 0.2 * 0.2, beta0 = (33.0 - 2.0 * 3.0) / (12 * SMC_PI);
  if (q2 <= lq2) return 0.5;
  return fmin(0.5, 1.0 / (beta0 * log(q2 / lq2)));
}

// natural cubic spline (gsl_interp_cspline) of one (iq, iy) table at kt
__device__ __forceinline__ double rcbk_spline(const KlnCfg& k, int iq, int iy, double x) {
  const size_t o = ((size_t)iq * k.maxY + iy) * k.maxKt;
  const double* xa = k.rkt + o; const double* ya = k.rna + o; const double* y2 = k.ry2 + o;
  int lo = 0, hi = k.maxKt - 1;
  while (hi - lo > 1) { const int m = (hi + lo) >> 1; if (xa[m] > x) hi = m; else lo = m; }
  const double h = xa[hi] - xa[lo], a = (xa[hi] - x) / h, b = (x - xa[lo]) / h;
  return a * ya[lo] + b * ya[hi] + ((a * a * a - a) * y2[lo] + (b * b * b - b) * y2[hi]) * (h * h) / 6.0;
}
// rcBKfunc::getFunc (src/rcBKfunc.h:65-121): nearest bin in Y = ln(x0/x), linear in Q0^2 between tables
__device__ double rcbk_func(const KlnCfg& k, double qs0_2, double x, double kt2, double alp) {
  if (x < 0. || x > 1. || qs0_2 < 0) return 0.;
  double Y = log(0.01 / x);
  if (Y < 0.0) { qs0_2 *= exp(0.3 * Y); Y = 0.; }
  int iy = (int)(Y / 0.1 + .5);
  if (iy >= k.maxY) iy = k.maxY - 1;
  const double Q02 = 4. / 8. * qs0_2;
  int iq = (int)(Q02 / k.dQ0); iq -= 1;
  int iqoffset = 1;
  if (k.model == 100) { iq -= 1; iqoffset++; }
  if (iq == k.maxQ0 - 1) iq--; else if (iq > k.maxQ0 - 1) iq = k.maxQ0 - 2;
  const double fac = kt2 / (6. * SMC_PI * SMC_PI * SMC_PI) / alp, kk = sqrt(kt2);
  if (iq >= 0) {
    const double v1 = rcbk_spline(k, iq, iy, kk), v2 = rcbk_spline(k, iq + 1, iy, kk);
    return (v1 + (v2 - v1) * (Q02 - (iq + iqoffset) * k.dQ0) / k.dQ0) * fac;
  }
  return (rcbk_spline(k, 0, iy, kk) * Q02 / (iqoffset * k.dQ0)) * fac;
}

// same integral with the tabulated uGD (KLNModel::func / SaturationScale / waveFunction for rcBKalbacete[Set2],
// src/KLNModel.cpp:219-277,360-399)
__global__ void __launch_bounds__(128) rcbk_table_kernel(KlnCfg k, double* table) {
  const int i = blockIdx.y, j = blockIdx.x, tid = threadIdx.x;
  __shared__ double red[4];
  if (i == 0 || j == 0) { if (tid == 0) table[(size_t)i * k.tmax + j] = 0.0; return; }
  const double ta = k.dT * i, tb = k.dT * j;
  const double Ptmin = 0.1, Ptmax = 12.0, CF = (3.0 * 3.0 - 1.0) / (2 * 3.0);
  const double q0 = (k.model == 100) ? 0.399 : 0.336;
  const double qs2a = ta * k.siginNN200 / 10. * q0, qs2b = tb * k.siginNN200 / 10. * q0;
  const double ey = exp(k.y);
  double sum = 0.0;
  for (int w = tid; w < k.npt * k.nkt; w += 128) {
    const int a = w / k.nkt, b = w % k.nkt;
    const double pt = Ptmin + k.xp[a] * (Ptmax - Ptmin), mt = pt;
    const double x1 = mt / k.ecm * ey, x2 = mt / k.ecm / ey;
    if (x1 > 1.0 || x2 > 1.0) continue;
    const double om1 = 1.0 - x1, om2 = 1.0 - x2, r1 = (om1 * om1) * (om1 * om1), r2 = (om2 * om2) * (om2 * om2), m2 = mt * mt;
    double jac = (k.pt_order == 2) ? (Ptmax - Ptmin) * pt * pt : 2.0 * SMC_PI * (Ptmax - Ptmin) * pt;
    jac = jac / m2 * 2.0 * SMC_PI * pt / 4.0;
    const double kt = pt * k.xk[b], base = pt * pt + kt * kt, cross = 2 * kt * pt;
    double sphi = 0.0;
    for (int p = 0; p < k.nphi; p++) {
      const double cph = k.cphi[p];
      const double k1 = 0.25 * (base + cross * cph), k2 = 0.25 * (base - cross * cph);
      const double f1 = rcbk_func(k, qs2a, x1, k1, kln_alpha_s(k1)) * r1, f2 = rcbk_func(k, qs2b, x2, k2, kln_alpha_s(k2)) * r2;
      sphi += kln_alpha_s(fmax(fmax(k1, k2), m2)) * f1 * f2;
    }
    sum += k.wp[a] * jac * k.wk[b] * kt * sphi / k.nphi;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    const double hbarC = 0.197327053, Norm = 2. / CF / (hbarC * hbarC);
    table[(size_t)i * k.tmax + j] = 2.0 * Norm * (red[0] + red[1] + red[2] + red[3]) * 9. / 32.;
  }
}

__global__ void __launch_bounds__(128) kln_table_kernel(KlnCfg k, double* table) {
  const int i = blockIdx.y, j = blockIdx.x, tid = threadIdx.x;
  __shared__ double red[4];
  if (i == 0 || j == 0) { if (tid == 0) table[(size_t)i * k.tmax + j] = 0.0; return; }   // MCnucl.cpp:937-944
  const double ta = k.dT * i, tb = k.dT * j;
  const double Ptmin = 0.1, Ptmax = 12.0, CF = (3.0 * 3.0 - 1.0) / (2 * 3.0), fac = CF * 2. / (3. * SMC_PI * SMC_PI);
  const double ey = exp(k.y);
  double sum = 0.0;
  for (int a = tid; a < k.npt; a += 128) {
    const double pt = Ptmin + k.xp[a] * (Ptmax - Ptmin), mt = pt;
    const double x1 = mt / k.ecm * ey, x2 = mt / k.ecm / ey;
    if (x1 > 1.0 || x2 > 1.0) continue;
    const double qs2a = ta * 2. / 1.53 * pow(0.01 / x1, k.lambda), qs2b = tb * 2. / 1.53 * pow(0.01 / x2, k.lambda);   // KLNModel.cpp:371-373
    const double om1 = 1.0 - x1, om2 = 1.0 - x2;
    const double fa = fac / kln_alpha_s(qs2a) * (om1 * om1) * (om1 * om1), fb = fac / kln_alpha_s(qs2b) * (om2 * om2) * (om2 * om2);
    const double m2 = mt * mt;
    double jac = (k.pt_order == 2) ? (Ptmax - Ptmin) * pt * pt : 2.0 * SMC_PI * (Ptmax - Ptmin) * pt;
    jac = jac / m2 * 2.0 * SMC_PI * pt / 4.0;       // 1/mt^2, d^2kT = 2 pi kt ktmax dx, symmetrisation 1/4
    double spt = 0.0;
    for (int b = 0; b < k.nkt; b++) {
      const double kt = pt * k.xk[b];
      const double base = pt * pt + kt * kt, cross = 2 * kt * pt;
      double sphi = 0.0;
      for (int p = 0; p < k.nphi; p++) {
        const double cph = k.cphi[p];
        const double k1 = 0.25 * (base + cross * cph), k2 = 0.25 * (base - cross * cph);
        const double f1 = (k1 <= qs2a) ? fa : fa * qs2a / k1;            // KLNfunc.h:14-17
        const double f2 = (k2 <= qs2b) ? fb : fb * qs2b / k2;
        const double sc = fmax(fmax(k1, k2), m2);
        sphi += kln_alpha_s(sc) * f1 * f2;
      }
      spt += k.wk[b] * kt * sphi;
    }
    sum += k.wp[a] * jac * spt / k.nphi;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    const double hbarC = 0.197327053, Norm = 2. / CF / (hbarC * hbarC);
    table[(size_t)i * k.tmax + j] = 2.0 * Norm * (red[0] + red[1] + red[2] + red[3]) * 9. / 32.;   // KLNModel.cpp:110
  }
}

cudaError_t launch_kln_table(const KlnCfg& k, double* table, cudaStream_t s) {
  dim3 g(k.tmax, k.tmax);
  if (k.model >= 100) rcbk_table_kernel<<<g, 128, 0, s>>>(k, table);
  else kln_table_kernel<<<g, 128, 0, s>>>(k, table);
  return cudaGetLastError();
}

}  // namespace smc
