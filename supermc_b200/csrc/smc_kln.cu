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
};

__device__ __forceinline__ double kln_alpha_s(double q2) {      // KLNModel.h:90-95 (alphaS=0.5, Lambda=0.2, Nf=3)
  const double lq2 = 0.2 * 0.2, beta0 = (33.0 - 2.0 * 3.0) / (12 * SMC_PI);
  if (q2 <= lq2) return 0.5;
  return fmin(0.5, 1.0 / (beta0 * log(q2 / lq2)));
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
  kln_table_kernel<<<g, 128, 0, s>>>(k, table);
  return cudaGetLastError();
}

}  // namespace smc
