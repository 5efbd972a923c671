// smc_profile3d.cu -- the 3-D extension of the reference (scripts/generate_3d_profiles/profile_3d.cpp, main.cpp): the
// participants (and binary collisions) of one event become Gaussians in (eta_s, x, y) on an neta x nx x ny lattice.
//
// profile_3d::set_variables (profile_3d.cpp:179-272): per source a space-time rapidity -- fixed +-2 (random_flag 0) or
// drawn by rejection from (1 -+ eta/y_beam) * exp(-(|eta| - 2.5)^2 / (2 * 0.5^2)) beyond |eta| > 2.5, tabulated on 1000
// points and linearly interpolated (set_eta_distribution / sample_eta_distribution_from_array, :122-177) -- and widths
// sigma_x = sigma_y = sqrt(sigma_in / 8 pi) [+- 0.3 for random_flag 2, 3], sigma_eta = 0.5 [+- 0.3].
// profile_3d::generate_3d_profile (:274-325): +-(int)(6 sigma / d) windows, no mask,
//     rho[j][k][l] += exp(-dis_eta - dis_x - dis_y) * norm_eta * norm_x * norm_y.
// The reference seeds its mt19937 with time(NULL); here the draws are Philox uniforms addressed by (seed, source, draw), so a
// run is reproducible, and the caller may also hand in the rapidities and widths (parity entry).
//
// B200 mapping: one thread per lattice cell, gathering over the sources in list order (the summation order of the
// reference), windows tested as integer ranges; 6.9 M cells x <= 416 sources.  The whole deposit is ~20 MFLOP + 17 M exps
// per event: a launch-latency-sized job, kept simple and exact rather than tiled.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/supermc_b200.h"
#include "smc_philox.h"
#include "smc_host_math.h"

namespace {

struct P3Src { double x, y, eta, sx, sy, se; int kl, kr, ll, lr, jl, jr; };

__global__ void profile3d_kernel(const P3Src* __restrict__ src, int n, int nx, int ny, int neta, double dx, double dy, double deta, double* __restrict__ rho) {
  extern __shared__ P3Src s_src[];
  const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x, ncell = (size_t)neta * nx * ny;
  const int l = (int)(cell % ny), k = (int)((cell / ny) % nx), j = (int)(cell / ((size_t)ny * nx));
  const double xg = (k - (nx - 1) / 2.) * dx, yg = (l - (ny - 1) / 2.) * dy, eg = (j - (neta - 1) / 2.) * deta;     // profile_3d.cpp:75-85
  double acc = 0.0;
  for (int i0 = 0; i0 < n; i0 += 128) {
    const int m = min(128, n - i0);
    __syncthreads();
    for (int q = threadIdx.x; q < m; q += blockDim.x) s_src[q] = src[i0 + q];
    __syncthreads();
    if (cell >= ncell) continue;
    for (int q = 0; q < m; q++) {
      const P3Src& s = s_src[q];
      if (j < s.jl || j >= s.jr || k < s.kl || k >= s.kr || l < s.ll || l >= s.lr) continue;
      const double dis_eta = (eg - s.eta) * (eg - s.eta) / (2. * s.se * s.se), norm_eta = 1. / sqrt(2. * M_PI * s.se * s.se);
      const double dis_x = (xg - s.x) * (xg - s.x) / (2. * s.sx * s.sx), norm_x = 1. / sqrt(2. * M_PI * s.sx * s.sx);
      const double dis_y = (yg - s.y) * (yg - s.y) / (2. * s.sy * s.sy), norm_y = 1. / sqrt(2. * M_PI * s.sy * s.sy);
      acc += exp(-dis_eta - dis_x - dis_y) * norm_eta * norm_x * norm_y;
    }
  }
  if (cell < ncell) rho[cell] = acc;
}

// profile_3d::binarySearch on the uniform eta table (profile_3d.cpp:431-468), skip_out_of_range = true
long bsearch_eta(const std::vector<double>& A, double value) {
  long lo = 0, hi = (long)A.size() - 1;
  if (value > A[hi] || value < A[lo]) return -1;
  long idx = (long)std::floor((hi + lo) / 2.);
  while (hi - lo > 1) { if (A[idx] < value) lo = idx; else hi = idx; idx = (long)std::floor((hi + lo) / 2.); }
  return lo;
}

}  // namespace

extern "C" int smc_profile3d(int device, const smc_profile3d_params* p, int n, const double* x, const double* y, const int* id,
                             const double* eta_in, const double* sigma3_in, double* rho_out, double* eta_used, double* sigma3_used) {
  if (!p || n < 0 || (n > 0 && (!x || !y || !id)) || !rho_out || p->nx < 1 || p->ny < 1 || p->neta < 1 || !(p->dx > 0) || !(p->dy > 0) || !(p->deta > 0)) return SMC_ERR_PARAM;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) return SMC_ERR_CUDA;       // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return SMC_ERR_CUDA;
  const double ecm = p->ecm, y_beam = std::atanh(std::sqrt(1. - 1. / std::pow(ecm / 2., 2)));              // profile_3d.cpp:25-29
  const double sigma_inelastic = smc_host::sigma_inel(ecm) * 0.1;
  // set_eta_distribution (:122-151)
  const int length = 1000; const double de = 2 * y_beam / (length - 1), eta_peak = 2.5, sig_out = 0.5;
  std::vector<double> te(length), tp(length), tt(length); double pmax = 0, tmax = 0;
  for (int i = 0; i < length; i++) {
    const double e = -y_beam + i * de; double f = 1.0;
    if (std::fabs(e) > eta_peak) f = std::exp(-std::pow(std::fabs(e) - eta_peak, 2) / (2. * sig_out * sig_out));
    te[i] = e; tp[i] = (1. - e / y_beam) * f; tt[i] = (1. + e / y_beam) * f;
    pmax = std::max(pmax, tp[i]); tmax = std::max(tmax, tt[i]);
  }
  const uint32_t k0 = (uint32_t)(uint64_t)p->seed, k1 = (uint32_t)((uint64_t)p->seed >> 32);
  auto uni = [&](int src, uint32_t draw) { smc_u4 o = smc_philox4x32_10((uint32_t)src, 0x3Du, draw >> 1, 0u, k0, k1); return (draw & 1) ? smc_u53(o.v[2], o.v[3]) : smc_u53(o.v[0], o.v[1]); };
  const double eta_0 = 2.0, s0 = std::sqrt(sigma_inelastic / (8 * M_PI)), se0 = 0.5, dsx = 0.3, dse = 0.3;   // set_variables (:179-189)
  std::vector<P3Src> src(n);
  for (int i = 0; i < n; i++) {
    P3Src s; s.x = x[i]; s.y = y[i]; uint32_t d = 0;
    auto sample_eta = [&]() {                                                                                  // :153-177
      for (;;) {
        const double e = y_beam * (1. - 2. * uni(i, d++));
        const long q = bsearch_eta(te, e);
        const double u2 = uni(i, d++);
        if (q < 0 || q + 1 >= length) continue;              // (the reference indexes out of range here; such a draw is rejected)
        const double fr = (e - te[q]) / (te[q + 1] - te[q]);
        const double pr = (id[i] == 1) ? tp[q] * (1. - fr) + tp[q + 1] * fr : tt[q] * (1. - fr) + tt[q + 1] * fr;
        if (!((id[i] == 1 ? pmax : tmax) * u2 > pr)) return e;
      }
    };
    s.sx = s.sy = s0; s.se = se0;
    if (eta_in) s.eta = eta_in[i];
    else if (p->random_flag == 0) s.eta = (id[i] == 1) ? eta_0 : -eta_0;
    else {
      s.eta = sample_eta();
      if (p->random_flag == 2) { s.se = se0 + dse * (1. - 2. * uni(i, d++)); s.sx = s.sy = s0 + dsx * (1. - 2. * uni(i, d++)); }
      else if (p->random_flag == 3) { s.sx = s0 + dsx * (1. - 2. * uni(i, d++)); s.sy = s0 + dsx * (1. - 2. * uni(i, d++)); s.se = se0 + dse * (1. - 2. * uni(i, d++)); }
    }
    if (sigma3_in) { s.sx = sigma3_in[3 * i]; s.sy = sigma3_in[3 * i + 1]; s.se = sigma3_in[3 * i + 2]; }
    // generate_3d_profile windows (:280-296)
    const int ix0 = (int)(s.x / p->dx + (p->nx - 1) / 2), iy0 = (int)(s.y / p->dy + (p->ny - 1) / 2), ie0 = (int)(s.eta / p->deta + (p->neta - 1) / 2);
    const int rx = (int)(6 * s.sx / p->dx), ry = (int)(6 * s.sy / p->dy), re = (int)(6 * s.se / p->deta);
    s.kl = std::max(ix0 - rx, 0); s.kr = std::min(ix0 + rx, p->nx); s.ll = std::max(iy0 - ry, 0); s.lr = std::min(iy0 + ry, p->ny);
    s.jl = std::max(ie0 - re, 0); s.jr = std::min(ie0 + re, p->neta);
    src[i] = s;
    if (eta_used) eta_used[i] = s.eta;
    if (sigma3_used) { sigma3_used[3 * i] = s.sx; sigma3_used[3 * i + 1] = s.sy; sigma3_used[3 * i + 2] = s.se; }
  }
  const size_t ncell = (size_t)p->neta * p->nx * p->ny;
  P3Src* d_src = nullptr; double* d_rho = nullptr;
  if (cudaMalloc(&d_src, std::max<size_t>(n, 1) * sizeof(P3Src)) != cudaSuccess || cudaMalloc(&d_rho, ncell * sizeof(double)) != cudaSuccess) { cudaFree(d_src); return SMC_ERR_NOMEM; }
  cudaMemcpy(d_src, src.data(), (size_t)n * sizeof(P3Src), cudaMemcpyHostToDevice);
  profile3d_kernel<<<(unsigned)((ncell + 255) / 256), 256, 128 * sizeof(P3Src)>>>(d_src, n, p->nx, p->ny, p->neta, p->dx, p->dy, p->deta, d_rho);
  const cudaError_t e = cudaMemcpy(rho_out, d_rho, ncell * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_src); cudaFree(d_rho);
  return e == cudaSuccess ? SMC_OK : SMC_ERR_CUDA;
}
