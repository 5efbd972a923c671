// smc_sample.cu -- K1 nucleus sampling + K2 binary-collision detection, one CTA (two warps) per event.
//
// Replaces, per event: the rejection loop of MakeDensity (reference src/MakeDensity.cpp:2147-2162),
// Nucleus::populate / getDeformRandomWS (src/Nucleus.cpp:187-317,578-621), the Particle constructor's
// AABB (src/Particle.cpp:16-99), MCnucl::getBinaryCollision / hit / createBinaryCollisions
// (src/MCnucl.cpp:217-385) and the Gamma multiplicity weights (src/MCnucl.cpp:1271-1301).
//
// B200 mapping: warp s owns nucleus s.  The Woods-Saxon rejection draws of a nucleus form one flat Philox
// stream: 32 draws are tested per step and the accepted radii, in stream order, are exactly what the
// reference's sequential do/while hands to candidates 0, 1, 2, ... (no divergent rejection loop).  The strictly
// sequential hard-core rejection ("nucleon k depends on 0..k-1") is kept *exactly* but evaluated 32 candidates
// at a time: distances to the nucleons already placed in single precision on float4 copies (a squared
// distance within 1e-4 of 0.81 falls back to the reference's double expression), conflicts inside the batch as
// bit masks resolved in candidate order -- the accepted set equals what a sequential loop over the same
// candidate stream produces.  Collisions are an all-pairs test, one projectile row per warp step, 32 target
// nucleons per instruction, with the reference's AABB sweep restated as a closed-form predicate so the set of
// pairs that consume a uniform is the reference's; pairs whose uniform is below a
// single-precision over-estimate of the hit probability queue up and are settled 32 at a time by the
// reference's double-precision expression.  Gamma weights are drawn densely over the compact participant and
// collision lists.
#include <cstdlib>
#include <algorithm>
#include "smc_common.cuh"

namespace smc {

struct SampleSmem {
  double* soa;      // [2][NROW][Amax]: structure of arrays (lane-consecutive => conflict-free), first in
                    // acceptance order (x y z, candidate id in the weight slot), then sorted in place by xL
  int Amax;
  uint32_t* hit;    // [Amax][HW] hit bit masks (row = projectile)
  int* ncB;         // [Amax]
  int* firstB;      // [Amax]
  int* rowoff;      // [Amax+1]
  int* misc;        // [16]
  double* wsq;      // [2][2][64] queue of accepted Woods-Saxon draws (r, cos theta) per warp
};

__device__ __forceinline__ void rot3(double cth, double phi, double& x, double& y, double& z) {
  // Point3D::rotate, src/MathBasics.cpp:41-50
  double sphi, cphi; sincos(phi, &sphi, &cphi);
  double sth = sqrt(1. - cth * cth), x0 = x, y0 = y, z0 = z;
  x = cth * cphi * x0 - sphi * y0 + sth * cphi * z0;
  y = cth * sphi * x0 + cphi * y0 + sth * sphi * z0;
  z = -sth * x0 + cth * z0;
}

#define S_(sm, side, f, i) (sm).soa[((size_t)(side) * NROW + (f)) * (sm).Amax + (i)]
#define SMC_MAXK 16      // ceil(512 / 32): nucleons per lane
#ifndef SMC_SORT_UNROLL
#define SMC_SORT_UNROLL 1
#endif
#ifndef SMC_BATCH_UNROLL
#define SMC_BATCH_UNROLL 4
#endif
constexpr int kSortUnroll = SMC_SORT_UNROLL, kBatchUnroll = SMC_BATCH_UNROLL;   // (macros are not expanded inside #pragma unroll)
#ifndef SMC_SORT_UNROLL
#define SMC_SORT_UNROLL 1
#endif
#ifndef SMC_BATCH_UNROLL
#define SMC_BATCH_UNROLL 4
#endif

struct Box { double xL, xR, yL, yR, xC, yC; };
__device__ __forceinline__ void box_center(Box& b, double x, double y) {   // Box2D::setCenter, src/Box2D.cpp:24-33
  b.xL = __dadd_rn(b.xL, __dadd_rn(x, -b.xC)); b.xR = __dadd_rn(b.xR, __dadd_rn(x, -b.xC));
  b.yL = __dadd_rn(b.yL, __dadd_rn(y, -b.yC)); b.yR = __dadd_rn(b.yR, __dadd_rn(y, -b.yC));
  b.xC = x; b.yC = y;
}
__device__ __forceinline__ void box_square(Box& b, double size) {          // Box2D::setDimensions, src/Box2D.cpp:35-41
  double h = size / 2;
  b.xL = __dadd_rn(b.xC, -h); b.xR = __dadd_rn(b.xC, h); b.yL = __dadd_rn(b.yC, -h); b.yR = __dadd_rn(b.yC, h);
}
__device__ __forceinline__ void box_union(Box& b, const Box& o) {          // Box2D::overUnion, src/Box2D.h:49-63
  b.xL = fmin(o.xL, b.xL); b.xR = fmax(o.xR, b.xR); b.yL = fmin(o.yL, b.yL); b.yR = fmax(o.yR, b.yR);
  b.xC = (b.xL + b.xR) / 2.0; b.yC = (b.yL + b.yR) / 2.0;
}

// Particle::Particle -> generateQuarkPositions -> calculateBounds (src/Particle.cpp:16-99)
template <bool NOQUARKS>
__device__ void particle_box(const DevCfg& c, const Store& st, const smc_stream& sq, uint32_t cand,
                             double x0, double y0, Box& out, double* extra) {
  Box base = {0, 0, 0, 0, 0, 0};
  box_center(base, x0, y0); box_square(base, 8 * c.w);
  out = base;
  if (extra) { extra[XBXL] = base.xL; extra[XBXR] = base.xR; extra[XBYL] = base.yL; extra[XBYR] = base.yR; for (int q = 0; q < 9; q++) extra[XQ + q] = 0.0;
               extra[XF] = extra[XF + 1] = extra[XF + 2] = 1.0 / 3.0; }      // Particle::resetFluctFactors, Particle.cpp:140-144
  if (NOQUARKS || c.quark_rows <= 0) {           // no table: r1 = r2 = 0, the three quark boxes sit on the nucleon
    Box b = {0, 0, 0, 0, 0, 0};
    box_center(b, 0.0, 0.0); box_square(b, 8 * c.quark_width); box_center(b, x0, y0);
    box_union(out, b); box_union(out, b); box_union(out, b);
    return;
  }
  if (NOQUARKS) return;
  double u0, u1, u2, u3;
  smc_uniform2(sq, cand, 0, &u0, &u1); smc_uniform2(sq, cand, 1, &u2, &u3);
  int index = (int)(250000 * u0);
  double r1 = 0, r2 = 0, z12 = 0;
  if (index < c.quark_rows) { r1 = st.quark_table[3 * index]; r2 = st.quark_table[3 * index + 1]; z12 = st.quark_table[3 * index + 2]; }
  r1 *= c.quark_R; r2 *= c.quark_R;
  double Theta12 = acos(z12), z1 = 2. * u1 - 1., Theta1 = acos(z1);
  double phi1 = 2 * SMC_PI * u2, phi2 = 2 * SMC_PI * u3;
  double s1, c1, s12, c12, sp1, cp1, s, cc;
  sincos(Theta1, &s1, &c1); sincos(Theta1 + Theta12, &s12, &c12); sincos(phi1, &sp1, &cp1); sincos(phi2, &s, &cc);
  double ux = s1 * cp1, uy = s1 * sp1, uz = z1, vx = s12 * cp1, vy = s12 * sp1, vz = c12;
  double r1x = r1 * ux, r1y = r1 * uy, r1z = r1 * uz;
  double r2z = (vx * (uz * ux * (1 - cc) - uy * s) + vy * (uz * uy * (1 - cc) + ux * s) + vz * (cc + uz * uz * (1 - cc))) * r2;
  double r2x = vx * (cc + ux * ux * (1 - cc)) + vy * (ux * uy * (1 - cc) - uz * s) + vz * (ux * uz * (1 - cc) + uy * s);
  double r2y = vx * (ux * uy * (1 - cc) + uz * s) + vy * (cc + uy * uy * (1 - cc)) + vz * (uy * uz * (1 - cc) - ux * s);
  r2x *= r2; r2y *= r2;
  double qx[3] = {r1x, r2x, -r1x - r2x}, qy[3] = {r1y, r2y, -r1y - r2y};
  if (extra) { const double qz[3] = {r1z, r2z, -r1z - r2z}; for (int q = 0; q < 3; q++) { extra[XQ + 3 * q] = qx[q]; extra[XQ + 3 * q + 1] = qy[q]; extra[XQ + 3 * q + 2] = qz[q]; } }
#pragma unroll
  for (int q = 0; q < 3; q++) {
    Box b = {0, 0, 0, 0, 0, 0};
    box_center(b, qx[q], qy[q]); box_square(b, 8 * c.quark_width); box_center(b, x0 + qx[q], y0 + qy[q]);
    box_union(out, b);
  }
}

__device__ __forceinline__ double sph_harm2(double ct) { return (3.0 * ct * ct - 1.0) * 0.31539156525252005; }
__device__ __forceinline__ double sph_harm4(double ct) { return (35.0 * ct * ct * ct * ct - 30.0 * ct * ct + 3.0) * 0.10578554691520431; }

// std::sort by xL (src/Nucleus.cpp:314): rank by counting (one pass over the keys updates all of this lane's
// nucleons), then an in-place permutation field by field.  MK = nucleons per lane (compile-time so that the
// rank/key/value arrays stay in registers).
template <int MK>
__device__ __forceinline__ void sort_by_xl(const DevCfg& c, const Store& st, const SampleSmem& sm, int e, int s, int A, int lane) {
  int rk[MK]; double key[MK];
#pragma unroll
  for (int m = 0; m < MK; m++) { const int k = lane + 32 * m; rk[m] = 0; key[m] = (k < A) ? S_(sm, s, NXL, k) : 0.0; }
  // rank = #{j : key_j < key_k or (key_j == key_k and j < k)}
  if (MK > 1 && MK <= 8) {
    // 64 buckets over [min, max] of the keys (the bucket index is a monotone function of the key): a key's rank is the
    // number of keys in lower buckets plus its rank among the keys of its own bucket -- about A/64 compares per key
    // instead of A.  Scratch: this warp's Woods-Saxon queue (idle now): 65 counters + A 16-bit indices.
    int* cnt = reinterpret_cast<int*>(sm.wsq + (size_t)s * 128);
    unsigned short* idx = reinterpret_cast<unsigned short*>(cnt + 68);
    double kmin = 1e300, kmax = -1e300;
#pragma unroll
    for (int m = 0; m < MK; m++) if (lane + 32 * m < A) { kmin = fmin(kmin, key[m]); kmax = fmax(kmax, key[m]); }
    for (int o = 16; o > 0; o >>= 1) { kmin = fmin(kmin, __shfl_xor_sync(0xffffffffu, kmin, o)); kmax = fmax(kmax, __shfl_xor_sync(0xffffffffu, kmax, o)); }
    const double scale = (kmax > kmin) ? 64.0 / (kmax - kmin) : 0.0;
    cnt[lane] = 0; cnt[lane + 32] = 0;
    __syncwarp();
    int bk[MK], pos[MK];
#pragma unroll
    for (int m = 0; m < MK; m++) {
      bk[m] = min(63, (int)((key[m] - kmin) * scale)); pos[m] = 0;
      if (lane + 32 * m < A) pos[m] = atomicAdd(&cnt[bk[m]], 1);
    }
    __syncwarp();
    const int c0 = cnt[2 * lane], c1 = cnt[2 * lane + 1]; int incl = c0 + c1;      // lane owns buckets 2 lane, 2 lane + 1
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    __syncwarp();
    cnt[2 * lane] = incl - c0 - c1; cnt[2 * lane + 1] = incl - c1; if (lane == 31) cnt[64] = incl;          // exclusive offsets, cnt[64] = A
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MK; m++) if (lane + 32 * m < A) idx[cnt[bk[m]] + pos[m]] = (unsigned short)(lane + 32 * m);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MK; m++) {
      const int k = lane + 32 * m;
      if (k < A) {
        const int t0 = cnt[bk[m]], t1 = cnt[bk[m] + 1];
        int r = t0;
#pragma unroll 2
        for (int t = t0; t < t1; t++) { const int j = idx[t]; const double o = S_(sm, s, NXL, j); r += (o < key[m]) || (o == key[m] && j < k); }
        rk[m] = r;
      }
    }
  } else {
  // all pairs; for keys of another 32-block the index comparison is known at compile time, which leaves one compare per pair
#pragma unroll
  for (int jb = 0; jb < MK; jb++) {
    const int jend = min(32, A - 32 * jb);
#pragma unroll kSortUnroll
    for (int jj = 0; jj < jend; jj++) {
      const double o = S_(sm, s, NXL, 32 * jb + jj);
#pragma unroll
      for (int m = 0; m < MK; m++) {
        if (m > jb) rk[m] += (o <= key[m]);
        else if (m < jb) rk[m] += (o < key[m]);
        else rk[m] += (o < key[m]) || (o == key[m] && jj < lane);
      }
    }
  }
  }
  __syncwarp();
#pragma unroll 1
  for (int f = 0; f < NROW; f++) {
    double v[MK];
#pragma unroll
    for (int m = 0; m < MK; m++) { const int k = lane + 32 * m; if (k < A) v[m] = S_(sm, s, f, k); }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MK; m++) { const int k = lane + 32 * m; if (k < A) S_(sm, s, f, rk[m]) = v[m]; }
    __syncwarp();
  }
  if (st.nuc_extra_tmp) {
#pragma unroll
    for (int m = 0; m < MK; m++) {
      const int k = lane + 32 * m;
      if (k < A) {
        const double* a = st.nuc_extra_tmp + (((size_t)e * 2 + s) * c.Amax + k) * NEXTRA; double* b2 = st.nuc_extra + (((size_t)e * 2 + s) * c.Amax + rk[m]) * NEXTRA;
#pragma unroll 1
        for (int f = 0; f < NEXTRA; f++) b2[f] = a[f];
      }
    }
  }
  __syncwarp();
}

// deuteron: inverse CDF of the Hulthen distribution by the reference's own Newton iteration with a numeric
// derivative (src/HulthenFunc.cpp:27-41, src/arsenal.cpp invertFunc: x0 = 1, dx = 1e-3, accuracy 1e-6)
__device__ double hulthen_cdf(double r) {
  const double alpha = .228, beta = 1.18;
  if (r <= 0) return 0;
  const double cc = (alpha * beta * (alpha + beta)) / ((alpha - beta) * (alpha - beta));
  return 2 * cc * (2 * (exp(-r * (alpha + beta)) / (alpha + beta)) - .5 * exp(-2 * alpha * r) / alpha - .5 * exp(-2 * beta * r) / beta + .5 / alpha + .5 / beta - 2 / (alpha + beta));
}
__device__ double hulthen_inv_cdf(double y) {
  const double xL = 0, xR = 100.0, dx = 0.001, accuracy = dx * 0.001;
  double XX2 = 1.0, XX1 = XX2 - 10 * accuracy;
  for (int it = 0; it <= 60 && fabs(XX2 - XX1) > accuracy; it++) {
    XX1 = XX2;
    const double F0 = hulthen_cdf(XX1) - y;
    const double X1 = (XX1 > xL + dx) ? XX1 - dx : xL, X2 = (XX1 < xR - dx) ? XX1 + dx : xR;
    XX2 = XX1 - F0 / ((hulthen_cdf(X1) - hulthen_cdf(X2)) / (X1 - X2));
  }
  return XX2;
}

// One nucleus by warp `s` (side).  Leaves A sorted rows in sm.pos[s].
// SPEC: 0 = every sampler; 1 = spherical Woods-Saxon on both sides; 2 = the same without a valence-quark table.  The
// specialised kernels do not carry the deformed / configuration-table / deuteron / quark-offset code at all: the hot
// instruction stream of the generic kernel (26 k SASS lines) did not fit the instruction cache (ncu: no_instruction
// 0.66 stalls per issue).  MK = nucleons per lane of the rank sort.
template <int SPEC, int MK>
__device__ void sample_nucleus(const DevCfg& c, const Store& st, const SampleSmem& sm, int e, int s, uint64_t ev,
                               uint32_t tr, double xCenter, double yCenter) {
  const int lane = threadIdx.x & 31, A = c.A[s];
  const smc_stream s_or = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_ORIENT, s);
  const smc_stream s_q = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_QUARK, s);
  double uo0, uo1; smc_uniform2(s_or, 0, 0, &uo0, &uo1);
  double ctr = 1.0 - 2.0 * uo0, phir = 2 * SMC_PI * uo1;       // Nucleus.cpp:193-197
  bool recentre = true;
  const int mode = SPEC ? 0 : c.sampler[s];
  const bool deformed = SPEC ? false : (c.deformed[s] != 0);
  if (mode == 1) {                                              // single nucleon, Nucleus.cpp:201-202
    if (lane == 0) { S_(sm, s, NX, 0) = xCenter; S_(sm, s, NY, 0) = yCenter; S_(sm, s, NZ, 0) = 0.0; S_(sm, s, NW, 0) = 0.0; }
    recentre = false;
  } else if (mode == 4) {                                       // deuteron, Nucleus.cpp:203-209,342-359
    if (lane == 0) {
      const smc_stream s_d = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_DEUT, s);
      const double d = hulthen_inv_cdf(1e-30 + (1.0 - 2e-30) * smc_uniform(s_d, 0, 0));
      double x1 = d / 2.0, y1 = 0.0, z1 = 0.0;
      rot3(ctr, phir, x1, y1, z1);
      S_(sm, s, NX, 0) = x1 + xCenter; S_(sm, s, NY, 0) = y1 + yCenter; S_(sm, s, NZ, 0) = z1; S_(sm, s, NW, 0) = 0.0;
      S_(sm, s, NX, 1) = -x1 + xCenter; S_(sm, s, NY, 1) = -y1 + yCenter; S_(sm, s, NZ, 1) = -z1; S_(sm, s, NW, 1) = 1.0;
    }
    recentre = false;
  } else if (mode == 2 || mode == 3) {                          // config tables, Nucleus.cpp:555-574,623-666
    const smc_stream s_c = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_CONFIG, s);
    double uc = smc_uniform(s_c, 0, 0);
    int icfg = (int)(uc * c.ncfg[s]); if (icfg >= c.ncfg[s]) icfg = c.ncfg[s] - 1;
    const double* cfg = st.cfg_table[s] + (size_t)icfg * A * 3;
    double mx = 0, my = 0, mz = 0;
    if (mode == 3) {
      for (int k = lane; k < A; k += 32) { mx += cfg[3 * k]; my += cfg[3 * k + 1]; mz += cfg[3 * k + 2]; }
      mx = warp_sum(mx) / A; my = warp_sum(my) / A; mz = warp_sum(mz) / A;
      double u2, u3; smc_uniform2(s_or, 0, 1, &u2, &u3);
      ctr = 1.0 - 2.0 * u2; phir = 2 * SMC_PI * u3;
    }
    for (int k = lane; k < A; k += 32) {
      double x = cfg[3 * k] - mx, y = cfg[3 * k + 1] - my, z = cfg[3 * k + 2] - mz;
      rot3(ctr, phir, x, y, z);
      if (mode == 2) { x += xCenter; y += yCenter; }
      S_(sm, s, NX, k) = x; S_(sm, s, NY, k) = y; S_(sm, s, NZ, k) = z; S_(sm, s, NW, k) = (double)k;
    }
    recentre = (mode == 3);
  } else {                                                      // Woods-Saxon + hard core, Nucleus.cpp:272-310
    const smc_stream s_ws = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_WS, s);
    const smc_stream s_an = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_ANGLE, s);
    const double rad = c.rad[s], dr = c.dr[s], rmaxCut = c.rmaxCut[s], rwMax = c.rwMax[s];
    const double rmin = 0.9 * 0.9;
    const float rmaxCut_f = (float)rmaxCut, rad_f = (float)rad, inv_dr_f = (float)(1.0 / dr);
    // The Woods-Saxon rejection draws form ONE flat stream per nucleus (draw n -> Philox (WS, n)): whether draw n is
    // accepted depends on nothing else, so 32 draws are tested at once and the accepted ones, in stream order, are
    // exactly the radii a sequential do/while loop hands to candidates 0, 1, 2, ...  (queue in shared memory)
    int placed = 0, nq = 0; uint32_t cand_base = 0, ndraw = 0;
    double* qr = sm.wsq + (size_t)s * 128; double* qc = qr + 64;
    const unsigned below = (1u << lane) - 1u;
    float4* pf = reinterpret_cast<float4*>(&S_(sm, s, NXL, 0));           // [Amax] placed nucleons, single precision (the box fields are not live yet)
    float4* bq = reinterpret_cast<float4*>(sm.hit) + s * 32;               // [32] this batch's candidates (the hit masks are not live yet)
    while (placed < A) {
      while (nq < 32) {
        const uint32_t n = ndraw + lane;
        double r, cx = 0.0; bool ok;
        if (deformed) {                                    // Nucleus.cpp:585-607
          double u0, u1; smc_uniform2(s_ws, n, 0, &u0, &u1);
          const double u2 = smc_uniform(s_ws, n, 2);
          r = rmaxCut * cbrt(u0); cx = 1.0 - 2.0 * u1;
          const double rad1 = rad * (1.0 + c.beta2[s] * sph_harm2(cx) + c.beta4[s] * sph_harm4(cx));
          const double rwMax1 = 1.0 / (1.0 + exp(-rad1 / dr));
          ok = !(u2 * rwMax1 > 1.0 / (1.0 + exp((r - rad1) / dr)));
        } else {                                                // Nucleus.cpp:610-613
          double u1, u2; smc_uniform2(s_ws, n, 0, &u1, &u2);
          // the accept test in single precision (error of the right-hand side < 2e-5 relative); a draw within 2e-4 of the
          // boundary is settled by the reference's double expression, so every decision is the reference's.  The queue
          // keeps u1: the radius rmaxCut * cbrt(u1) is taken in double precision for the accepted draws only (below)
          const float rf = rmaxCut_f * cbrtf((float)u1);
          const float ff = __fdividef(1.0f, 1.0f + __expf((rf - rad_f) * inv_dr_f)), lf = (float)(u2 * rwMax);
          ok = !(lf > ff);
          if (fabsf(lf - ff) <= 2e-4f * ff) ok = !(u2 * rwMax > 1.0 / (1.0 + exp((rmaxCut * cbrt(u1) - rad) / dr)));
          r = u1;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) { const int pos = nq + __popc(m & below); qr[pos] = r; qc[pos] = cx; }
        nq += __popc(m); ndraw += 32;
        __syncwarp();
      }
      const uint32_t cand = cand_base + lane;
      const double r = deformed ? qr[lane] : rmaxCut * cbrt(qr[lane]), cxq = qc[lane];
      { double t0 = 0, t1 = 0; const bool mv = lane + 32 < nq;
        if (mv) { t0 = qr[lane + 32]; t1 = qc[lane + 32]; }
        __syncwarp();
        if (mv) { qr[lane] = t0; qc[lane] = t1; }
        nq -= 32; __syncwarp(); }
      double x, y, z;
      if (deformed) {
        const double cx = cxq, sx = sqrt(1.0 - cx * cx); double sp, cp;
        sincos(2 * SMC_PI * smc_uniform(s_an, cand, 1), &sp, &cp);
        x = r * sx * cp; y = r * sx * sp; z = r * cx;
        rot3(ctr, phir, x, y, z);
      } else {                                                  // Nucleus.cpp:614-619
        double ua, ub; smc_uniform2(s_an, cand, 0, &ua, &ub);
        double cx = 1.0 - 2.0 * ua, sx = sqrt(1.0 - cx * cx), sp, cp;
        sincos(2 * SMC_PI * ub, &sp, &cp);
        x = r * sx * cp; y = r * sx * sp; z = r * cx;
      }
      // Hard core (Nucleus.cpp:284-293).  The distance tests run in single precision on float4 copies of the
      // positions; a squared distance within 1e-4 of 0.81 (the single-precision error is < 1e-5) is settled by the
      // reference's double-precision expression, so every decision is the reference's.
      const float xf = (float)x, yf = (float)y, zf = (float)z;
      float r2min = 1e30f;
      for (int i = 0; i < placed; i++) {
        const float4 p = pf[i];
        const float ax = xf - p.x, ay = yf - p.y, az = zf - p.z;
        r2min = fminf(r2min, ax * ax + ay * ay + az * az);
      }
      bool bad = r2min < 0.81f - 1e-4f;
      if (__any_sync(0xffffffffu, !bad && r2min < 0.81f + 1e-4f)) {
        for (int i = 0; i < placed; i++) {
          double ax = x - S_(sm, s, NX, i), ay = y - S_(sm, s, NY, i), az = z - S_(sm, s, NZ, i);
          bad |= (ax * ax + ay * ay + az * az < rmin);
        }
      }
      // conflicts inside the batch: cm = earlier lanes closer than the core
      bq[lane] = make_float4(xf, yf, zf, 0.f);
      __syncwarp();
      unsigned cm = 0, dm = 0;
#pragma unroll kBatchUnroll
      for (int j = 0; j < 31; j++) {
        const float4 p = bq[j];
        const float ax = xf - p.x, ay = yf - p.y, az = zf - p.z, r2 = ax * ax + ay * ay + az * az;
        if (r2 < 0.81f + 1e-4f) { if (r2 < 0.81f - 1e-4f) cm |= 1u << j; else dm |= 1u << j; }
      }
      cm &= below; dm &= below;
      while (__any_sync(0xffffffffu, dm != 0)) {                // doubtful pairs: exact
        const int j = dm ? __ffs(dm) - 1 : lane;
        const double lx = __shfl_sync(0xffffffffu, x, j), ly = __shfl_sync(0xffffffffu, y, j), lz = __shfl_sync(0xffffffffu, z, j);
        if (dm) { const double ax = x - lx, ay = y - ly, az = z - lz; if (ax * ax + ay * ay + az * az < rmin) cm |= 1u << j; dm &= dm - 1; }
      }
      // sequential acceptance in candidate order: a lane without a conflict among the earlier admissible lanes is in;
      // the (few) others are in iff none of the lanes they conflict with was accepted
      const unsigned okb = __ballot_sync(0xffffffffu, !bad);
      unsigned todo = __ballot_sync(0xffffffffu, !bad && (cm & okb) != 0);
      unsigned acc = okb & ~todo;
      while (todo) {
        const int l = __ffs(todo) - 1; todo &= todo - 1;
        const unsigned cml = __shfl_sync(0xffffffffu, cm, l);
        if (!(cml & acc)) acc |= 1u << l;
      }
      const int rank = __popc(acc & below);
      if (((acc >> lane) & 1u) && placed + rank < A) {
        const int q = placed + rank;
        S_(sm, s, NX, q) = x; S_(sm, s, NY, q) = y; S_(sm, s, NZ, q) = z; S_(sm, s, NW, q) = (double)cand;
        pf[q] = make_float4(xf, yf, zf, 0.f);
      }
      const int nacc = min(__popc(acc), A - placed);
      placed += nacc; cand_base += 32;
      __syncwarp();
    }
  }
  __syncwarp();
  // centre of mass shift (Nucleus.cpp:301-309) + AABB + sort key
  double mx = 0, my = 0, mz = 0;
  if (recentre) {
    for (int k = lane; k < A; k += 32) { mx += S_(sm, s, NX, k); my += S_(sm, s, NY, k); mz += S_(sm, s, NZ, k); }
    mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz);
  }
  __syncwarp();
  for (int k = lane; k < A; k += 32) {
    double x0 = S_(sm, s, NX, k), y0 = S_(sm, s, NY, k), z0 = S_(sm, s, NZ, k);
    uint32_t cand = (uint32_t)S_(sm, s, NW, k);
    double* ex = st.nuc_extra_tmp ? st.nuc_extra_tmp + (((size_t)e * 2 + s) * c.Amax + k) * NEXTRA : nullptr;
    Box bx; particle_box<SPEC == 2>(c, st, s_q, cand, x0, y0, bx, ex);
    if (recentre) {
      double x = x0 - mx / A + xCenter, y = y0 - my / A + yCenter, z = z0 - mz / A;
      box_center(bx, x, bx.yC); box_center(bx, bx.xC, y);       // Particle::setX / setY, src/Particle.cpp:176-185
      x0 = x; y0 = y; z0 = z;
    }
    if (ex) { ex[XCX] = bx.xC; ex[XCY] = bx.yC; }
    S_(sm, s, NX, k) = x0; S_(sm, s, NY, k) = y0; S_(sm, s, NZ, k) = z0; S_(sm, s, NXL, k) = bx.xL; S_(sm, s, NXR, k) = bx.xR;
    S_(sm, s, NYL, k) = bx.yL; S_(sm, s, NYR, k) = bx.yR; S_(sm, s, NW, k) = 1.0;
  }
  __syncwarp();
  if (A <= 32) sort_by_xl<1>(c, st, sm, e, s, A, lane);
  else sort_by_xl<MK>(c, st, sm, e, s, A, lane);
}

// Marsaglia-Tsang gamma(shape a, scale th); a<1 boosted by U^(1/a).  Law-equivalent to gsl_ran_gamma
// (reference src/MCnucl.cpp:1285,1298); the reference's stream is a separate mt19937 (quirk Q8).
__device__ double gamma_variate(const smc_stream& s, uint32_t cand, double a, double th) {
  double boost = 1.0, aa = a;
  const bool small = a < 1.0;
  if (small) aa = a + 1.0;
  const double d = aa - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
  for (uint32_t it = 0; it < 64; it++) {
    double u1, u2, u3, u4;
    smc_uniform2(s, cand, 2 * it, &u1, &u2); smc_uniform2(s, cand, 2 * it + 1, &u3, &u4);
    // the normal deviate and the acceptance test run in single precision (the weight itself stays double)
    const float xf = sqrtf(-2.0f * __logf(fmaxf(1.0f - (float)u1, 1e-37f))) * cospif(2.0f * (float)u2);
    const double x = (double)xf;
    double v = 1.0 + cc * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const float df = (float)d, vf = (float)v;
    if (__logf(fmaxf(1.0f - (float)u3, 1e-37f)) < 0.5f * xf * xf + df - df * vf + df * __logf(vf)) {
      if (small) boost = pow(1.0 - u4, 1.0 / a);
      return d * v * th * boost;
    }
  }
  return a * th;
}

template <bool GIVEN, int SPEC, int MK>
__global__ void __launch_bounds__(64, 6) sample_collide_kernel(DevCfg c, Store st, int nev) {
  extern __shared__ double smem_d[];
  const int e = blockIdx.x;
  if (e >= nev) return;
  if (st.redo && !st.redo[e]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = c.A[0], B = c.A[1], Amax = c.Amax, HW = (Amax + 31) / 32;
  SampleSmem sm;
  sm.soa = smem_d; sm.Amax = Amax;
  sm.hit = (uint32_t*)(sm.soa + 2 * Amax * NROW + 256);
  sm.ncB = (int*)(sm.hit + max((size_t)Amax * HW, (size_t)256)); sm.firstB = sm.ncB + Amax; sm.rowoff = sm.firstB + Amax; sm.misc = sm.rowoff + Amax + 1;
  sm.wsq = smem_d + 2 * Amax * NROW;
  const uint64_t ev = st.event_id[e];
  double* gn = st.nuc + (size_t)e * 2 * Amax * NROW;
  int* hi = st.hdr_i + (size_t)e * HDR_I;
  double* hd = st.hdr_d + (size_t)e * HDR_D;
  double b = 0.0;
  uint32_t tr = GIVEN ? 0u : (uint32_t)st.try_start[e];
  int ncoll = 0, np1 = 0, np2 = 0;
  bool accepted = false;
  for (int guard = 0; guard < 100000 && !accepted; guard++, tr++) {
    if (GIVEN) {
      b = hd[HD_B];
      for (int k = tid; k < 2 * Amax * NROW; k += 64) { const int sd = k / (Amax * NROW), i = (k / NROW) % Amax, f = k % NROW; S_(sm, sd, f, i) = gn[k]; }
    } else {
      const smc_stream s_b = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_B, 0);
      b = sqrt((c.bmax * c.bmax - c.bmin * c.bmin) * smc_uniform(s_b, 0, 0) + c.bmin * c.bmin);   // MakeDensity.cpp:2149
      sample_nucleus<SPEC, MK>(c, st, sm, e, warp, ev, tr, warp == 0 ? b / 2.0 : -b / 2.0, 0.0);                 // MCnucl.cpp:208-214
    }
    // both nuclei must be complete before the hit masks are cleared: while a warp samples, the candidates of its
    // current batch live in that very region (bq in sample_nucleus) -- found by compute-sanitizer racecheck
    __syncthreads();
    for (int k = tid; k < Amax; k += 64) { sm.ncB[k] = 0; sm.firstB[k] = 0x7fffffff; }
    for (int k = tid; k < Amax * HW; k += 64) sm.hit[k] = 0;
    // Window [start, end) of target nucleons the sweep tests for every projectile row (MCnucl.cpp:253-270), one row per lane:
    //   start = first j with tXR_j >= pXL_i.  The reference carries its starting index from row to row, but rows are sorted by
    //           xL: whatever an earlier row skipped has tXR_j < pXL of that row <= pXL_i, so the carry does not change the result.
    //   end:    pair j is tested iff pXR_i >= tXL of the box looked at before it (the box at `start` for j = start and start + 1);
    //           targets are sorted by xL, so with U = #{j : tXL_j <= pXR_i} that is "start < U and j <= U".
    // Both come from binary searches on tXL (start: from the first box whose xL could reach pXL_i given the widest box, then the
    // exact predicate).  Packed into rowoff[] (free until the row counts are taken).
    {
      double wmax = 0.0;
#pragma unroll 1
      for (int j = lane; j < B; j += 32) wmax = fmax(wmax, S_(sm, 1, NXR, j) - S_(sm, 1, NXL, j));
      for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
#pragma unroll 1
      for (int i = tid; i < A; i += 64) {
        const double pXL = S_(sm, 0, NXL, i), pXR = S_(sm, 0, NXR, i);
        const double key = pXL - wmax * (1.0 + 1e-12) - 1e-12;       // boxes that start left of it end left of pXL
        int lo = 0, hi2 = B;
        while (lo < hi2) { const int mid = (lo + hi2) >> 1; if (S_(sm, 1, NXL, mid) < key) lo = mid + 1; else hi2 = mid; }
        while (lo < B && !(S_(sm, 1, NXR, lo) >= pXL)) lo++;           // skip loop of the sweep, MCnucl.cpp:255-261
        int ul = 0, uh = B;
        while (ul < uh) { const int mid = (ul + uh) >> 1; if (S_(sm, 1, NXL, mid) <= pXR) ul = mid + 1; else uh = mid; }
        const int end = (lo < ul) ? min(ul + 1, B) : lo;
        sm.rowoff[i] = lo | (end << 16);
      }
    }
    __syncthreads();
    // ---- collisions: rows of the projectile, 32 target nucleons per step ----
    const smc_stream s_p = smc_make_stream(c.seed_lo, c.seed_hi, ev, tr, SMC_K_PAIR, 0);
    const float hit_c1f = (float)(c.sigma_gg / (4. * SMC_PI * c.w * c.w)), hit_c2f = (float)(1.0 / (4. * c.w * c.w));
    const double* pu = (GIVEN && st.pair_u) ? st.pair_u + (size_t)e * A * B : nullptr;
    // pairs whose uniform is below a single-precision over-estimate of the hit probability wait in a per-warp queue
    // and are settled 32 at a time by the double-precision expression of the reference (dense lanes)
    double* qu = sm.wsq + (size_t)warp * 128; int* qij = reinterpret_cast<int*>(qu + 64);
    const unsigned below = (1u << lane) - 1u;
    int nq = 0;
    auto record_hit = [&](int i, int j) { atomicOr(&sm.hit[(size_t)i * HW + (j >> 5)], 1u << (j & 31)); atomicAdd(&sm.ncB[j], 1); atomicMin(&sm.firstB[j], i); };
    auto settle = [&](int cnt) {
      if (lane < cnt) {
        const int ij = qij[lane], i = ij >> 16, j = ij & 0xffff; const double u = qu[lane];
        const double ddx = S_(sm, 1, NX, j) - S_(sm, 0, NX, i), ddy = S_(sm, 1, NY, j) - S_(sm, 0, NY, i);
        const double bb = sqrt(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));              // MCnucl.cpp:359-360
        const double prob = 1. - exp(-c.sigma_gg * exp(-bb * bb / (4. * c.w * c.w)) / (4. * SMC_PI * c.w * c.w));
        if (u < prob) record_hit(i, j);
      }
      __syncwarp();
    };
    for (int i = warp; i < A; i += 2) {
      const int win = sm.rowoff[i], start = win & 0xffff, end = win >> 16;
      if (start >= B) break;                                    // rows are sorted by xL: no later row finds a start either
      if (start >= end) continue;
      const double px = S_(sm, 0, NX, i), py = S_(sm, 0, NY, i), pYL = S_(sm, 0, NYL, i), pYR = S_(sm, 0, NYR, i);
      for (int j0 = start; j0 < end; j0 += 32) {
        {
          const int j = j0 + lane; const bool tst = j < end;
          const int jj = tst ? j : 0;
          bool maybe = false; double u = 0.0;
          if (tst && pYL <= S_(sm, 1, NYR, jj) && pYR >= S_(sm, 1, NYL, jj)) {
            const double ddx = S_(sm, 1, NX, jj) - px, ddy = S_(sm, 1, NY, jj) - py;
            if (c.crit == 1) {
              const double bb = sqrt(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));        // MCnucl.cpp:359-366
              if (__dmul_rn(bb, bb) <= c.dsq) record_hit(i, j);
            } else if (SPEC == 0 && c.crit == 3) {
              // GaussianNucleonsCal::testFluctuatedCollision (GaussianNucleonsCal.cpp:70-97): overlap of the 3 x 3 valence-quark
              // Gaussians (offsets from the sorted extras rows, written before the __syncthreads above or staged by the host)
              const double* qa = st.nuc_extra + (((size_t)e * 2 + 0) * Amax + i) * NEXTRA + XQ;
              const double* qb = st.nuc_extra + (((size_t)e * 2 + 1) * Amax + jj) * NEXTRA + XQ;
              const double tx = S_(sm, 1, NX, jj), ty = S_(sm, 1, NY, jj), gw2 = c.quark_width * c.quark_width;
              double overlap = 0;
              for (int a = 0; a < 3; a++) for (int q = 0; q < 3; q++) {
                const double mx = qa[3 * a] + px, my = qa[3 * a + 1] + py, yx = qb[3 * q] + tx, yy = qb[3 * q + 1] + ty;
                const double d = (mx - yx) * (mx - yx) + (my - yy) * (my - yy);
                overlap += (1 / (4 * SMC_PI * gw2)) * exp(-d / (4 * gw2)) / 9;
              }
              u = pu ? pu[(size_t)i * B + j] : smc_uniform(s_p, (uint32_t)i, (uint32_t)j);
              if (u < 1. - exp(-c.sigma_gg * overlap)) record_hit(i, j);
            } else {
              u = pu ? pu[(size_t)i * B + j] : smc_uniform(s_p, (uint32_t)i, (uint32_t)j);
              // P = 1 - exp(-t) <= t
              const float tub = hit_c1f * __expf(-(float)(ddx * ddx + ddy * ddy) * hit_c2f) * 1.01f;
              maybe = (u <= (double)tub);
            }
          }
          const unsigned mm = __ballot_sync(0xffffffffu, maybe);
          if (mm) {
            if (maybe) { const int pos = nq + __popc(mm & below); qij[pos] = (i << 16) | j; qu[pos] = u; }
            nq += __popc(mm);
            __syncwarp();
            if (nq >= 32) {
              settle(32);
              int t0 = 0; double t1 = 0.0; const bool mv = lane + 32 < nq;
              if (mv) { t0 = qij[lane + 32]; t1 = qu[lane + 32]; }
              __syncwarp();
              if (mv) { qij[lane] = t0; qu[lane] = t1; }
              nq -= 32; __syncwarp();
            }
          }
        }
      }
    }
    settle(nq);
    __syncthreads();
    for (int i = tid; i < A; i += 64) { int r = 0; for (int w = 0; w < HW; w++) r += __popc(sm.hit[(size_t)i * HW + w]); sm.rowoff[i] = r; }
    __syncthreads();
    // ---- counts ----
    if (warp == 0) {
      int n1 = 0, nc = 0;
      for (int i = lane; i < A; i += 32) { int r = sm.rowoff[i]; nc += r; n1 += (r > 0); }
      int n2 = 0;
      for (int j = lane; j < B; j += 32) n2 += (sm.ncB[j] > 0);
      for (int o = 16; o > 0; o >>= 1) { n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o); nc += __shfl_xor_sync(0xffffffffu, nc, o); }
      if (lane == 0) { sm.misc[0] = n1; sm.misc[1] = n2; sm.misc[2] = nc; }
    }
    __syncthreads();
    np1 = sm.misc[0]; np2 = sm.misc[1]; ncoll = sm.misc[2];
    accepted = (ncoll > 0) && (np1 + np2 <= c.npmax) && (np1 + np2 >= c.npmin);                     // MakeDensity.cpp:2147, MCnucl.cpp:388-393
    if (GIVEN) { tr++; break; }
    __syncthreads();
  }
  // ---- emit the event record ----
  if (tid == 0) {
    hi[H_NP1] = np1; hi[H_NP2] = np2; hi[H_NCOLL] = ncoll; hi[H_TRIES] = (st.redo ? hi[H_TRIES] : 0) + (int)tr - (GIVEN ? 0 : st.try_start[e]);
    hi[H_STATUS] = accepted ? (ncoll > c.ncoll_cap ? 4 : 0) : 100;
    hd[HD_B] = b;
    if (!GIVEN) st.try_start[e] = (int)tr;
  }
  const uint32_t trw = tr - 1;    // the accepted try
  // exclusive prefix of row hit counts -> collision offsets in (i,j) order (createBinaryCollisions, MCnucl.cpp:326-352)
  if (warp == 0) {
    int run = 0;
    for (int i0 = 0; i0 < A; i0 += 32) {
      int i = i0 + lane; int v = (i < A) ? sm.rowoff[i] : 0, incl = v;
      for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
      if (i < A) sm.rowoff[i] = run + incl - v;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) sm.rowoff[A] = run;
  }
  __syncthreads();
  const bool givenw = GIVEN && hi[H_GIVENW];
  for (int k = tid; k < A + B; k += 64) {
    const int s = k >= A, i = s ? k - A : k;
    const int nc = s ? sm.ncB[i] : (sm.rowoff[i + 1] - sm.rowoff[i]);
    if (!givenw) S_(sm, s, NW, i) = 1.0;
    st.nuc_ncoll[((size_t)e * 2 + s) * Amax + i] = nc;
    if (s) st.nuc_first[(size_t)e * Amax + i] = sm.firstB[i];
  }
  // compact participant / spectator lists (ordered), warp 0 = proj, warp 1 = targ
  int* pidx = st.part_idx + (size_t)e * 2 * Amax;
  {
    const int s = warp, n = c.A[s];
    int npart = 0, nspec = 0;
    const int pbase = s ? np1 : 0, sbase = s ? (A - np1) : 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const bool in = i < n;
      const int nc = in ? (s ? sm.ncB[i] : (sm.rowoff[i + 1] - sm.rowoff[i])) : 0;
      const unsigned mp = __ballot_sync(0xffffffffu, in && nc > 0), ms = __ballot_sync(0xffffffffu, in && nc == 0);
      const unsigned below = (1u << lane) - 1u;
      if (in && nc > 0) pidx[pbase + npart + __popc(mp & below)] = (s << 16) | i;
      if (in && nc == 0) st.spec_idx[(size_t)e * 2 * Amax + sbase + nspec + __popc(ms & below)] = (s << 16) | i;
      npart += __popc(mp); nspec += __popc(ms);
    }
    if (lane == 0) hi[s ? H_NSPEC2 : H_NSPEC1] = nspec;
  }
  // collision list in (i,j) order (createBinaryCollisions, MCnucl.cpp:326-352): the sparse pass only records the pair
  int* cij = st.coll_ij + (size_t)e * c.ncoll_cap;
  if (accepted) {
    for (int i = tid; i < A; i += 64) {            // one projectile row per lane: its hit words in order, bit by bit
      int off = sm.rowoff[i];
      if (sm.rowoff[i + 1] == off) continue;
#pragma unroll 1
      for (int wj = 0; wj < HW; wj++) {
        unsigned hm = sm.hit[(size_t)i * HW + wj];
        while (hm) { const int b = __ffs(hm) - 1; hm &= hm - 1; if (off < c.ncoll_cap) cij[off] = (i << 16) | (wj * 32 + b); off++; }
      }
    }
  }
  __syncthreads();
  // Gamma multiplicity weights, one variate per lane (dense over the compact lists).  Nucleon weights
  // (selectFluctFactors, MCnucl.cpp:310-324) are re-drawn at every hit, last wins => one draw per wounded nucleon.
  if (!givenw && c.cc_fluct > 5 && accepted) {
    if (SPEC == 0 && c.shape_of_entropy == 3) {
      // one weight per valence quark, shape k/3 (selectFluctFactors / sampleFluctuationFactorforParticipant, MCnucl.cpp:310-324,
      // 1271-1288); the nucleon's own factor stays 1
      for (int k = tid; k < 3 * (np1 + np2); k += 64) {
        const int id = pidx[k / 3], s = id >> 16, i = id & 0xffff, q = k % 3;
        const smc_stream sg = smc_make_stream(c.seed_lo, c.seed_hi, ev, trw, SMC_K_GAMMA_PART, s);
        st.nuc_extra[(((size_t)e * 2 + s) * Amax + i) * NEXTRA + XF + q] = gamma_variate(sg, (uint32_t)(3 * i + q), c.gam_k_part / 3.0, c.gam_th_part);
      }
    } else {
      for (int k = tid; k < np1 + np2; k += 64) {
        const int id = pidx[k], s = id >> 16, i = id & 0xffff;
        const smc_stream sg = smc_make_stream(c.seed_lo, c.seed_hi, ev, trw, SMC_K_GAMMA_PART, s);
        S_(sm, s, NW, i) = gamma_variate(sg, (uint32_t)i, c.gam_k_part, c.gam_th_part);
      }
    }
  }
  // collisions: midpoints + weights
  if (accepted) {
    const smc_stream sgc = smc_make_stream(c.seed_lo, c.seed_hi, ev, trw, SMC_K_GAMMA_COLL, 0);
    const double* cw = (GIVEN && st.coll_w) ? st.coll_w + (size_t)e * c.ncoll_cap * 2 : nullptr;
    const int nck = min(ncoll, c.ncoll_cap);
    for (int k = tid; k < nck; k += 64) {
      const int ij = cij[k], i = ij >> 16, j = ij & 0xffff;
      double* cr = st.coll + ((size_t)e * c.ncoll_cap + k) * CROW;
      cr[CX] = (S_(sm, 0, NX, i) + S_(sm, 1, NX, j)) / 2.0;                                           // MCnucl.cpp:339-340
      cr[CY] = (S_(sm, 0, NY, i) + S_(sm, 1, NY, j)) / 2.0;
      double wv = 1.0, addw = 0.0;
      if (c.which_mc_model == 5 && c.sub_model == 2)                                           // integer division in the reference,
        addw = (double)((sm.rowoff[i + 1] - sm.rowoff[i] == 1 ? 1 : 0) + (sm.ncB[j] == 1 ? 1 : 0));   // MCnucl.cpp:345-348
      if (cw) { wv = cw[2 * k]; addw = cw[2 * k + 1]; }
      else if (c.cc_fluct > 5) wv = gamma_variate(sgc, (uint32_t)k, c.gam_k_bin, c.gam_th_bin);
      cr[CW] = wv; cr[CADDW] = addw;
    }
  }
  __syncthreads();
  // rows [side][i][NROW] in global memory <- structure of arrays in shared memory: one 64-byte row per thread step
#pragma unroll 1
  for (int sd = 0; sd < 2; sd++)
#pragma unroll 1
    for (int i = tid; i < Amax; i += 64) {
      double2* row = reinterpret_cast<double2*>(gn + ((size_t)sd * Amax + i) * NROW);
#pragma unroll
      for (int f = 0; f < NROW; f += 2) row[f >> 1] = make_double2(S_(sm, sd, f, i), S_(sm, sd, f + 1, i));
    }
}

size_t sample_smem_bytes(int Amax) {
  const int HW = (Amax + 31) / 32;
  size_t d = (size_t)(2 * Amax * NROW) + 256;
  size_t i = std::max((size_t)Amax * HW, (size_t)256) + 3 * Amax + 1 + 16;
  return d * sizeof(double) + i * sizeof(int);
}

template <bool GIVEN, int SPEC, int MK>
static cudaError_t launch_sc(const DevCfg& c, const Store& st, int nev, size_t smem, cudaStream_t s) {
  cudaFuncSetAttribute(sample_collide_kernel<GIVEN, SPEC, MK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(sample_collide_kernel<GIVEN, SPEC, MK>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  sample_collide_kernel<GIVEN, SPEC, MK><<<nev, 64, smem, s>>>(c, st, nev);
  return cudaGetLastError();
}

cudaError_t launch_sample_collide(const DevCfg& c, const Store& st, int nev, bool given, cudaStream_t s) {
  static const size_t pad = getenv("SMC_SAMPLE_PAD") ? (size_t)atoi(getenv("SMC_SAMPLE_PAD")) : 0;     // tuning aid: occupancy sensitivity
  static const bool generic = getenv("SMC_SAMPLE_GENERIC") != nullptr;                                   // A/B: the unspecialised kernel
  const size_t smem = sample_smem_bytes(c.Amax) + pad;
  const bool big = c.Amax > 256;
  if (given) return big ? launch_sc<true, 0, SMC_MAXK>(c, st, nev, smem, s) : launch_sc<true, 0, 8>(c, st, nev, smem, s);
  const bool ws = !generic && c.sampler[0] == 0 && c.sampler[1] == 0 && !c.deformed[0] && !c.deformed[1] && c.crit != 3 && c.shape_of_entropy != 3;
  const int spec = ws ? (c.quark_rows > 0 ? 1 : 2) : 0;
  if (big) return spec == 2 ? launch_sc<false, 2, SMC_MAXK>(c, st, nev, smem, s) : spec == 1 ? launch_sc<false, 1, SMC_MAXK>(c, st, nev, smem, s) : launch_sc<false, 0, SMC_MAXK>(c, st, nev, smem, s);
  return spec == 2 ? launch_sc<false, 2, 8>(c, st, nev, smem, s) : spec == 1 ? launch_sc<false, 1, 8>(c, st, nev, smem, s) : launch_sc<false, 0, 8>(c, st, nev, smem, s);
}

}  // namespace smc
