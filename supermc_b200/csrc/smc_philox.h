// smc_philox.h -- Philox4x32-10 counter-based uniforms (Salmon, Moraes, Dror, Shaw, SC'11) and the
// counter layout of the superMC event streams.  Host + device.
//
// The reference draws everything from three global sequential generators (drand48, rand, mt19937;
// SURVEY.md quirk Q8).  Here every uniform has an address
//     (seed, event id, try, nucleus, kind, cand, slot)
// so any thread can produce it and the set of events does not depend on batch size or GPU count:
//     key  = (seed_lo, seed_hi)
//     ctr0 = event id, low 32      ctr1 = event id, high 32
//     ctr2 = try << 8 | kind << 1 | nucleus
//     ctr3 = cand << 12 | slot >> 1          component (slot & 1) of the 4x32 output
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SMC_HD __host__ __device__ __forceinline__
#else
#define SMC_HD static inline
#endif

enum { SMC_K_B = 0, SMC_K_ORIENT = 1, SMC_K_WS = 2, SMC_K_ANGLE = 3, SMC_K_QUARK = 4, SMC_K_PAIR = 5,
       SMC_K_GAMMA_PART = 6, SMC_K_GAMMA_COLL = 7, SMC_K_CONFIG = 8, SMC_K_DEUT = 9, SMC_K_NBD = 10 };

struct smc_u4 { uint32_t v[4]; };

SMC_HD uint32_t smc_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

SMC_HD smc_u4 smc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = smc_mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = smc_mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  smc_u4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3; return o;
}

// 53-bit uniform in [0,1) from two 32-bit words
SMC_HD double smc_u53(uint32_t a, uint32_t b) {
  uint64_t m = ((uint64_t)a << 21) | ((uint64_t)b >> 11);
  return (double)m * (1.0 / 9007199254740992.0);
}

struct smc_stream {   // everything but (cand, slot)
  uint32_t k0, k1, c0, c1, c2;
};
SMC_HD smc_stream smc_make_stream(uint32_t seed_lo, uint32_t seed_hi, uint64_t event, uint32_t tr, int kind, int nuc) {
  smc_stream s; s.k0 = seed_lo; s.k1 = seed_hi; s.c0 = (uint32_t)event; s.c1 = (uint32_t)(event >> 32);
  s.c2 = (tr << 8) | ((uint32_t)kind << 1) | (uint32_t)nuc; return s;
}
// both uniforms of one Philox call: slots 2q and 2q+1
SMC_HD void smc_uniform2(const smc_stream& s, uint32_t cand, uint32_t q, double* u0, double* u1) {
  smc_u4 o = smc_philox4x32_10(s.c0, s.c1, s.c2, (cand << 12) | q, s.k0, s.k1);
  *u0 = smc_u53(o.v[0], o.v[1]); *u1 = smc_u53(o.v[2], o.v[3]);
}
SMC_HD double smc_uniform(const smc_stream& s, uint32_t cand, uint32_t slot) {
  double a, b; smc_uniform2(s, cand, slot >> 1, &a, &b); return (slot & 1) ? b : a;
}
// per-lattice-cell uniform of the NBD multiplicity fluctuations: ctr3 is the cell index itself (a lattice can have more
// than 2^20 cells), ctr2 = pass << 8 | SMC_K_NBD << 1 where `pass` counts the re-deposits of an event (operation 3)
SMC_HD double smc_uniform_cell(uint32_t seed_lo, uint32_t seed_hi, uint64_t event, uint32_t pass, uint32_t cell) {
  smc_u4 o = smc_philox4x32_10((uint32_t)event, (uint32_t)(event >> 32), (pass << 8) | ((uint32_t)SMC_K_NBD << 1), cell, seed_lo, seed_hi);
  return smc_u53(o.v[0], o.v[1]);
}
