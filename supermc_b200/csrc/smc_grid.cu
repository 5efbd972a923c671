// smc_grid.cu -- K3 density accumulation and K4 eccentricity moments.
//
// K3 replaces MCnucl::setThickness, addDensity, the binary-collision loop of setDensity,
// calculate_rho_binary and calculate_spectator_density (reference src/MCnucl.cpp:432-478, 481-531,
// 534-614, 688-811, 822-866; Particle::getSmoothTn src/Particle.cpp:122-132).
// K4 replaces MakeDensity::dumpEccentricities / MCnucl::getHotSpots (src/MakeDensity.cpp:2244-2430,
// src/MCnucl.cpp:1303-1323) and GlueDensity::calcCMAngle (src/GlueDensity.cpp:87-144).
//
// B200 mapping of K3 (gather, no atomics): the reference scatters every source over its window with one
// exp per (source, cell).  Here bbox_kernel expands every source once per event into a record (position,
// folded weight, exact window); a deposit CTA owns DEP_BAND rows x DEP_COLS columns of one event, each warp a
// 16 x 32 tile and each lane a 4 x 4 micro-tile in registers.  Sources whose window meets the CTA are compacted
// in order and processed in chunks: per chunk the CTA builds, in shared memory, the *separable* factors
// W*exp(-dx^2/2w^2) (rows) and exp(-dy^2/2w^2) (columns) with a two-multiply Gaussian recurrence (a few exps
// per source instead of ~2000) and, per (source, row), the column interval of the reference's circle test
// `(x-xg)^2 + (y-yg)^2 <= thr` as bit masks (single-precision estimate, settled by the exact double predicate
// near cell boundaries: the masks are the reference's bit for bit).  A lane then spends 5 shared loads,
// 4 shifts, 4 R2P and 16 predicated DFMAs per (source, tile).  Tables of chunk n+1 are built by whichever warps
// finish chunk n first (shared item counter, one barrier per chunk).  Deposits are deterministic (fixed
// summation order) and each grid cell is written exactly once.
#include <algorithm>
#include <cstdlib>
#include <cuda_pipeline.h>
#include "smc_common.cuh"

namespace smc {

enum { GK_RHO = 0, GK_TA1 = 1, GK_TA2 = 2, GK_RHO_BINARY = 3, GK_SPEC_A = 4, GK_SPEC_B = 5, GK_RHOA = 6, GK_RHOB = 7 };

struct Src { double x, y, W, thr; int iL, iR, jL, jR; int flat; };

__device__ __forceinline__ int src_count(const DevCfg& c, const int* hi, int kind) {
  const int np1 = hi[H_NP1], np2 = hi[H_NP2];
  int nc = hi[H_NCOLL]; if (nc > c.ncoll_cap) nc = c.ncoll_cap;
  const int nq = (c.shape_of_entropy == 3) ? 3 : 1;        // three valence-quark sources per wounded nucleon (addDensity, MCnucl.cpp:856)
  switch (kind) {
    case GK_RHO: return ((c.sub_model == 1) ? nq * (np1 + np2) : 0) + ((c.alpha > 1e-8) ? nc : 0);
    case GK_RHOA: return nq * np1;
    case GK_RHOB: return nq * np2;
    case GK_TA1: return np1;
    case GK_TA2: return np2;
    case GK_RHO_BINARY: return nc;
    case GK_SPEC_A: return hi[H_NSPEC1];
    case GK_SPEC_B: return hi[H_NSPEC2];
  }
  return 0;
}

// source k of `kind` for event e -> position, folded weight, window, mask threshold
__device__ void load_src(const DevCfg& c, const Store& st, int e, const int* hi, int kind, int k, Src& s) {
  const int Amax = c.Amax;
  const double* nuc = st.nuc + (size_t)e * 2 * Amax * NROW;
  const int np1 = hi[H_NP1], np2 = hi[H_NP2];
  bool box_window = false, is_coll = false;
  const double* row = nullptr; const double* quark = nullptr;      // quark: extras row + which valence quark (shape_of_entropy 3)
  int qi = 0;
  const int nq = (c.shape_of_entropy == 3) ? 3 : 1;
  double wgt = 1.0;
  int shape = c.shape_of_nucleons;          // which "shape" switch the reference consults for this deposit
  s.thr = c.thrB;
  if (kind == GK_RHO) {
    const int nwn = (c.sub_model == 1) ? nq * (np1 + np2) : 0;
    shape = c.shape_of_entropy;
    if (k < nwn) {                                                                   // addDensity, MCnucl.cpp:822-866
      const int id = st.part_idx[(size_t)e * 2 * Amax + k / nq];
      row = nuc + ((size_t)(id >> 16) * Amax + (id & 0xffff)) * NROW;
      wgt = row[NW] * ((1.0 - c.alpha) / 2.); box_window = true; s.thr = c.thrA;
      if (nq == 3) { quark = st.nuc_extra + (((size_t)e * 2 + (id >> 16)) * Amax + (id & 0xffff)) * NEXTRA; qi = k % 3; wgt = quark[XF + qi] * ((1.0 - c.alpha) / 2.); }
    } else {                                                                         // MCnucl.cpp:724-759
      const double* cr = st.coll + ((size_t)e * c.ncoll_cap + (k - nwn)) * CROW;
      row = cr; is_coll = true;
      const double fl = (c.cc_fluct > 5) ? cr[CW] : 1.0;
      wgt = fl * (c.alpha + (1. - c.alpha) * cr[CADDW]); s.thr = c.thrB;
    }
  } else if (kind == GK_RHOA || kind == GK_RHOB) {                                   // model 7, MCnucl.cpp:780-811
    const int id = st.part_idx[(size_t)e * 2 * Amax + (kind == GK_RHOB ? np1 : 0) + k / nq];
    row = nuc + ((size_t)(id >> 16) * Amax + (id & 0xffff)) * NROW;
    wgt = row[NW]; box_window = true; s.thr = c.thrA; shape = c.shape_of_entropy;
    if (nq == 3) { quark = st.nuc_extra + (((size_t)e * 2 + (id >> 16)) * Amax + (id & 0xffff)) * NEXTRA; qi = k % 3; wgt = quark[XF + qi]; }
  } else if (kind == GK_TA1 || kind == GK_TA2) {                                     // setThickness, MCnucl.cpp:432-478
    const int id = st.part_idx[(size_t)e * 2 * Amax + (kind == GK_TA2 ? np1 : 0) + k];
    row = nuc + ((size_t)(id >> 16) * Amax + (id & 0xffff)) * NROW;
    s.thr = c.thrA;
  } else if (kind == GK_RHO_BINARY) {                                                // MCnucl.cpp:481-531
    row = st.coll + ((size_t)e * c.ncoll_cap + k) * CROW; is_coll = true;
  } else {                                                                           // spectators, MCnucl.cpp:534-614
    const int id = st.spec_idx[(size_t)e * 2 * Amax + (kind == GK_SPEC_B ? (c.A[0] - np1) : 0) + k];
    row = nuc + ((size_t)(id >> 16) * Amax + (id & 0xffff)) * NROW;
  }
  s.x = row[0]; s.y = row[1];
  s.flat = (shape == 1);
  if (s.flat) { s.W = wgt * c.areai; s.thr = c.dsq; } else s.W = wgt * c.norm;
  if (quark) {      // Quark::getSmoothTn (Quark.cpp:14-22): a Gaussian of width quark_width at nucleon + offset, cut at d > 5 * width (sic)
    s.x = row[0] + quark[XQ + 3 * qi]; s.y = row[1] + quark[XQ + 3 * qi + 1];
    s.flat = 2; s.W = wgt * c.q_norm; s.thr = c.q_thr;
  }
  // +-d_max window (quirk Q7); for AABB windows (quirk Q3) the same range, widened by a cell, is only a
  // safe clip: cells farther than the mask radius contribute nothing in the reference either
  const double reach = quark ? c.q_reach : (box_window && s.flat) ? c.rclip_flat : c.dmax;
  int iL = cell_of(__dadd_rn(s.x, -reach), c.Xmin, c.dx), iR = cell_of(__dadd_rn(s.x, reach), c.Xmin, c.dx);
  int jL = cell_of(__dadd_rn(s.y, -reach), c.Ymin, c.dy), jR = cell_of(__dadd_rn(s.y, reach), c.Ymin, c.dy);
  if (box_window && !is_coll) {
    const int bl = cell_of(row[NXL], c.Xmin, c.dx), br = cell_of(row[NXR], c.Xmin, c.dx);
    const int cl = cell_of(row[NYL], c.Ymin, c.dy), cr2 = cell_of(row[NYR], c.Ymin, c.dy);
    iL = max(bl, iL - 1); iR = min(br, iR + 2); jL = max(cl, jL - 1); jR = min(cr2, jR + 2);
  }
  s.iL = max(0, iL); s.iR = min(c.Maxx, iR); s.jL = max(0, jL); s.jR = min(c.Maxy, jR);
}

struct KindList { int n; int kind[8]; };

// Deposit geometry.  A CTA owns DEP_BAND rows x DEP_COLS columns of one event's lattice; warp (wr, ws) owns a
// 16 x 32 tile of it and every lane a 4 x 4 register micro-tile (rows 4*lr..4*lr+3, columns 4*lc..4*lc+3 of the
// tile), so one source costs a lane 5 shared-memory loads, 4 shifts, 4 R2P and 16 predicated DFMAs.  Sources are processed in chunks of DEP_CH whose tables live in shared memory:
//   xg[t][r]      W * exp(-dx^2/2w^2) of row r            (two-multiply Gaussian recurrence, DEP_XP rows per item)
//   yg[t][c]      exp(-dy^2/2w^2) of column c             (same, DEP_YP columns per item)
//   mk[t][s][r]   32-bit mask of row r over stripe s: window  AND  the reference's circle test
// The circle test  (x-xg)^2 + (y-yg)^2 <= thr  (each term rounded as the reference rounds it) is monotone in
// |y-yg|, so per (source,row) it is a column interval.  Its end points come from a single-precision estimate and
// are settled with the exact double-precision predicate whenever the estimate lies within 2e-3 cells of a cell
// boundary (the estimate's own error is < 1e-4 cells), i.e. the masks are the reference's, bit for bit.
#ifndef DEP_BAND
#define DEP_BAND 32
#endif
#define DEP_WROWS 16
#define DEP_NWR (DEP_BAND / DEP_WROWS)
#ifndef DEP_NSTR
#define DEP_NSTR 4
#endif
#define DEP_COLS (DEP_NSTR * 32)
#ifndef DEP_CH
#define DEP_CH 16
#endif
#define DEP_XP 4
#ifndef DEP_YP
#define DEP_YP 64      // columns of yg one build item covers (16 / 32 / 64 / 128 measured: 745 / 748 / 756 / 704 k events/s with DEP_XG 16 / 16 / 16 / 32)
#endif
#ifndef DEP_MINCTA
#define DEP_MINCTA 3
#endif
#ifndef DEP_GRAB
#define DEP_GRAB 32     // items a warp takes per shared-counter grab
#endif
#define DEP_THREADS (DEP_NWR * DEP_NSTR * 32)
#ifndef SMC_DEP_PERSIST_DEFAULT
#define SMC_DEP_PERSIST_DEFAULT 1     // 1: persistent CTAs over bbox_kernel's tile list; k > 1: k x as many CTAs as fit
#endif
#ifndef DEP_XG
#define DEP_XG 16       // rows of xg one build item covers (one exp pair + a DEP_XG-step recurrence)
#endif
#define DEP_NXG (DEP_BAND / DEP_XG)
#define DEP_NXI (DEP_BAND / DEP_XP)
#define DEP_NYI (DEP_COLS / DEP_YP)

// rows are padded by 16 bytes: builder lanes work on different sources t at the same row/column, and a row
// pitch that is a multiple of 128 bytes would put all of their stores into the same banks
struct DepTab {
  double xg[DEP_CH][DEP_BAND + 2];
  uint32_t mk[DEP_CH][DEP_NSTR * DEP_BAND + 4];     // [t][s * DEP_BAND + r]
  double yg[DEP_CH][DEP_COLS + 2];
  int4 desc[DEP_CH];                         // iL, iR, jL, jR
};

// Tile lists of the persistent deposit launch: (event, kind index, column group, band) of every tile that exists,
// in DEP_NCLS weight classes (class 0 = the middle bands of an event's rectangle, the last class = its edges) that
// are walked in order, heaviest first, so that the launch does not end on a heavy tile.  bbox_kernel appends the
// tiles of its event; the order inside a class does not matter: every tile is written by exactly one CTA and its
// result does not depend on who computes it or when.
#define DEP_NCLS 8
struct DepWork { int2* items; int* ctr; int cap; };      // ctr[0] = next tile to hand out, ctr[1 + c] = tiles in class c
__device__ __forceinline__ DepWork dep_work(const Store& st) {
  DepWork w; w.items = reinterpret_cast<int2*>(reinterpret_cast<char*>(st.src_rec) + st.work_off); w.cap = st.work_cap;
  w.ctr = reinterpret_cast<int*>(w.items + (size_t)DEP_NCLS * st.work_cap); return w;
}
__device__ __forceinline__ int dep_tile_class(int band, int nb, int sgroup) { return min(abs(2 * band - (nb - 1)) + 3 * sgroup, DEP_NCLS - 1); }

// bounding rectangles (cells) of the source windows of one event: one for the participant / collision deposits, one for the
// spectator deposits (they cover different parts of the lattice); nothing outside them is ever written or read
__global__ void bbox_kernel(DevCfg c, Store st, KindList kl, int nev, int list_tiles) {
  __shared__ int red[4][8];
  __shared__ int s_cnt[10], s_nb[2];
  const int e = blockIdx.x + st.e0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (st.redo && !st.redo[e]) return;
  int* hi = st.hdr_i + (size_t)e * HDR_I;
  int lo[4] = {c.Maxx, c.Maxy, c.Maxx, c.Maxy}, hh[4] = {0, 0, 0, 0};      // (rows, columns) of the two rectangles
  const int status = hi[H_STATUS];
  if (status == 0 || status == 4) {
    for (int q = 0; q < kl.n; q++) {
      const int kind = kl.kind[q];
      const int w = (kind == GK_SPEC_A || kind == GK_SPEC_B) ? 2 : 0;
      const int ns = src_count(c, hi, kind);
      SrcRec* recs = st.src_rec + ((size_t)e * kl.n + q) * st.src_stride;
      for (int k = tid; k < ns; k += blockDim.x) {
        Src s; load_src(c, st, e, hi, kind, k, s);
        SrcRec r; r.x = s.x; r.y = s.y; r.W = s.W; r.thr = s.thr; r.iL = (short)s.iL; r.iR = (short)s.iR; r.jL = (short)s.jL; r.jR = (short)s.jR; r.flat = s.flat; r.pad = 0;
        recs[k] = r;
        if (s.iL < s.iR && s.jL < s.jR) { lo[w] = min(lo[w], s.iL); hh[w] = max(hh[w], s.iR); lo[w + 1] = min(lo[w + 1], s.jL); hh[w + 1] = max(hh[w + 1], s.jR); }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++)
    for (int o = 16; o > 0; o >>= 1) { lo[q] = min(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], o)); hh[q] = max(hh[q], __shfl_xor_sync(0xffffffffu, hh[q], o)); }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 4; q++) { red[warp][q] = lo[q]; red[warp][4 + q] = hh[q]; }
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
      for (int q = 0; q < 4; q++) { lo[q] = min(lo[q], red[w][q]); hh[q] = max(hh[q], red[w][4 + q]); }
    for (int w = 0; w < 4; w += 2) if (hh[w] <= lo[w] || hh[w + 1] <= lo[w + 1]) { lo[w] = hh[w] = lo[w + 1] = hh[w + 1] = 0; }
    hi[H_RLO] = lo[0]; hi[H_RHI] = hh[0]; hi[H_CLO] = lo[1]; hi[H_CHI] = hh[1];
    hi[H_SRLO] = lo[2]; hi[H_SRHI] = hh[2]; hi[H_SCLO] = lo[3]; hi[H_SCHI] = hh[3];
    if (list_tiles) {
      int tot = 0;
      for (int w = 0; w < 2; w++) s_nb[w] = (hh[2 * w] - lo[2 * w] + DEP_BAND - 1) / DEP_BAND;
      for (int q = 0; q < kl.n; q++) {
        const int w = (kl.kind[q] == GK_SPEC_A || kl.kind[q] == GK_SPEC_B) ? 1 : 0;
        s_cnt[q] = tot;
        if (status == 0 || status == 4) tot += s_nb[w] * ((hh[2 * w + 1] - lo[2 * w + 1] + DEP_COLS - 1) / DEP_COLS);
      }
      s_cnt[kl.n] = tot;
    }
  }
  if (!list_tiles) return;
  __syncthreads();
  const int tot = s_cnt[kl.n];
  const DepWork wk = dep_work(st);
  for (int k = tid; k < tot; k += blockDim.x) {
    int q = 0;
    while (k >= s_cnt[q + 1]) q++;
    const int w = (kl.kind[q] == GK_SPEC_A || kl.kind[q] == GK_SPEC_B) ? 1 : 0;
    const int t = k - s_cnt[q], nb = s_nb[w], band = t % nb, sgroup = t / nb;
    const int cls = dep_tile_class(band, nb, sgroup);
    const int pos = atomicAdd(wk.ctr + 1 + cls, 1);
    wk.items[(size_t)cls * wk.cap + pos] = make_int2(e, (q << 16) | (sgroup << 8) | band);
  }
}

__device__ __forceinline__ uint32_t ones_below(int n) {      // bits [0, n) of a word, n clamped to [0, 32]
  return __funnelshift_lc(0xffffffffu, 0u, (uint32_t)n);
}

// one row of a lane's micro-tile: four DFMAs guarded by bits 0..3 of the (pre-shifted) row mask.  Written as PTX
// *branches*: ptxas turns those into predicated DFMAs fed by one R2P, whereas `if (bit) a = fma(..)` (or a PTX
// `@p fma`) comes out as an unconditional DFMA plus two FSELs per cell.
__device__ __forceinline__ void row_fma(double (&a)[4], double gx, double y0, double y1, double y2, double y3, uint32_t m) {
  asm volatile("{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b32 t;\n\t"
      "and.b32 t, %9, 1; setp.eq.u32 p0, t, 0;\n\t"
      "and.b32 t, %9, 2; setp.eq.u32 p1, t, 0;\n\t"
      "and.b32 t, %9, 4; setp.eq.u32 p2, t, 0;\n\t"
      "and.b32 t, %9, 8; setp.eq.u32 p3, t, 0;\n\t"
      "@p0 bra L0;\n\t fma.rn.f64 %0, %4, %5, %0;\n\tL0:\n\t"
      "@p1 bra L1;\n\t fma.rn.f64 %1, %4, %6, %1;\n\tL1:\n\t"
      "@p2 bra L2;\n\t fma.rn.f64 %2, %4, %7, %2;\n\tL2:\n\t"
      "@p3 bra L3;\n\t fma.rn.f64 %3, %4, %8, %3;\n\tL3:\n\t}"
      : "+d"(a[0]), "+d"(a[1]), "+d"(a[2]), "+d"(a[3])
      : "d"(gx), "d"(y0), "d"(y1), "d"(y2), "d"(y3), "r"(m));
}
// position of column c (0..31 of a stripe) in yg: lane lc owns columns 4*lc..4*lc+3 and reads them with two
// 128-bit loads; this order makes both conflict-free (8 lanes x 16 B contiguous each)
__device__ __forceinline__ int yg_pos(int c) { return ((c >> 1) & 1) * 16 + (c >> 2) * 2 + (c & 1); }

// PERSIST: one resident CTA per slot of the GPU walks the tile list of bbox_kernel (shared counter) instead of one
// CTA per possible tile of every event: most possible tiles lie outside their event's bounding rectangle, and those
// CTAs cost a launch slot with 74 KB of shared memory each just to find that out.
template <bool PERSIST>
__global__ void __launch_bounds__(DEP_THREADS, DEP_MINCTA) deposit_kernel(DevCfg c, Store st, KindList kl, int nev, int nbands) {
  extern __shared__ __align__(16) unsigned char dep_smem[];
  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  bool first_tile = true; int next_item = 0;
  __shared__ int s_ccnt[DEP_NCLS];
  if (PERSIST) { if (tid < DEP_NCLS) s_ccnt[tid] = dep_work(st).ctr[1 + tid]; }      // final: bbox_kernel has completed
 for (;;) {
  int e, band, sgroup, kq;
  if (PERSIST) {
    const DepWork wk = dep_work(st);
    if (first_tile) { if (tid == 0) s_item = atomicAdd(wk.ctr, 1); first_tile = false; }
    else if (tid == 0) s_item = next_item;      // fetched while the previous tile was being processed
    __syncthreads();
    int it = s_item, cls = 0;
    while (cls < DEP_NCLS && it >= s_ccnt[cls]) { it -= s_ccnt[cls]; cls++; }     // class lists are walked in order
    if (cls == DEP_NCLS) return;
    if (tid == 0) next_item = atomicAdd(wk.ctr, 1);
    const int2 w = wk.items[(size_t)cls * wk.cap + it];
    e = w.x; kq = w.y >> 16; sgroup = (w.y >> 8) & 0xff; band = w.y & 0xff;
  } else {
    e = blockIdx.x + st.e0; band = blockIdx.y % nbands; sgroup = blockIdx.y / nbands; kq = blockIdx.z;
    if (st.redo && !st.redo[e]) return;
  }
  const int kind = kl.kind[kq], tile_slot = sgroup * nbands + band;
  const int* hi = st.hdr_i + (size_t)e * HDR_I;
  // bands and column groups are laid out from the corner of the event's own bounding rectangle (bbox_kernel),
  // so a CTA is either inside the populated region or exits at once; spectator grids have a rectangle of their own
  const bool spec = (kind == GK_SPEC_A || kind == GK_SPEC_B);
  const int r_org = hi[spec ? H_SRLO : H_RLO], r_end = hi[spec ? H_SRHI : H_RHI];
  const int c_org = hi[spec ? H_SCLO : H_CLO], c_end = hi[spec ? H_SCHI : H_CHI];
  const int r0 = r_org + band * DEP_BAND, c0 = c_org + sgroup * DEP_COLS;
  if (!PERSIST && (r0 >= r_end || c0 >= c_end)) return;
  const int wr = warp / DEP_NSTR, ws = warp % DEP_NSTR, lr = lane >> 3, lc = lane & 7;
  const int rw0 = r0 + wr * DEP_WROWS, sc0 = c0 + ws * 32;
  DepTab* tab = reinterpret_cast<DepTab*>(dep_smem);
  SrcRec* srcs = reinterpret_cast<SrcRec*>(tab + 2);           // [2][DEP_CH]
  int* wtot = reinterpret_cast<int*>(srcs + 2 * DEP_CH);       // [32]
  unsigned short* act = reinterpret_cast<unsigned short*>(wtot + 32);
  const SrcRec* recs = st.src_rec + ((size_t)e * kl.n + kq) * st.src_stride;
  const int slot = st.kind_slot[kind];
  double* grid = st.grids + ((size_t)e * st.nkinds + slot) * (size_t)c.Maxx * c.Maxy;

  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;

  const int status = hi[H_STATUS];
  const int nsrc = (status == 0 || status == 4) ? src_count(c, hi, kind) : 0;
  // ---- ordered compaction of the sources whose window meets this CTA (records written by bbox_kernel) ----
  int nact = 0;
  for (int base = 0; base < nsrc; base += DEP_THREADS) {
    const int k = base + tid;
    bool on = false;
    if (k < nsrc) {
      const short4 w = *reinterpret_cast<const short4*>(&recs[k].iL);
      on = (w.x < w.y) && (w.z < w.w) && (w.x < r0 + DEP_BAND) && (w.y > r0) && (w.z < c0 + DEP_COLS) && (w.w > c0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) wtot[warp] = __popc(m);
    __syncthreads();
    int off = nact;
    for (int w2 = 0; w2 < warp; w2++) off += wtot[w2];
    if (on) act[off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
    for (int w2 = 0; w2 < DEP_THREADS / 32; w2++) nact += wtot[w2];
    __syncthreads();
  }
  const int nchunks = (nact + DEP_CH - 1) / DEP_CH;

  // One table item = DEP_XP rows or DEP_YP columns of one source of the chunk (or, last, the source
  // record of the chunk after it).  Items of chunk n+1 are handed out dynamically (shared counter) into the
  // *other* table buffer while chunk n is being consumed: warps whose tile holds few sources spend their slack
  // building tables, and there is one barrier per chunk.
  const int nyq = min(DEP_NYI, (c.wmax + 2 * DEP_YP - 2) / DEP_YP), nitems = DEP_CH * (DEP_NXG + DEP_NXI + nyq);
  auto build_item = [&](int chunk, int id) {
    DepTab& T = tab[chunk & 1];
    const int nch = min(DEP_CH, nact - chunk * DEP_CH);
    if (id < DEP_CH * DEP_NXG) {                                            // ---- rows: xg, DEP_XG rows per item ----
      const int t = id % DEP_CH, part = id / DEP_CH;
      if (t >= nch) return;
      const SrcRec s = srcs[(chunk & 1) * DEP_CH + t];
      const int rb = part * DEP_XG, i0 = r0 + rb;
      const int ka = max(s.iL - i0, 0), kb = min(s.iR - i0, DEP_XG);
      if (ka >= kb) return;
      double g = s.W, q = 1.0, rec = 1.0;
      if (s.flat != 1) {
        const double i2 = s.flat == 2 ? c.q_inv2w2 : c.inv2w2;
        const double d = __dadd_rn(s.x, -xg_of(c, i0 + ka));
        g = s.W * exp(-__dmul_rn(d, d) * i2); q = exp((2.0 * d * c.dx - c.dx * c.dx) * i2); rec = s.flat == 2 ? c.q_recx : c.recx;
      }
      for (int k = ka; k < kb; k++) { T.xg[t][rb + k] = g; g *= q; q *= rec; }
    } else if (id < DEP_CH * (DEP_NXG + DEP_NXI)) {                         // ---- rows: masks, DEP_XP rows per item ----
      const int id1 = id - DEP_CH * DEP_NXG, t = id1 % DEP_CH, part = id1 / DEP_CH;
      if (t >= nch) return;
      const SrcRec s = srcs[(chunk & 1) * DEP_CH + t];
      if (part == 0) T.desc[t] = make_int4(s.iL, s.iR, s.jL, s.jR);
      const int rb = part * DEP_XP, i0 = r0 + rb;
      const int ka = max(s.iL - i0, 0), kb = min(s.iR - i0, DEP_XP);
      const float fyf = (float)(s.y - c.Ymin) * c.inv_dy_f;
      uint32_t w[DEP_NSTR][DEP_XP];
#pragma unroll
      for (int k = 0; k < DEP_XP; k++) {
#pragma unroll
        for (int s2 = 0; s2 < DEP_NSTR; s2++) w[s2][k] = 0u;
        if (k >= ka && k < kb) {
          const double d = __dadd_rn(s.x, -xg_of(c, i0 + k));
          const double d2 = __dmul_rn(d, d);
          const double rem = __dadd_rn(s.thr, -d2);
          if (rem >= 0.0) {
            float hf; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(hf) : "f"((float)rem));
            hf *= c.inv_dy_f;
            const float tl = fyf - hf, th = fyf + hf, rl = rintf(tl), rh = rintf(th);
            int lo = (int)ceilf(tl), hi2 = (int)floorf(th) + 1;
            if (fabsf(tl - rl) < 2e-3f) {                                   // exact predicate at the doubtful cell
              const int cn = (int)rl; const double dy = __dadd_rn(s.y, -yg_of(c, cn));
              lo = (__dadd_rn(d2, __dmul_rn(dy, dy)) <= s.thr) ? cn : cn + 1;
            }
            if (fabsf(th - rh) < 2e-3f) {
              const int cn = (int)rh; const double dy = __dadd_rn(s.y, -yg_of(c, cn));
              hi2 = (__dadd_rn(d2, __dmul_rn(dy, dy)) <= s.thr) ? cn + 1 : cn;
            }
            lo = max(lo, (int)s.jL); hi2 = min(hi2, (int)s.jR);
            const int a = lo - c0, b = hi2 - c0;
            if (a < b) {
#pragma unroll
              for (int s2 = 0; s2 < DEP_NSTR; s2++)
                w[s2][k] = ones_below(__viaddmax_s32(b, -32 * s2, 0)) & ~ones_below(__viaddmax_s32(a, -32 * s2, 0));
            }
          }
        }
      }
#pragma unroll
      for (int s2 = 0; s2 < DEP_NSTR; s2++) *reinterpret_cast<uint4*>(&T.mk[t][s2 * DEP_BAND + rb]) = make_uint4(w[s2][0], w[s2][1], w[s2][2], w[s2][3]);
    } else if (id < nitems) {                                               // ---- columns: yg ----
      const int id2 = id - DEP_CH * (DEP_NXG + DEP_NXI), t = id2 % DEP_CH;
      if (t >= nch) return;
      const SrcRec s = srcs[(chunk & 1) * DEP_CH + t];
      const int part = max(s.jL - c0, 0) / DEP_YP + id2 / DEP_CH;          // a window spans at most nyq parts
      if (part >= DEP_NYI) return;
      const int cb = part * DEP_YP, j0 = c0 + cb;
      const int ka = max(s.jL - j0, 0) & ~1, kb = min(s.jR - j0, DEP_YP);   // pairs of columns (an extra one is masked off)
      if (ka >= kb) return;
      double g = 1.0, q = 1.0, rec = 1.0;
      if (s.flat != 1) {
        const double i2 = s.flat == 2 ? c.q_inv2w2 : c.inv2w2;
        const double d = __dadd_rn(s.y, -yg_of(c, j0 + ka));
        g = exp(-__dmul_rn(d, d) * i2); q = exp((2.0 * d * c.dy - c.dy * c.dy) * i2); rec = s.flat == 2 ? c.q_recy : c.recy;
      }
      for (int k = ka; k < kb; k += 2) {
        const int cc = cb + k;
        const double v0 = g; g *= q; q *= rec;
        const double v1 = g; g *= q; q *= rec;
        *reinterpret_cast<double2*>(&T.yg[t][(cc & ~31) + yg_pos(cc & 31)]) = make_double2(v0, v1);
      }
    } else {                                                                // ---- source records of chunk+1 ----
      const int t = id - nitems, q = chunk + 1;
      if (t < DEP_CH && q * DEP_CH + t < nact) srcs[(q & 1) * DEP_CH + t] = recs[act[q * DEP_CH + t]];
    }
  };

  if (tid < 2) wtot[tid] = 0;                                   // wtot[0..1] double as the item counters from here on
  if (tid < 2 * DEP_CH && tid < nact) srcs[tid] = recs[act[tid]];
  __syncthreads();
  if (nchunks > 0) for (int id = tid; id < nitems; id += DEP_THREADS) build_item(0, id);
  __syncthreads();
  const int sh = 4 * lc;
  for (int n = 0; n < nchunks; n++) {
    const DepTab& T = tab[n & 1];
    const int nch = min(DEP_CH, nact - n * DEP_CH);
    // ---- accumulate: every warp walks the sources of the chunk whose window meets its tile ----
    bool hit = false;
    if (lane < nch) { const int4 d = T.desc[lane]; hit = (d.y > rw0) && (d.x < rw0 + DEP_WROWS) && (d.w > sc0) && (d.z < sc0 + 32); }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    while (m) {
      const int t = __ffs(m) - 1; m &= m - 1;
      const uint4 mw = *reinterpret_cast<const uint4*>(&T.mk[t][ws * DEP_BAND + wr * DEP_WROWS + 4 * lr]);
      const double2 xa = *reinterpret_cast<const double2*>(&T.xg[t][wr * DEP_WROWS + 4 * lr]);
      const double2 xb = *reinterpret_cast<const double2*>(&T.xg[t][wr * DEP_WROWS + 4 * lr + 2]);
      const double2 ya = *reinterpret_cast<const double2*>(&T.yg[t][ws * 32 + 2 * lc]);          // columns 4lc, 4lc+1
      const double2 yb = *reinterpret_cast<const double2*>(&T.yg[t][ws * 32 + 16 + 2 * lc]);     // columns 4lc+2, 4lc+3
      const uint32_t mr[4] = {mw.x >> sh, mw.y >> sh, mw.z >> sh, mw.w >> sh};
      const double gx[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
      for (int a = 0; a < 4; a++) row_fma(acc[a], gx[a], ya.x, ya.y, yb.x, yb.y, mr[a]);
    }
    // ---- then help building the tables of the next chunk (and the source records of the one after it) ----
    if (n + 1 < nchunks) {
#ifdef DEP_STATIC
      for (int id = tid; id < nitems + DEP_CH; id += DEP_THREADS) build_item(n + 1, id);
#else
      int* counter = &wtot[(n + 1) & 1];
      for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, DEP_GRAB);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= nitems + DEP_CH) break;
#pragma unroll 1
        for (int g2 = 0; g2 < DEP_GRAB; g2 += 32) if (base + g2 < nitems + DEP_CH) build_item(n + 1, base + g2 + lane);
      }
#endif
    }
    if (tid == 0) wtot[n & 1] = 0;                         // counter of chunk n+2
    __syncthreads();
  }
  // only the event's bounding rectangle is ever read back (moments / combine walk the rectangle, the getters blank the
  // rest, profile modes start from a zeroed lattice): cells of the tile beyond it are exact zeros and are not stored
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int i = rw0 + 4 * lr + a;
    if (i < r_end) {
      double* g = grid + (size_t)i * c.Maxy;
      const int j = sc0 + 4 * lc;
#pragma unroll
      for (int b = 0; b < 4; b++) if (j + b < c_end) g[j + b] = acc[a][b];
    }
  }
  // The moments kernel needs sum(rho), sum(x rho), sum(y rho) before it can do anything else (centre of mass,
  // MakeDensity.cpp:2273-2282).  The tile is still in registers here: every warp leaves its three partial sums in a
  // fixed slot and the moments kernel adds the slots of the event in a fixed order (deterministic, no atomics).
  if (kind == GK_RHO && st.cm_part) {
    double s0 = 0, sx = 0, sy = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      double cs = 0;
#pragma unroll
      for (int a = 0; a < 4; a++) { cs += acc[a][b]; }
      s0 += cs; sy += yg_of(c, sc0 + 4 * lc + b) * cs;
    }
#pragma unroll
    for (int a = 0; a < 4; a++) sx += xg_of(c, rw0 + 4 * lr + a) * ((acc[a][0] + acc[a][1]) + (acc[a][2] + acc[a][3]));
    s0 = warp_sum(s0); sx = warp_sum(sx); sy = warp_sum(sy);
    if (lane == 0) {
      double* o = st.cm_part + (((size_t)e * st.cm_slots + tile_slot) * (DEP_THREADS / 32) + warp) * 4;
      o[0] = s0; o[1] = sx; o[2] = sy;
    }
  }
  if (!PERSIST) return;
  __syncthreads();          // the next tile reuses the tables and s_item
 }
}

// Bands / column groups start at the event's own first row / column (>= 0), so ceil(Maxx / DEP_BAND) bands and
// ceil(Maxy / DEP_COLS) groups always reach the end of the rectangle
int deposit_cm_slots(const DevCfg& c) { return ((c.Maxy + DEP_COLS - 1) / DEP_COLS) * ((c.Maxx + DEP_BAND - 1) / DEP_BAND); }

size_t deposit_smem_bytes(const DevCfg& c, int nsrc_max) {
  size_t b = 2 * sizeof(DepTab) + 2 * DEP_CH * sizeof(SrcRec) + 32 * sizeof(int);
  b += (size_t)(nsrc_max + 8) * sizeof(unsigned short);
  return (b + 15) & ~(size_t)15;
}

size_t deposit_work_bytes(const DevCfg& c, int batch, int nk) { return (size_t)DEP_NCLS * ((size_t)batch * nk * deposit_cm_slots(c)) * sizeof(int2) + (1 + DEP_NCLS) * sizeof(int) + 16; }

cudaError_t launch_deposit(const DevCfg& c, const Store& st, const int* kinds, int nk, int nev, cudaStream_t s) {
  KindList kl; kl.n = nk; for (int i = 0; i < nk; i++) kl.kind[i] = kinds[i];
  static const int persist = getenv("SMC_DEP_PERSIST") ? atoi(getenv("SMC_DEP_PERSIST")) : SMC_DEP_PERSIST_DEFAULT;
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
  const int ngroups = (c.Maxy + DEP_COLS - 1) / DEP_COLS;
  const int nbands = (c.Maxx + DEP_BAND - 1) / DEP_BAND;
  const bool use_list = persist && nbands < 256 && ngroups < 256;
  if (use_list) cudaMemsetAsync(reinterpret_cast<char*>(st.src_rec) + st.work_off + (size_t)DEP_NCLS * st.work_cap * sizeof(int2), 0, (1 + DEP_NCLS) * sizeof(int), s);
  bbox_kernel<<<nev, 128, 0, s>>>(c, st, kl, nev, use_list ? 1 : 0);
  const size_t smem = deposit_smem_bytes(c, (c.shape_of_entropy == 3 ? 6 : 2) * c.Amax + c.ncoll_cap);
  if (use_list) {
    cudaFuncSetAttribute(deposit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long want = (long)nev * nbands * ngroups * nk;
    // SMC_DEP_CTAS < DEP_MINCTA leaves room on every SM for CTAs of the other pipeline slots' kernels (sampler, moments)
    static const int per_sm = getenv("SMC_DEP_CTAS") ? std::max(1, std::min(atoi(getenv("SMC_DEP_CTAS")), DEP_MINCTA)) : DEP_MINCTA;
    const int grid = (int)std::min<long>(want, (long)n_sm * per_sm * persist);
    deposit_kernel<true><<<grid, DEP_THREADS, smem, s>>>(c, st, kl, nev, nbands);
  } else {
    cudaFuncSetAttribute(deposit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 g(nev, nbands * ngroups, nk);
    deposit_kernel<false><<<g, DEP_THREADS, smem, s>>>(c, st, kl, nev, nbands);
  }
  return cudaGetLastError();
}

#define COMB_BLOCKS 32      // blocks per event of combine_kernel (= centre-of-mass partial-sum slots it leaves)
#define COMB_THREADS 256
#define COMB_ILP 4
// ---- pointwise combinations -------------------------------------------------------------------
// which_mc_model 7: rho = sqrt(rhoA*rhoB) (MCnucl.cpp:797-803); which_mc_model 1: 6-point table lookup
// (MCnucl.cpp:654-687, arsenal.cpp:33-54)
__global__ void combine_kernel(DevCfg c, Store st, int nev) {
  const int e = blockIdx.y + st.e0;
  if (st.redo && !st.redo[e]) return;
  const size_t G = (size_t)c.Maxx * c.Maxy;
  int* hi = st.hdr_i + (size_t)e * HDR_I;
  double* base = st.grids + (size_t)e * st.nkinds * G;
  double* rho = base + (size_t)st.kind_slot[GK_RHO] * G;
  // Only the event's bounding rectangle (bbox_kernel: union of the source windows of every deposited kind) can
  // hold a non-zero input: outside it rho is 0 in the reference too (sqrt(0 * 0); table[0][j] = table[i][0] = 0,
  // MCnucl.cpp:937-944), and nothing downstream reads it (moments walk the same rectangle, getters zero the rest)
  const int ilo = hi[H_RLO], ihi = hi[H_RHI], jlo = hi[H_CLO], jhi = hi[H_CHI];
  const double* ga = base + (size_t)st.kind_slot[c.which_mc_model == 7 ? GK_RHOA : GK_TA1] * G;
  const double* gb = base + (size_t)st.kind_slot[c.which_mc_model == 7 ? GK_RHOB : GK_TA2] * G;
  bool overflow = false;
  const double inv_dT = 1.0 / c.kln_dT;
  double s0 = 0, sx = 0, sy = 0;          // centre-of-mass sums of this block's cells (MakeDensity.cpp:2273-2282)
  // the rectangle as a linear list of cells, COMB_ILP cells per thread in flight: the kernel is bound by the latency
  // of the dependent loads (thicknesses -> table entries), not by their volume
  const int wj = max(jhi - jlo, 0), ncell = max(ihi - ilo, 0) * wj, stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + threadIdx.x; base < ncell; base += COMB_ILP * stride) {
    double a[COMB_ILP], b[COMB_ILP]; int ci[COMB_ILP], cj[COMB_ILP];
#pragma unroll
    for (int u = 0; u < COMB_ILP; u++) {
      const int idx = base + u * stride;
      a[u] = 0.0; b[u] = 0.0; ci[u] = -1; cj[u] = 0;
      if (idx < ncell) {
        const int ir = idx / wj; ci[u] = ilo + ir; cj[u] = jlo + (idx - ir * wj);
        const size_t k = (size_t)ci[u] * c.Maxy + cj[u];
        a[u] = ga[k]; b[u] = gb[k];
      }
    }
#pragma unroll
    for (int u = 0; u < COMB_ILP; u++) {
      if (ci[u] < 0) continue;
      const size_t k = (size_t)ci[u] * c.Maxy + cj[u];
      double r;
      if (a[u] == 0.0 || b[u] == 0.0) {
        r = 0.0;           // sqrt(0) and the lookup at the table's zero row / column (table[0][j] = table[i][0] = 0) are exactly 0
      } else if (c.which_mc_model == 7) {
        r = sqrt(a[u] * b[u]);
      } else {
        // TA / dT (MCnucl.cpp:660-661) with the reciprocal of the constant divisor hoisted: q = a * (1/dT), then one
        // exact-remainder correction step, q + fma(-q, dT, a) * (1/dT) -- Markstein's final division step, which returns
        // the correctly rounded quotient when 1/dT is correctly rounded (3 instructions instead of a division each)
        const double q1 = __dmul_rn(a[u], inv_dT), q2 = __dmul_rn(b[u], inv_dT);
        const double di = __fma_rn(__fma_rn(-q1, c.kln_dT, a[u]), inv_dT, q1), dj = __fma_rn(__fma_rn(-q2, c.kln_dT, b[u]), inv_dT, q2);
        if (di < 0 || di >= c.kln_tmax - 2 || dj < 0 || dj >= c.kln_tmax - 2) { overflow = true; rho[k] = 0.0; continue; }
        const int ii = (int)di, jj = (int)dj;               // floor of a non-negative number
        const double x = di - ii, y = dj - jj;
        const double* T = st.kln_table; const int tm = c.kln_tmax;
        const double v00 = T[ii * tm + jj], v01 = T[ii * tm + jj + 1], v02 = T[ii * tm + jj + 2];
        const double v10 = T[(ii + 1) * tm + jj], v11 = T[(ii + 1) * tm + jj + 1], v20 = T[(ii + 2) * tm + jj];
        const double axx = 1.0 / 2.0 * (v00 - 2 * v10 + v20), axy = v00 - v01 - v10 + v11, ayy = 1.0 / 2.0 * (v00 - 2 * v01 + v02);
        const double bx = 1.0 / 2.0 * (-3.0 * v00 + 4 * v10 - v20), by = 1.0 / 2.0 * (-3.0 * v00 + 4 * v01 - v02);
        r = axx * x * x + axy * x * y + ayy * y * y + bx * x + by * y + v00;
      }
      rho[k] = r;
      s0 += r; sx += xg_of(c, ci[u]) * r; sy += yg_of(c, cj[u]) * r;
    }
  }
  if (overflow) hi[H_STATUS] = 4;
  // per-block partial sums in a fixed slot; the moments kernel adds the COMB_BLOCKS slots in order (deterministic)
  if (st.cm_part) {
    __shared__ double red[3][COMB_THREADS / 32];
    s0 = warp_sum(s0); sx = warp_sum(sx); sy = warp_sum(sy);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = sx; red[2][threadIdx.x >> 5] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t0 = 0, tx = 0, ty = 0;
      for (int w = 0; w < COMB_THREADS / 32; w++) { t0 += red[0][w]; tx += red[1][w]; ty += red[2][w]; }
      double* o = st.cm_part + ((size_t)e * st.cm_slots * (DEP_THREADS / 32) + blockIdx.x) * 4;
      o[0] = t0; o[1] = tx; o[2] = ty;
    }
  }
}
cudaError_t launch_combine(const DevCfg& c, const Store& st, int nev, cudaStream_t s) {
  dim3 g(COMB_BLOCKS, nev);
  combine_kernel<<<g, COMB_THREADS, 0, s>>>(c, st, nev);
  return cudaGetLastError();
}

// ---- NBD multiplicity fluctuations (cc_fluctuation_model 1, 2) ------------------------------------------
// MCnucl::fluctuateCurrentDensity (MCnucl.cpp:868-905) replaces rho dx dy of every cell by a draw of NBD::rand(p, r)
// (NBD.cpp:31-90).  That sampler works through RandomVariable's step-function envelope (RandomVariable.cpp:190-286):
// M = s + 6 intervals of width std = sqrt(p r)/(1-p) from mode - s std on, interval m of height h_m = the larger pmf
// of its two ends; a point is drawn uniformly from interval m with probability ~ std h_m and kept with probability
// min(1, pmf/h_m); the integer part is returned.  Its law is therefore
//     P(k) ~ sum_m |[e_m, e_m+1) n [k, k+1)| min(h_m, pmf(k)),     k = floor(e_0) ... floor(e_M)
// (a truncated NBD: cells with 6 std < 1 always give 0).  The reference walks the lattice with one sequential drand48;
// here every cell inverts that law at its own Philox uniform, which makes the draw addressable and exactly
// reproducible by the oracle (smc_o_nbd_quantile).
__device__ double nbd_pmf(double p, double r, double k_in) {
  if (k_in < 0) return 0;
  const int k = (int)floor(k_in);
  return exp(lgamma(k + r) - lgamma(k + 1.0) - lgamma(r)) * pow(1 - p, r) * pow(p, (double)k);
}
// pmf(k+1) = pmf(k) p (k + r) / (k + 1): one multiply chain serves all thirteen envelope edges and both passes over the
// support, so a cell costs one pow (or three lgamma when the support starts far from 0) instead of ~20 pmf evaluations
// (the uniform is a callable: most cells of an event's rim return 0 before they would use it, and a Philox call costs ~100 instructions)
template <class U>
__device__ long nbd_quantile(double p, double r, U uniform) {
  const double ZERO = 1e-15;
  if (p < ZERO || p + ZERO > 1.0) return 0;
  const double mode = (r <= 1) ? 1e-30 : p * (r - 1) / (1 - p), sd = sqrt(p * r) / (1 - p);
  int sl;
  for (sl = 6; sl > 0; sl--) if (mode - sd * sl >= 0) break;
  const int M = sl + 6;
  double edge[13], hgt[12];
  double LB = mode - sl * sd, RB = LB + sd;
  if (sl == 0 && LB + 6.5 * sd <= 1.0) return 0;          // the whole envelope lies inside the cell k = 0 (with slack for the summed edge)
  edge[0] = LB;
  for (int m = 0; m < M; m++) { edge[m + 1] = RB; RB += sd; }
  if (edge[M] <= 1.0) return 0;
  const long ka = (long)floor(edge[0]), kb = (long)floor(edge[M]);
  double pka;                                              // pmf(ka)
  if (ka <= 64) { pka = pow(1 - p, r); for (long k = 0; k < ka; k++) pka *= p * ((double)k + r) / (double)(k + 1); }
  else pka = nbd_pmf(p, r, (double)ka);
  {                                                        // heights: the larger pmf of the two ends of every interval
    long kc = ka; double pc = pka, pl = pka;
    for (int m = 0; m < M; m++) {
      const long km = (long)floor(edge[m + 1]);
      while (kc < km) { pc *= p * ((double)kc + r) / (double)(kc + 1); kc++; }
      hgt[m] = pl > pc ? pl : pc; pl = pc;
    }
  }
  double tot = 0, u = 0;
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) u = uniform();
    double acc = 0, pk = pka; const double target = u * tot;
    for (long k = ka; k <= kb; k++) {
      double w = 0;
      for (int m = 0; m < M; m++) {
        const double lo = fmax(edge[m], (double)k), hi = fmin(edge[m + 1], (double)(k + 1));
        if (hi > lo) w += (hi - lo) * fmin(hgt[m], pk);
      }
      acc += w;
      if (pass == 1 && acc > target) return k;
      pk *= p * ((double)k + r) / (double)(k + 1);
    }
    tot = acc;
  }
  return kb;
}
__global__ void fluctuate_kernel(DevCfg c, Store st, int nev) {
  const int e = blockIdx.y + st.e0;
  if (st.redo && !st.redo[e]) return;
  const int* hi = st.hdr_i + (size_t)e * HDR_I;
  if (!(hi[H_STATUS] == 0 || hi[H_STATUS] == 4)) return;
  const size_t G = (size_t)c.Maxx * c.Maxy;
  double* base = st.grids + (size_t)e * st.nkinds * G;
  double* rho = base + (size_t)st.kind_slot[GK_RHO] * G;
  const double* ta = c.cc_fluct == 2 ? base + (size_t)st.kind_slot[GK_TA1] * G : nullptr;
  const double* tb = c.cc_fluct == 2 ? base + (size_t)st.kind_slot[GK_TA2] * G : nullptr;
  // outside the bounding rectangle rho = 0 and the draw is 0 (p = 0 / nb < 1e-10), MCnucl.cpp:877-894
  const int ilo = hi[H_RLO], ihi = hi[H_RHI], jlo = hi[H_CLO], jhi = hi[H_CHI];
  const int wj = max(jhi - jlo, 0), ncell = max(ihi - ilo, 0) * wj;
  const double cell = c.dx * c.dy, kpp = 1.0 / SMC_PI * c.dx * c.dy * 1.0 * (0.25 * 0.25 / 0.197327053 / 0.197327053);
  const uint64_t ev = st.event_id[e];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ncell; idx += gridDim.x * blockDim.x) {
    const int ir = idx / wj, i = ilo + ir, j = jlo + (idx - ir * wj);
    const size_t q = (size_t)i * c.Maxy + j;
    const double nb = rho[q] * cell;
    auto draw = [&]() { return smc_uniform_cell(c.seed_lo, c.seed_hi, ev, (uint32_t)st.nbd_pass, (uint32_t)q); };
    double n;
    if (c.cc_fluct == 1) {
      n = (nb == 0.0) ? 0.0 : (double)nbd_quantile(nb / (nb + c.cc_k), c.cc_k, draw);
    } else {
      const double k = kpp * fmin(ta[q], tb[q]) * c.siginNN / 10;
      if (nb < 1e-10) n = nb;
      else n = (double)nbd_quantile(nb / (nb + k), k, draw);
    }
    rho[q] = n / cell;
  }
}
cudaError_t launch_fluctuate(const DevCfg& c, const Store& st, int nev, cudaStream_t s) {
  dim3 g(COMB_BLOCKS, nev);
  fluctuate_kernel<<<g, COMB_THREADS, 0, s>>>(c, st, nev);
  return cudaGetLastError();
}

// ---- K4: moments --------------------------------------------------------------------------------
#ifndef MOM_THREADS
#define MOM_THREADS 128    // 3 CTAs/SM leave 170 registers per thread: the 57 accumulators stay in registers
#endif
#ifndef MOM_LD
#define MOM_LD 8      // density loads a lane keeps in flight
#endif
#ifndef MOM_MINCTA
#define MOM_MINCTA 3
#endif
#ifndef MOM_UNROLL
#define MOM_UNROLL 1    // cells of a lane's column step in flight at once
#endif
#define SMC_PRAGMA_(x) _Pragma(#x)
#define SMC_UNROLL(n) SMC_PRAGMA_(unroll n)
__device__ __forceinline__ double block_sum(double v, double* red, int tid) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < MOM_THREADS / 32; w++) s += red[w];
  return s;
}
__device__ __forceinline__ double block_min(double v, double* red, int tid) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = red[0];
  for (int w = 1; w < MOM_THREADS / 32; w++) s = fmin(s, red[w]);
  return s;
}

__global__ void __launch_bounds__(MOM_THREADS, MOM_MINCTA) moments_kernel(DevCfg c, Store st, int nev) {
  extern __shared__ double smem_d[];
  const int e = blockIdx.x + st.e0, tid = threadIdx.x;
  if (st.redo && !st.redo[e]) return;
  const int* hi = st.hdr_i + (size_t)e * HDR_I;
  double* out = st.mom_out + (size_t)e * MOM_OUT;
  const int status = hi[H_STATUS];
  if (!(status == 0 || status == 4)) { if (tid < MOM_OUT) out[tid] = 0.0; return; }
  const int Maxx = c.Maxx, Maxy = c.Maxy, MW = (Maxy + 31) / 32;
  double* red = smem_d;                       // [MOM_THREADS/32 + 1][64]
  double* stage = red + (MOM_THREADS / 32 + 1) * 64;                   // [2][MOM_LD][MOM_THREADS] density staging
  uint32_t* mask = (uint32_t*)(stage + 2 * MOM_LD * MOM_THREADS);      // [Maxx][MW]
  const size_t G = (size_t)Maxx * Maxy;
  const double* rho = st.grids + ((size_t)e * st.nkinds + st.kind_slot[GK_RHO]) * G;
  const int Amax = c.Amax, np = hi[H_NP1] + hi[H_NP2];
  const double* nuc = st.nuc + (size_t)e * 2 * Amax * NROW;
  const int* pidx = st.part_idx + (size_t)e * 2 * Amax;
  for (int k = tid; k < Maxx * MW; k += MOM_THREADS) mask[k] = 0u;
  // hot-spot region: union of participant AABBs and the zero-size collision boxes at the origin
  // (MCnucl.cpp:1303-1323, CollisionPair.h:15-17 -- quirk Q13)
  double xl = 1e300, xr = -1e300, yl = 1e300, yr = -1e300;
  for (int k = tid; k < np; k += MOM_THREADS) {
    const int id = pidx[k]; const double* r = nuc + ((size_t)(id >> 16) * Amax + (id & 0xffff)) * NROW;
    xl = fmin(xl, r[NXL]); xr = fmax(xr, r[NXR]); yl = fmin(yl, r[NYL]); yr = fmax(yr, r[NYR]);
  }
  if (hi[H_NCOLL] > 0) { xl = fmin(xl, 0.0); xr = fmax(xr, 0.0); yl = fmin(yl, 0.0); yr = fmax(yr, 0.0); }
  const double rXL = block_min(xl, red, tid), rXR = -block_min(-xr, red, tid);
  const double rYL = block_min(yl, red, tid), rYR = -block_min(-yr, red, tid);
  const int nX = (int)__ddiv_rn(__dadd_rn(rXR, -rXL), c.dx), nY = (int)__ddiv_rn(__dadd_rn(rYR, -rYL), c.dy);   // MakeDensity.cpp:2343-2344
  const int i0 = cell_of(rXL, c.Xmin, c.dx), j0 = cell_of(rYL, c.Ymin, c.dy);                                    // :2399-2400
  __syncthreads();
  // boolean mask (MakeDensity.cpp:2354-2358), stored in grid coordinates; one (box,row) per thread step
  for (int k = tid; k < np; k += MOM_THREADS) {
    const int id = pidx[k]; const double* r = nuc + ((size_t)(id >> 16) * Amax + (id & 0xffff)) * NROW;
    int x0 = cell_of(r[NXL], rXL, c.dx), x1 = cell_of(r[NXR], rXL, c.dx);
    int y0 = cell_of(r[NYL], rYL, c.dy), y1 = (int)__ddiv_rn(__dadd_rn(r[NYR], -rYL), c.dx);                    // sic: dx (quirk Q5)
    x0 = max(x0, 0); x1 = min(x1, nX); y0 = max(y0, 0); y1 = min(y1, nY);
    int ja = max(y0 + j0, 0), jb = min(y1 + j0, Maxy);
    if (ja >= jb) continue;
    for (int xi = x0; xi < x1; xi++) {
      const int i = xi + i0;
      if (i < 0 || i >= Maxx) continue;
      for (int wd = ja >> 5; wd <= (jb - 1) >> 5; wd++) {
        const int lo = max(ja - wd * 32, 0), hi2 = min(jb - wd * 32, 32);
        const uint32_t bits = (hi2 - lo >= 32) ? 0xffffffffu : (((1u << (hi2 - lo)) - 1u) << lo);
        atomicOr(&mask[i * MW + wd], bits);
      }
    }
  }
  // bounding rectangle of non-zero density = union of the source windows (bbox_kernel); the derived densities
  // (sqrt(rhoA rhoB), the MC-KLN table lookup) vanish outside it as well (combine_kernel)
  const int ilo = hi[H_RLO], ihi = hi[H_RHI], jlo = hi[H_CLO], jhi = hi[H_CHI];
  const int lane = tid & 31, warp = tid >> 5;
  // warp w walks rows ilo + w, ilo + w + 8, ...; lanes walk the columns of the row (no index divisions, the
  // row coordinate is hoisted)
  // ---- pass 1: centre of mass (MakeDensity.cpp:2273-2282) ----
  double total, xc, yc;
  if (c.which_mc_model == 5 && st.cm_part) {
    // MC-Glauber: the deposit CTAs left sum(rho), sum(x rho), sum(y rho) of their tiles; add them in slot order
    const int nbands = (Maxx + DEP_BAND - 1) / DEP_BAND;
    const int nb = (max(ihi - ilo, 0) + DEP_BAND - 1) / DEP_BAND, ng = (max(jhi - jlo, 0) + DEP_COLS - 1) / DEP_COLS;
    const int NWD = DEP_THREADS / 32, nslot = nb * ng * NWD;
    const double* part = st.cm_part + (size_t)e * st.cm_slots * NWD * 4;
    double s0 = 0, sx = 0, sy = 0;
    for (int q = tid; q < nslot; q += MOM_THREADS) {
      const int w2 = q % NWD, b2 = (q / NWD) % nb, g2 = q / (NWD * nb);
      const double* o = part + ((size_t)(g2 * nbands + b2) * NWD + w2) * 4;
      s0 += o[0]; sx += o[1]; sy += o[2];
    }
    const double t0 = block_sum(s0, red, tid);
    total = t0 * c.finalFactor; xc = block_sum(sx, red, tid) / t0; yc = block_sum(sy, red, tid) / t0;
  } else if (st.cm_part && st.cm_slots * (DEP_THREADS / 32) >= COMB_BLOCKS) {
    // derived densities (sqrt scaling, MC-KLN): combine_kernel left the sums of its row blocks
    const double* part = st.cm_part + (size_t)e * st.cm_slots * (DEP_THREADS / 32) * 4;
    double s0 = 0, sx = 0, sy = 0;
    if (tid < COMB_BLOCKS) { s0 = part[tid * 4]; sx = part[tid * 4 + 1]; sy = part[tid * 4 + 2]; }
    const double t0 = block_sum(s0, red, tid);
    total = t0 * c.finalFactor; xc = block_sum(sx, red, tid) / t0; yc = block_sum(sy, red, tid) / t0;
  } else {
    double s0 = 0, sx = 0, sy = 0;
    for (int i = ilo + warp; i < ihi; i += MOM_THREADS / 32) {
      const double xg = xg_of(c, i); const double* row = rho + (size_t)i * Maxy;
      for (int jb = jlo + lane; jb < jhi; jb += 32 * MOM_LD) {
        double dv[MOM_LD];
#pragma unroll
        for (int u = 0; u < MOM_LD; u++) { const int j = jb + 32 * u; dv[u] = (j < jhi) ? row[j] : 0.0; }   // MOM_LD loads in flight
#pragma unroll
        for (int u = 0; u < MOM_LD; u++) {
          const int j = jb + 32 * u;
          const double d = dv[u] * c.finalFactor;
          s0 += d; sx += xg * d; sy += yg_of(c, j) * d;
        }
      }
    }
    total = block_sum(s0, red, tid);
    xc = block_sum(sx, red, tid) / total; yc = block_sum(sy, red, tid) / total;
  }
  // ---- pass 2: <r^n>, eps_n, eps'_n (MakeDensity.cpp:2285-2298, 2389-2430) ----
  double rn[10], mr[10], mi[10], pr[10], pi[10], npw[10], nrm = 0, nnz = 0;
#pragma unroll
  for (int n = 0; n < 10; n++) { rn[n] = 0; mr[n] = 0; mi[n] = 0; pr[n] = 0; pi[n] = 0; npw[n] = 0; }
  // The density of chunk t+1 (MOM_LD column steps of one row) streams into shared memory with cp.async while
  // chunk t is being processed: the loads of a lane never wait in registers
  const int nj = max(jhi - jlo, 0), cpr = (nj + 32 * MOM_LD - 1) / (32 * MOM_LD);
  const int nrows_w = (ihi - ilo - warp + MOM_THREADS / 32 - 1) / (MOM_THREADS / 32), T = max(nrows_w, 0) * cpr;
  auto issue = [&](int t) {
    const int i = ilo + warp + (t / cpr) * (MOM_THREADS / 32), jb = jlo + lane + (t % cpr) * 32 * MOM_LD;
    const double* row = rho + (size_t)i * Maxy; double* dst = stage + (size_t)(t & 1) * MOM_LD * MOM_THREADS + tid;
#pragma unroll
    for (int u = 0; u < MOM_LD; u++) {
      const int j = jb + 32 * u;
      if (j < jhi) __pipeline_memcpy_async(dst + u * MOM_THREADS, row + j, sizeof(double)); else dst[u * MOM_THREADS] = 0.0;
    }
    __pipeline_commit();
  };
  if (T > 0) issue(0);
  for (int t = 0; t < T; t++) {
    if (t + 1 < T) { issue(t + 1); __pipeline_wait_prior(1); } else __pipeline_wait_prior(0);
    const int i = ilo + warp + (t / cpr) * (MOM_THREADS / 32), jb = jlo + lane + (t % cpr) * 32 * MOM_LD;
    const double x = xg_of(c, i) - xc, xx = x * x;
    const uint32_t* mrow = mask + i * MW;
    const double* src = stage + (size_t)(t & 1) * MOM_LD * MOM_THREADS + tid;
    {
      SMC_UNROLL(MOM_UNROLL)
      for (int u = 0; u < MOM_LD; u++) {
      const int j = jb + 32 * u;
      const double d = src[u * MOM_THREADS] * c.finalFactor;
      if (d == 0.0) continue;
      nnz += 1.0;
      const double y = yg_of(c, j) - yc;
      const double r2 = xx + y * y;
      double ux = 1.0, uy = 0.0, r = 0.0;                          // atan2(0,0) = 0
      if (r2 > 0.0) { const double ri = rsqrt(r2); r = r2 * ri; ux = x * ri; uy = y * ri; }
      // hot-spot mask as a 0/1 weight: the masked sums take w * (...) instead of five predicated updates per order
      const double w = (double)((mrow[j >> 5] >> (j & 31)) & 1u);
      const double d2 = d * r2, d2m = d2 * w;
      rn[0] += d;
      nrm += d2m;
      // A_n + i B_n = rho w e^{i n phi} by the three-term recurrence c_n = 2 cos(phi) c_(n-1) - c_(n-2) (linear, so the cell's
      // masked density rides along): nine FP64 operations per order -- 2 recurrence, r^n, <r^n>, 2 for eps_n, 3 for eps'_n
      const double dw = d * w, tux = 2.0 * ux;
      double An = dw, Bn = 0.0, Ap = dw * ux, Bp = -dw * uy, rpow = 1.0;      // (Ap, Bp): order n - 2; n = 1 starts from -phi
#pragma unroll
      for (int n = 1; n < 10; n++) {
        const double A2 = fma(tux, An, -Ap), B2 = fma(tux, Bn, -Bp);
        Ap = An; Bp = Bn; An = A2; Bn = B2;
        rpow *= r;
        rn[n] = fma(d, rpow, rn[n]);
        mr[n] = fma(r2, An, mr[n]); mi[n] = fma(r2, Bn, mi[n]);
        const double wt = (n == 1) ? r2 * rpow : rpow;                        // eps'_1 is weighted with r^3
        pr[n] = fma(wt, An, pr[n]); pi[n] = fma(wt, Bn, pi[n]); npw[n] = fma(dw, wt, npw[n]);
      }
      }
    }
  }
  // ---- one reduction for all 57 sums: a transposing butterfly leaves every lane with two fully warp-reduced
  // values (5 stages, 62 shuffles instead of 57 x 5), then one pass over the per-warp partials in shared memory
  double v[64];
#pragma unroll
  for (int n = 0; n < 10; n++) v[n] = rn[n];
  v[10] = nrm; v[11] = nnz;
#pragma unroll
  for (int n = 1; n < 10; n++) { v[12 + (n - 1) * 5] = mr[n]; v[13 + (n - 1) * 5] = mi[n]; v[14 + (n - 1) * 5] = pr[n]; v[15 + (n - 1) * 5] = pi[n]; v[16 + (n - 1) * 5] = npw[n]; }
#pragma unroll
  for (int k = 57; k < 64; k++) v[k] = 0.0;
#pragma unroll
  for (int off = 16, cnt = 32; off > 0; off >>= 1, cnt >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; i++) {
      const double keep = up ? v[2 * i + 1] : v[2 * i], send = up ? v[2 * i] : v[2 * i + 1];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  const int bl = (int)(__brev((unsigned)lane) >> 27);          // lane holds sums number bl and 32 + bl
  __syncthreads();
  red[warp * 64 + bl] = v[0]; red[warp * 64 + 32 + bl] = v[1];
  __syncthreads();
  double* fin = red + (MOM_THREADS / 32) * 64;
  if (tid < 64) { double t = 0; for (int w = 0; w < MOM_THREADS / 32; w++) t += red[w * 64 + tid]; fin[tid] = t; }
  __syncthreads();
  if (tid == 0) {
    const double eps = 1e-15, dn = fin[0], nrmS = fin[10], nnzS = fin[11];
    out[45] = dn / dn; out[46] = total * c.dx * c.dy; out[47] = xc; out[48] = yc; out[49] = (total / c.finalFactor) * c.dx * c.dy; out[50] = nnzS;
    for (int n = 1; n < 10; n++) {
      const double a = fin[12 + (n - 1) * 5], b = fin[13 + (n - 1) * 5], cc = fin[14 + (n - 1) * 5], dd = fin[15 + (n - 1) * 5], ee = fin[16 + (n - 1) * 5], ff = fin[n];
      const bool on = (n >= c.ecc_from && n <= c.ecc_to);        // orders outside [from,to] stay 0 (MakeDensity.cpp:2389,2475)
      double* o = out + (n - 1) * 5;
      o[0] = on ? -a / (nrmS + eps) : 0.0; o[1] = on ? -b / (nrmS + eps) : 0.0;
      o[2] = on ? -cc / (ee + eps) : 0.0; o[3] = on ? -dd / (ee + eps) : 0.0;
      o[4] = ff / dn;
    }
  }
}

cudaError_t launch_moments(const DevCfg& c, const Store& st, int nev, cudaStream_t s) {
  static const size_t pad = getenv("SMC_MOM_PAD") ? (size_t)atoi(getenv("SMC_MOM_PAD")) : 0;      // tuning aid: caps the CTAs per SM so that other kernels' CTAs co-reside
  const size_t smem = ((MOM_THREADS / 32 + 1) * 64 + 2 * MOM_LD * MOM_THREADS) * sizeof(double) + (size_t)c.Maxx * ((c.Maxy + 31) / 32) * sizeof(uint32_t) + pad;
  cudaFuncSetAttribute(moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  moments_kernel<<<nev, MOM_THREADS, smem, s>>>(c, st, nev);
  return cudaGetLastError();
}

}  // namespace smc
