// smc_common.cuh -- shared device-side definitions of the B200 superMC hot path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "smc_philox.h"

#define SMC_PI 3.14159265358979323846

namespace smc {

// nucleon row layout (also the C-ABI layout of smc_event_in.proj/targ)
enum { NX = 0, NY = 1, NZ = 2, NXL = 3, NXR = 4, NYL = 5, NYR = 6, NW = 7, NROW = 8 };
// extra nucleon state for operation 3: stale base box (Particle::baseBox), valence-quark offsets, AABB centre
enum { XBXL = 0, XBXR = 1, XBYL = 2, XBYR = 3, XQ = 4 /* 9 doubles: x y z of the three valence quarks */, XCX = 13, XCY = 14,
       XF = 15 /* 3 doubles: per-quark multiplicity weights (shape_of_entropy 3, Quark::fluctFactor) */, NEXTRA = 20 };
// collision row layout
enum { CX = 0, CY = 1, CW = 2, CADDW = 3, CROW = 4 };

// Everything a kernel needs about the run; passed by value as a kernel parameter.
struct DevCfg {
  int Maxx, Maxy;
  double Xmin, Ymin, dx, dy;
  double w;               // nucleon = entropy gaussian width (reference quirk Q1, MCnucl.cpp:94-96)
  double dsq, siginNN, sigma_gg, alpha;
  double dmax;            // window half-size: 5w (gaussian) or 2 sqrt(dsq) (disk)   MCnucl.cpp:440-444
  double thrA;            // largest dc with sqrt(dc) <= 5w   (Particle::getSmoothTn mask, Particle.cpp:124-126)
  double thrB;            // 25 w^2                            (MCnucl.cpp:647,754)
  double norm;            // 1/(2 pi w^2)
  double inv2w2;          // 1/(2 w^2)
  double areai;           // 10/sigma_in (disk deposits)
  double recx, recy;      // exp(-dx^2/w^2), exp(-dy^2/w^2): second-order ratio of the Gaussian recurrence
  double rclip_flat;      // sqrt(dsq): reach of a disk deposit
  int kln_tmax_param;     // the `tmax` parameter (overflow check of calculateThickness, MCnucl.cpp:413-429)
  double hit_c1, hit_c2;  // sigma_gg/(4 pi w^2), 1/(4 w^2)   GaussianNucleonsCal.cpp:59-67
  double finalFactor;
  int shape_of_nucleons, shape_of_entropy, crit;   // crit: 1 disk, 2 gaussian
  int which_mc_model, sub_model, cc_fluct;
  double cc_k;            // NBD k of cc_fluctuation_model 1
  int A[2];               // mass numbers
  int deformed[2];
  int sampler[2];         // 0 WS, 1 single nucleon, 2 config table (no recentre), 3 config table (NN-corr), 4 deuteron
  double rad[2], dr[2], rmaxCut[2], rwMax[2], beta2[2], beta4[2];
  double bmin, bmax;
  int npmin, npmax;
  double gam_k_part, gam_th_part, gam_k_bin, gam_th_bin;   // MCnucl.cpp:1271-1301
  double quark_width, quark_R;
  // shape_of_entropy 3: wounded nucleons deposit three quark Gaussians of width quark_width (Quark.cpp:14-22)
  double q_inv2w2, q_recx, q_recy, q_norm, q_thr, q_reach;
  int quark_rows;
  int ncfg[2];
  uint32_t seed_lo, seed_hi;
  int Amax;               // row stride of the nucleon arrays
  int ncoll_cap;
  int ecc_from, ecc_to;
  // KLN table lookup
  int kln_tmax; double kln_dT;
  // deposit geometry
  int wmax;               // max window cells per axis (+ slack)
  float inv_dx_f, inv_dy_f;   // single-precision 1/dx, 1/dy (interval estimates of the deposit masks)
};

// one deposit source, expanded once per event by bbox_kernel: position, folded weight, mask threshold, window
struct SrcRec { double x, y, W, thr; short iL, iR, jL, jR; int flat; int pad; };   // 48 bytes; flat: 0 gaussian(w), 1 disk, 2 gaussian(quark_width)

// device-resident event records for one batch
struct Store {
  double* nuc;        // [batch][2][Amax][NROW]
  int* nuc_ncoll;     // [batch][2][Amax]
  int* nuc_first;     // [batch][Amax]   first-hit rank of target nucleons (participant order of the reference)
  double* coll;       // [batch][ncoll_cap][CROW]
  int* coll_ij;       // [batch][ncoll_cap]  (i<<16 | j)
  int* part_idx;      // [batch][2*Amax]   compact participants: side<<16 | i  (proj first)
  int* spec_idx;      // [batch][2*Amax]   compact spectators
  int* hdr_i;         // [batch][HDR_I]
  double* hdr_d;      // [batch][HDR_D]
  const double* quark_table;   // [quark_rows][3]
  const double* cfg_table[2];  // [ncfg][A][3]
  const double* pair_u;        // [batch][A0*A1] or null
  const double* coll_w;        // [batch][ncoll_cap][2] or null
  const uint64_t* event_id;    // [batch]
  int* try_start;              // [batch] first try index to use (dS/dy-cut re-runs)
  const int* redo;             // [batch] or null: only events with redo[e] != 0 are (re)processed
  double* nuc_extra;           // [batch][2][Amax][NEXTRA] or null: state the averaged-profile path needs (quirk Q4)
  double* nuc_extra_tmp;       // same, acceptance order (scratch of the sampler)
  double* cm;                  // [batch][4] xcm, ycm, angle, weight  (GlueDensity::calcCMAngle)
  double* grids;      // [batch][nkinds][G]
  int kind_slot[8];   // grid kind -> slot in grids (or -1)
  int nkinds;
  double* mom_out;    // [batch][MOM_OUT]
  double* kln_table;  // [tmax][tmax]
  SrcRec* src_rec;    // [batch][deposit kinds][src_stride]
  int src_stride;
  size_t work_off;    // byte offset, from src_rec, of the deposit tile list (int2 items[work_cap], then two counters)
  int work_cap;
  double* cm_part;    // [batch][cm_slots][4] per deposit CTA: sum rho, sum x rho, sum y rho of its tile (MC-Glauber rho only)
  int cm_slots;
  int nbd_pass;       // how often this batch has been (re)deposited: part of the NBD uniforms' address (operation 3)
  int e0;             // first event of the launch (the grid stages run in L2-sized sub-batches)
};
// H_RLO..H_CHI: bounding rectangle (cells) of the participant / collision deposits; H_SRLO..H_SCHI: that of the spectator deposits
enum { H_NP1 = 0, H_NP2, H_NCOLL, H_TRIES, H_NSPEC1, H_NSPEC2, H_STATUS, H_RLO, H_RHI, H_CLO, H_CHI, H_GIVENW, H_SRLO, H_SRHI, H_SCLO, H_SCHI, HDR_I = 16 };
enum { HD_B = 0, HDR_D = 4 };
enum { MOM_OUT = 64 };   // 0..44 mom[9][5], 45 rn0, 46 total, 47 xc, 48 yc, 49 dsdy

__device__ __forceinline__ double xg_of(const DevCfg& c, int i) { return __dadd_rn(c.Xmin, __dmul_rn((double)i, c.dx)); }
__device__ __forceinline__ double yg_of(const DevCfg& c, int j) { return __dadd_rn(c.Ymin, __dmul_rn((double)j, c.dy)); }
// (int)((v - lo)/d) exactly as the reference evaluates it (quirk Q7)
__device__ __forceinline__ int cell_of(double v, double lo, double d) { return (int)__ddiv_rn(__dadd_rn(v, -lo), d); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace smc
