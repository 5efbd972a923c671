// smc_avg.cu -- K6: averaged smooth profiles (operation 3).
//
// Replaces the body of MakeDensity::generate_profile_average (reference src/MakeDensity.cpp:1240-1577):
// per accepted event and per order n, [recentre -> redeposit -> accumulate] for the reaction-plane
// average, then [recentre + rotate by the participant-plane angle Psi_n -> redeposit -> accumulate],
// for the entropy branch and again for the energy branch, mutating the event's positions cumulatively
// exactly as the reference does.  GlueDensity::calcCMAngle (src/GlueDensity.cpp:87-144),
// MCnucl::recenterGrid / rotateGrid (src/MCnucl.cpp:1119-1174), Particle::rotate / calculateBounds
// (src/Particle.cpp:94-99,191-200; the stale-baseBox behaviour is SURVEY.md quirk Q4).
//
// The reference keeps a sequential running mean (old*(k-1)+new)/k; here every GPU keeps plain sums in
// device memory plus an event counter, so that the only cross-GPU step is one sum-allreduce of the
// accumulator block (smc_avg_device_buffer) -- the mean differs from the sequential form at 1e-16.
#include <algorithm>
#include <cstring>
#include "smc_ctx.h"

namespace smc {

#define AVG_THREADS 256
__device__ __forceinline__ double bsum(double v, double* red, int tid) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < AVG_THREADS / 32; w++) s += red[w];
  return s;
}

// GlueDensity::calcCMAngle: centre of mass and n-th order participant-plane angle of rho*scale
__global__ void __launch_bounds__(AVG_THREADS) cm_angle_kernel(DevCfg c, Store st, int order, double scale) {
  __shared__ double red[AVG_THREADS / 32];
  const int e = blockIdx.x, tid = threadIdx.x;
  const int* hi = st.hdr_i + (size_t)e * HDR_I;
  if (hi[H_STATUS] != 0) return;
  const size_t G = (size_t)c.Maxx * c.Maxy;
  const double* rho = st.grids + ((size_t)e * st.nkinds + st.kind_slot[GK_RHO]) * G;
  // rho vanishes outside the event's bounding rectangle (bbox_kernel) and that part of the lattice is not even written
  // (a flat walk over the rectangle's cells: a row/column walk without the two index divisions measured 14 % slower --
  // rectangles are ~130 columns wide, a fifth of the lanes of a row-per-warp walk idle)
  const int ilo = hi[H_RLO], jlo = hi[H_CLO], wj = max(hi[H_CHI] - jlo, 0), ncell = max(hi[H_RHI] - ilo, 0) * wj;
  double w = 0, sx = 0, sy = 0;
  for (int q = tid; q < ncell; q += AVG_THREADS) {
    const int i = ilo + q / wj, j = jlo + q % wj;
    const double wei = rho[(size_t)i * c.Maxy + j] * scale * c.dx * c.dy;
    w += wei; sx += xg_of(c, i) * wei; sy += yg_of(c, j) * wei;
  }
  const double weight = bsum(w, red, tid);
  const double xc = bsum(sx, red, tid) / weight, yc = bsum(sy, red, tid) / weight;
  double nr = 0, ni = 0;
  for (int q = tid; q < ncell; q += AVG_THREADS) {
    const int i = ilo + q / wj, j = jlo + q % wj;
    const double d = rho[(size_t)i * c.Maxy + j] * scale;
    if (d == 0.0) continue;
    const double x = xg_of(c, i) - xc, y = yg_of(c, j) - yc;
    double a = 1.0, b = 0.0;                 // (x + i y)^n = r^n e^{i n theta}
    for (int k2 = 0; k2 < order; k2++) { const double a2 = a * x - b * y; b = a * y + b * x; a = a2; }
    nr += a * d; ni += b * d;
  }
  const double Nr = bsum(nr, red, tid), Ni = bsum(ni, red, tid);
  if (tid == 0) { double* o = st.cm + (size_t)e * 4; o[0] = xc; o[1] = yc; o[2] = -atan2(-Ni, -Nr) / order; o[3] = weight; }
}

// recenterGrid (rotate=0) or rotateGrid (rotate=1) applied to participants, collisions and spectators
__global__ void transform_kernel(DevCfg c, Store st, int rotate) {
  const int e = blockIdx.x;
  const int* hi = st.hdr_i + (size_t)e * HDR_I;
  if (hi[H_STATUS] != 0) return;
  const double xc = st.cm[(size_t)e * 4], yc = st.cm[(size_t)e * 4 + 1], ang = st.cm[(size_t)e * 4 + 2];
  double sn, cs; sincos(ang, &sn, &cs);
  const int Amax = c.Amax;
  for (int k = threadIdx.x; k < c.A[0] + c.A[1]; k += blockDim.x) {
    const int s = k >= c.A[0], i = s ? k - c.A[0] : k;
    double* r = st.nuc + (((size_t)e * 2 + s) * Amax + i) * NROW;
    double* x = st.nuc_extra + (((size_t)e * 2 + s) * Amax + i) * NEXTRA;
    const bool part = st.nuc_ncoll[((size_t)e * 2 + s) * Amax + i] > 0;
    double px = r[NX] - xc, py = r[NY] - yc;               // setX(getX()-x), setY(getY()-y)   MCnucl.cpp:1125-1137
    if (part) {                                             // Box2D::setCenter keeps the extent, moves the centre
      const double ddx = __dadd_rn(px, -x[XCX]), ddy = __dadd_rn(py, -x[XCY]);
      r[NXL] = __dadd_rn(r[NXL], ddx); r[NXR] = __dadd_rn(r[NXR], ddx); r[NYL] = __dadd_rn(r[NYL], ddy); r[NYR] = __dadd_rn(r[NYR], ddy);
      x[XCX] = px; x[XCY] = py;
    }
    if (rotate) {                                           // Point3D::rotate(cos 0, angle)   MathBasics.cpp:41-50
      const double x0 = px, y0 = py;
      px = cs * x0 - sn * y0; py = sn * x0 + cs * y0;
      if (part) {
        // Particle::rotate turns the quark offsets too; calculateBounds() then rebuilds the AABB from the
        // STALE constructor-time base box and the quark boxes at the current position (quirk Q4)
        double xl = x[XBXL], xr = x[XBXR], yl = x[XBYL], yr = x[XBYR];
        const double hq = 8 * c.quark_width / 2;
        for (int q = 0; q < 3; q++) {
          const double qx = x[XQ + 3 * q], qy = x[XQ + 3 * q + 1];
          const double rx = cs * qx - sn * qy, ry = sn * qx + cs * qy;
          x[XQ + 3 * q] = rx; x[XQ + 3 * q + 1] = ry;
          xl = fmin(xl, px + rx - hq); xr = fmax(xr, px + rx + hq); yl = fmin(yl, py + ry - hq); yr = fmax(yr, py + ry + hq);
        }
        r[NXL] = xl; r[NXR] = xr; r[NYL] = yl; r[NYR] = yr;
        x[XCX] = (xl + xr) / 2.0; x[XCY] = (yl + yr) / 2.0;
      }
    }
    r[NX] = px; r[NY] = py;
  }
  int nc = hi[H_NCOLL]; if (nc > c.ncoll_cap) nc = c.ncoll_cap;
  for (int k = threadIdx.x; k < nc; k += blockDim.x) {
    double* cr = st.coll + ((size_t)e * c.ncoll_cap + k) * CROW;
    double px = cr[CX] - xc, py = cr[CY] - yc;
    if (rotate) { const double x0 = px, y0 = py; px = cs * x0 - sn * y0; py = sn * x0 + cs * y0; }
    cr[CX] = px; cr[CY] = py;
  }
}

// Accumulation of one density evaluation of a batch into the running sums (MakeDensity.cpp:1299-1330, 1341-1386).
// One thread owns one lattice cell and a slice of the batch's events; it reads rho, TA1, TA2, rho_binary (and the spectator
// lattices of the rotated pass) of every event of the slice whose bounding rectangle holds the cell -- the deposits of this mode
// only write rectangles (the spectator lattices one of their own) -- and keeps all seven sums in registers: every lattice is read once
// (TA1 and TA2 feed three sums), ACC_UNROLL events are in flight per thread, the rectangle test is done once per event.  The
// slices leave partial sums that acc_finish_kernel adds in slice order: the result does not depend on the launch geometry.
struct AccList { int with_spec; int dst0; double sd_scale; };      // dst0: accumulator slot of quantity 0 (the seven are consecutive)
#define ACC_SLICE 128      // events per slice (<= 8 slices per batch of 1024)
#define ACC_UNROLL 2
__global__ void __launch_bounds__(256, 3) accumulate_kernel(DevCfg c, Store st, AccList al, double* part, int nev) {
  __shared__ short4 rect[ACC_SLICE], rect2[ACC_SLICE];      // rectangle of the participant / collision deposits, and of the spectator deposits
  const size_t G = (size_t)c.Maxx * c.Maxy;
  const int e0 = blockIdx.y * ACC_SLICE, ne = min(ACC_SLICE, nev - e0);
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(k / c.Maxy), j = (int)(k % c.Maxy);
  for (int e = threadIdx.x; e < ACC_SLICE; e += blockDim.x) {
    short4 r = make_short4(0, 0, 0, 0), r2 = r;                   // empty: the event does not count (not accepted, or beyond the batch)
    if (e < ne) {
      const int* hi = st.hdr_i + (size_t)(e0 + e) * HDR_I;
      if (hi[H_STATUS] == 0) { r = make_short4((short)hi[H_RLO], (short)hi[H_RHI], (short)hi[H_CLO], (short)hi[H_CHI]);
                               r2 = make_short4((short)hi[H_SRLO], (short)hi[H_SRHI], (short)hi[H_SCLO], (short)hi[H_SCHI]); }
    }
    rect[e] = r; rect2[e] = r2;
  }
  __syncthreads();
  if (k >= G) return;
  const size_t g_rho = (size_t)st.kind_slot[GK_RHO] * G + k, g_ta1 = (size_t)st.kind_slot[GK_TA1] * G + k, g_ta2 = (size_t)st.kind_slot[GK_TA2] * G + k;
  const size_t g_bin = (size_t)st.kind_slot[GK_RHO_BINARY] * G + k, g_sa = (size_t)st.kind_slot[GK_SPEC_A] * G + k, g_sb = (size_t)st.kind_slot[GK_SPEC_B] * G + k;
  const bool spec = al.with_spec != 0;
  double s_sd = 0, s_tatb = 0, s_bin = 0, s_ta = 0, s_tb = 0, s_sa = 0, s_sb = 0;
  for (int eb = 0; eb < ne; eb += ACC_UNROLL) {
    double v_rho[ACC_UNROLL], v_ta1[ACC_UNROLL], v_ta2[ACC_UNROLL], v_bin[ACC_UNROLL], v_sa[ACC_UNROLL], v_sb[ACC_UNROLL];
#pragma unroll
    for (int u = 0; u < ACC_UNROLL; u++) {
      const short4 r = rect[min(eb + u, ACC_SLICE - 1)], r2 = rect2[min(eb + u, ACC_SLICE - 1)];
      const bool live = eb + u < ne, in = live && i >= r.x && i < r.y && j >= r.z && j < r.w;
      const bool in2 = spec && live && i >= r2.x && i < r2.y && j >= r2.z && j < r2.w;
      const double* base = st.grids + (size_t)(e0 + eb + u) * st.nkinds * G;
      v_rho[u] = in ? base[g_rho] : 0.0; v_ta1[u] = in ? base[g_ta1] : 0.0; v_ta2[u] = in ? base[g_ta2] : 0.0; v_bin[u] = in ? base[g_bin] : 0.0;
      v_sa[u] = in2 ? base[g_sa] : 0.0; v_sb[u] = in2 ? base[g_sb] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < ACC_UNROLL; u++) {                        // event order
      s_sd += v_rho[u] * al.sd_scale; s_tatb += v_ta1[u] * v_ta2[u]; s_bin += v_bin[u]; s_ta += v_ta1[u]; s_tb += v_ta2[u];
      s_sa += v_sa[u]; s_sb += v_sb[u];
    }
  }
  double* o = part + (size_t)blockIdx.y * SMC_AVG_QUANTITIES * G + k;
  o[(size_t)SMC_AVG_SD * G] = s_sd; o[(size_t)SMC_AVG_TATB * G] = s_tatb; o[(size_t)SMC_AVG_RHO_BINARY * G] = s_bin;
  o[(size_t)SMC_AVG_TA * G] = s_ta; o[(size_t)SMC_AVG_TB * G] = s_tb;
  if (spec) { o[(size_t)SMC_AVG_SPEC_A * G] = s_sa; o[(size_t)SMC_AVG_SPEC_B * G] = s_sb; }
}
// acc[dst[q]] += sum over the slices, in slice order
__global__ void __launch_bounds__(256) acc_finish_kernel(DevCfg c, AccList al, const double* part, double* acc, int nslices) {
  const size_t G = (size_t)c.Maxx * c.Maxy;
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (k >= G || (!al.with_spec && (q == SMC_AVG_SPEC_A || q == SMC_AVG_SPEC_B))) return;
  double s = 0.0;
  for (int sl = 0; sl < nslices; sl++) s += part[((size_t)sl * SMC_AVG_QUANTITIES + q) * G + k];
  acc[(size_t)(al.dst0 + q) * G + k] += s;
}
}  // namespace smc

// accumulator slot of (order index, variant 0 rotated / 1 reaction plane, branch 0 sd / 1 ed, quantity)
static inline int avg_slot(int io, int variant, int branch, int quantity) { return ((io * 2 + variant) * 2 + branch) * SMC_AVG_QUANTITIES + quantity; }

extern "C" int smc_avg_begin(smc_ctx* ctx, int from_order, int to_order, int with_rp, int branches) {
  if (!ctx || from_order < 1 || to_order < from_order || to_order > 9 || !(branches & 3)) return SMC_ERR_PARAM;
  CK(cudaSetDevice(ctx->device));
  ctx->avg_from = from_order; ctx->avg_to = to_order; ctx->avg_rp = with_rp ? 1 : 0; ctx->avg_ed = branches; ctx->avg_count = 0;
  const int norders = to_order - from_order + 1;
  const int64_t want = (int64_t)norders * 4 * SMC_AVG_QUANTITIES * (int64_t)ctx->G;
  if (ctx->d_avg && want != ctx->avg_doubles) { cudaFree(ctx->d_avg); ctx->d_avg = nullptr; }      // a run of the same shape reuses the block
  ctx->avg_doubles = want;
  if (!ctx->d_avg) CK(cudaMalloc(&ctx->d_avg, (size_t)ctx->avg_doubles * sizeof(double)));
  CK(cudaMemset(ctx->d_avg, 0, (size_t)ctx->avg_doubles * sizeof(double)));
  { const int rc = smc_ensure_extra(ctx); if (rc) return rc; }
  return SMC_OK;
}

// the per-order sequence on a batch whose event records are on the device
static int avg_sequence(smc_ctx* ctx, int m) {
  const smc::DevCfg& c = ctx->cfg; smc::Store& st = ctx->st;
  int kinds[8], nd = 0, rc;
  if ((rc = smc_plan_kinds(ctx, SMC_RUN_THICKNESS | SMC_RUN_RHO_BINARY | SMC_RUN_SPECTATORS, kinds, &nd))) return rc;
  ctx->need_zero = false;
  if (c.which_mc_model == 1 && !st.kln_table) FAIL(SMC_ERR_STATE, "MC-KLN needs its table first");
  // The reference recomputes every lattice at each of the five density steps of an order; only some are read before the
  // next move of the event: the order's first density feeds calcCMAngle alone (rho), the reaction-plane pass accumulates
  // everything but the spectator lattices (MakeDensity.cpp:1289-1330), the rotated pass everything.  Deposit what is read.
  int k_rho[8], n_rho = 0, k_rp[8], n_rp = 0;
  for (int i = 0; i < nd; i++) {
    const int k = kinds[i];
    const bool for_rho = (c.which_mc_model == 5) ? k == smc::GK_RHO : (c.which_mc_model == 7) ? (k == smc::GK_RHOA || k == smc::GK_RHOB) : (k == smc::GK_TA1 || k == smc::GK_TA2);
    if (for_rho || (c.cc_fluct == 2 && (k == smc::GK_TA1 || k == smc::GK_TA2))) k_rho[n_rho++] = k;
    if (k != smc::GK_SPEC_A && k != smc::GK_SPEC_B) k_rp[n_rp++] = k;
  }
  auto density = [&](const int* ks, int nk) -> int {          // calculateThickness + setDensity + calculate_rho_binary + calculate_spectator_density
    // no zero fill: deposit, combine, cm_angle and accumulate all work on the event's bounding rectangles (one for the
    // participant / collision lattices, one for the spectator lattices)
    CK(smc::launch_deposit(c, st, ks, nk, m, ctx->stream)); ctx->launches += 2;
    if (c.which_mc_model != 5) { CK(smc::launch_combine(c, st, m, ctx->stream)); ctx->launches++; }
    if (c.cc_fluct == 1 || c.cc_fluct == 2) { st.nbd_pass++; CK(smc::launch_fluctuate(c, st, m, ctx->stream)); ctx->launches++; }   // fresh draws per setDensity
    return SMC_OK;
  };
  auto cm = [&](int order, double scale) -> int { smc::cm_angle_kernel<<<m, AVG_THREADS, 0, ctx->stream>>>(c, st, order, scale); ctx->launches++; CK(cudaGetLastError()); return SMC_OK; };
  auto tf = [&](int rot) -> int { smc::transform_kernel<<<m, 128, 0, ctx->stream>>>(c, st, rot); ctx->launches++; CK(cudaGetLastError()); return SMC_OK; };
  auto acc = [&](int io, int variant, int branch) -> int {
    smc::AccList al;
    al.dst0 = avg_slot(io, variant, branch, 0);
    al.with_spec = (variant == 0); al.sd_scale = c.finalFactor;                        // setSd/setEd: rho * finalFactor
    const int nslices = (m + ACC_SLICE - 1) / ACC_SLICE;
    const size_t need = (size_t)nslices * SMC_AVG_QUANTITIES * ctx->G * sizeof(double);
    if (need > ctx->avg_part_bytes) {
      CK(cudaStreamSynchronize(ctx->stream));
      if (ctx->d_avg_part) { cudaFree(ctx->d_avg_part); ctx->d_avg_part = nullptr; ctx->avg_part_bytes = 0; }
      CK(cudaMalloc(&ctx->d_avg_part, need)); ctx->avg_part_bytes = need;
    }
    dim3 g((unsigned)((ctx->G + 255) / 256), nslices), g2((unsigned)((ctx->G + 255) / 256), SMC_AVG_QUANTITIES);
    smc::accumulate_kernel<<<g, 256, 0, ctx->stream>>>(c, st, al, ctx->d_avg_part, m);
    smc::acc_finish_kernel<<<g2, 256, 0, ctx->stream>>>(c, al, ctx->d_avg_part, ctx->d_avg, nslices); ctx->launches += 2; CK(cudaGetLastError());
    return SMC_OK;
  };
  ctx->epoch++;                          // positions and boxes are about to move: host mirrors of the lists are stale
  // order -> rapidity slice -> branch, as the reference nests them (MakeDensity.cpp:1268-1275): the moves are cumulative
  // across slices too ("different rapidity slices are rotated separately, and this does not quite make sense").  Every
  // slice is written to the same files at the end, so only the last slice's sums are kept.
  const size_t tsz = (size_t)c.kln_tmax * c.kln_tmax;
  for (int order = ctx->avg_from; order <= ctx->avg_to; order++) {
    const int io = order - ctx->avg_from;
    for (int iy = 0; iy < ctx->ny; iy++) {
      if (ctx->d_kln) st.kln_table = ctx->d_kln + (size_t)iy * tsz;
      const bool keep = (iy == ctx->ny - 1);
      if ((rc = density(k_rho, n_rho))) return rc;                                       // MakeDensity.cpp:1275
      for (int branch = 0; branch < 2; branch++) {
        if (!(ctx->avg_ed & (1 << branch))) continue;
        double scale = branch == 1 ? c.finalFactor : 1.0;                                // ed branch: setRho(rho*finalFactor) first (:1391-1396)
        if (ctx->avg_rp) {                                                               // :1289-1330 / :1397-1438
          if ((rc = cm(order, scale)) || (rc = tf(0)) || (rc = density(k_rp, n_rp)) || (keep && (rc = acc(io, 1, branch)))) return rc;
          scale = 1.0;
        }
        if ((rc = cm(order, scale)) || (rc = tf(1)) || (rc = density(kinds, nd)) || (keep && (rc = acc(io, 0, branch)))) return rc;   // :1331-1386 / :1439-1493
      }
    }
  }
  if (ctx->d_kln) st.kln_table = ctx->d_kln;
  return SMC_OK;
}

static int64_t count_ok(smc_ctx* ctx, int m) { int64_t n = 0; for (int e = 0; e < m; e++) n += ctx->h_hdr_i[(size_t)e * smc::HDR_I + smc::H_STATUS] == 0; return n; }

extern "C" int smc_avg_run(smc_ctx* ctx, uint64_t first_event_id, int n, smc_event_out* out) {
  if (!ctx || n < 0) return SMC_ERR_PARAM;
  if (!ctx->d_avg) FAIL(SMC_ERR_STATE, "smc_avg_begin first");
  CK(cudaSetDevice(ctx->device));
  int kinds[8], nd = 0, rc;
  std::vector<smc_event_out> tmp;
  for (int off = 0; off < n; off += ctx->batch) {
    const int m = std::min(ctx->batch, n - off);
    // the un-rotated event: moments for the caller, sum(rho) for the dS/dy window -- only the density itself is needed
    // here (avg_sequence plans and deposits the profile lattices of every step itself)
    if ((rc = smc_plan_kinds(ctx, SMC_RUN_MOMENTS, kinds, &nd))) return rc;
    ctx->need_zero = false;                // every consumer of this mode walks rectangles (smc_get_grid blanks the rest)
    if ((rc = smc_sample_batch(ctx, first_event_id + (uint64_t)off, m))) return rc;
    if ((rc = smc_events_first_pass(ctx, m, kinds, nd))) return rc;
    if (out) smc_fill_out(ctx, m, out + off);
    if ((rc = avg_sequence(ctx, m))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->avg_count += count_ok(ctx, m);
    ctx->last_n = m;
  }
  return SMC_OK;
}

extern "C" int smc_avg_run_from_positions(smc_ctx* ctx, int n, const smc_event_in* in, smc_event_out* out) {
  if (!ctx || n < 0 || (n > 0 && !in)) return SMC_ERR_PARAM;
  if (!ctx->d_avg) FAIL(SMC_ERR_STATE, "smc_avg_begin first");
  CK(cudaSetDevice(ctx->device));
  int kinds[8], nd = 0, rc; bool any_u, any_w;
  if ((rc = smc_check_positions(ctx, n, in, &any_u, &any_w))) return rc;
  for (int off = 0; off < n; off += ctx->batch) {
    const int m = std::min(ctx->batch, n - off);
    if ((rc = smc_plan_kinds(ctx, SMC_RUN_MOMENTS, kinds, &nd))) return rc;
    ctx->need_zero = false;
    if ((rc = smc_stage_positions(ctx, off, m, in, any_u, any_w))) return rc;
    if ((rc = smc_run_grid_stages(ctx, m, kinds, nd))) return rc;
    if ((rc = smc_fetch_results(ctx, m))) return rc;
    if (out) smc_fill_out(ctx, m, out + off);
    if ((rc = avg_sequence(ctx, m))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->avg_count += count_ok(ctx, m);
    ctx->last_n = m;
  }
  return SMC_OK;
}

extern "C" int smc_avg_device_buffer(smc_ctx* ctx, void** dev_ptr, int64_t* n_doubles) {
  if (!ctx || !dev_ptr || !n_doubles) return SMC_ERR_PARAM;
  if (!ctx->d_avg) FAIL(SMC_ERR_STATE, "smc_avg_begin first");
  *dev_ptr = ctx->d_avg; *n_doubles = ctx->avg_doubles; return SMC_OK;
}
extern "C" int smc_avg_count(smc_ctx* ctx, int64_t* count) { if (!ctx || !count) return SMC_ERR_PARAM; *count = ctx->avg_count; return SMC_OK; }
extern "C" int smc_avg_set_count(smc_ctx* ctx, int64_t count) { if (!ctx) return SMC_ERR_PARAM; ctx->avg_count = count; return SMC_OK; }

extern "C" int smc_avg_get(smc_ctx* ctx, int order, int variant, int quantity, int branch, double* host) {
  if (!ctx || !host || variant < 0 || variant > 1 || branch < 0 || branch > 1 || quantity < 0 || quantity >= SMC_AVG_QUANTITIES) return SMC_ERR_PARAM;
  if (!ctx->d_avg) FAIL(SMC_ERR_STATE, "smc_avg_begin first");
  if (order < ctx->avg_from || order > ctx->avg_to) FAIL(SMC_ERR_PARAM, "order outside [average_from_order, average_to_order]");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(host, ctx->d_avg + (size_t)avg_slot(order - ctx->avg_from, variant, branch, quantity) * ctx->G, ctx->G * sizeof(double), cudaMemcpyDeviceToHost));
  const double inv = ctx->avg_count > 0 ? 1.0 / (double)ctx->avg_count : 0.0;
  for (size_t k = 0; k < ctx->G; k++) host[k] *= inv;
  return SMC_OK;
}
