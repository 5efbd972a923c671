// smc_api.cu -- the C ABI (include/supermc_b200.h): context, device memory, batch orchestration.
// One context per GPU; every compute entry point fails loudly (SMC_ERR_CUDA) without a device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cub/cub.cuh>
#include "../../include/supermc_b200.h"
#include "smc_common.cuh"
#include "smc_host_math.h"

#include "smc_ctx.h"

static void slot_store(smc_ctx* ctx, smc_slot& sl);

extern "C" int smc_abi_version(void) { return SMC_ABI_VERSION; }

extern "C" void smc_params_default(smc_params* p) {   // reference parameters.dat
  std::memset(p, 0, sizeof *p);
  p->which_mc_model = 7; p->sub_model = 1; p->lambda = 0.138; p->tmax = 71; p->tmax_subdivision = 3;
  p->alpha = 0.118; p->aproj = 208; p->atarg = 208; p->shape_of_nucleons = 2; p->collision_criterion = 2;
  p->shape_of_entropy = 2; p->quark_width = 0.3; p->gauss_nucl_width = 0.812; p->ecm = 5020.; p->bmin = 0.; p->bmax = 20.;
  p->npmin = 2; p->npmax = 500; p->cutdsdy = 0; p->cutdsdy_lowerbound = 593.51; p->cutdsdy_upperbound = 889.53;
  p->randomseed = 1; p->finalfactor = 40.0; p->ecc_from_order = 1; p->ecc_to_order = 9;
  p->maxx = 15.; p->maxy = 15.; p->dx = 0.1; p->dy = 0.1; p->cc_fluctuation_model = 6; p->cc_fluctuation_gamma_theta = 0.75;
  p->pt_order = 1; p->gaussian_lambda = 4.14; p->cc_fluctuation_k = 0.75; p->ny = 1; p->ymax = 0.0; p->max_batch = 0; p->ncoll_cap = 0;
}

template <typename T> static int dalloc(smc_ctx* ctx, T** p, size_t n) {
  void* v = nullptr;
  cudaError_t e = cudaMalloc(&v, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return SMC_ERR_NOMEM; }
  cudaMemset(v, 0, std::max<size_t>(n, 1) * sizeof(T));      // record slots the kernels never write are copied to the host too
  ctx->owned.push_back(v); *p = (T*)v; return SMC_OK;
}

static int sampler_mode(int A, int nn_corr) {
  if (A == 1) return 1;
  if (A == 2) return 4;
  if (A == 3 || A == 4 || A == 12 || A == 16) return 2;
  if (nn_corr == 1 && (A == 197 || A == 208)) return 3;
  return 0;
}

extern "C" int smc_create(const smc_params* p, int device, smc_ctx** out) {
  if (!p || !out) return SMC_ERR_PARAM;
  *out = nullptr;
  smc_ctx* ctx = new smc_ctx();
  ctx->p = *p; ctx->device = device; ctx->launches = 0; ctx->last_ms = 0; ctx->last_n = 0; ctx->last_flags = 0;
  ctx->d_grids = nullptr; ctx->grids_bytes = 0; ctx->d_srcrec = nullptr; ctx->srcrec_bytes = 0; ctx->d_cmpart = nullptr; ctx->d_pair_u = nullptr; ctx->pair_u_bytes = 0; ctx->d_coll_w = nullptr; ctx->coll_w_bytes = 0;
  ctx->d_rcbk = nullptr; ctx->rcbk_q = ctx->rcbk_y = ctx->rcbk_k = 0;
  ctx->d_quark = nullptr; ctx->d_cfgtab[0] = ctx->d_cfgtab[1] = nullptr; ctx->d_kln = nullptr; ctx->d_avg = nullptr; ctx->d_avg_part = nullptr; ctx->avg_part_bytes = 0; ctx->avg_doubles = 0; ctx->avg_count = 0;
  ctx->stream = nullptr; ctx->ev0 = ctx->ev1 = nullptr; ctx->profile = 0; ctx->cur_slot = 0; ctx->comm = nullptr; ctx->epoch = 1; ctx->lists.epoch = 0; ctx->lists.n = 0; ctx->ny = p->ny; ctx->slice = 0; std::memset(&ctx->sortbuf, 0, sizeof ctx->sortbuf);
  std::memset(ctx->slots, 0, sizeof ctx->slots);
  for (int i = 0; i < 8; i++) { ctx->stage_ms[i] = 0; ctx->pev[i] = nullptr; }
  *out = ctx;      // returned even on failure so the caller can read smc_last_error
  const bool timing = getenv("SMC_TIMING") != nullptr;
  const auto tc0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) { if (timing) std::fprintf(stderr, "#   smc_create %-28s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count()); };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) FAIL(SMC_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) FAIL(SMC_ERR_CUDA, "device is not sm_100-class (this library holds sm_100a code only)");
  CK(cudaStreamCreate(&ctx->stream)); CK(cudaEventCreate(&ctx->ev0)); CK(cudaEventCreate(&ctx->ev1));
  for (int i = 0; i < 8; i++) CK(cudaEventCreate(&ctx->pev[i]));
  lap("CUDA context, stream, events");

  // ---- parameter checks (the reference prints and exits) ----
  if (p->which_mc_model != 1 && p->which_mc_model != 5 && p->which_mc_model != 7) FAIL(SMC_ERR_PARAM, "which_mc_model must be 1, 5 or 7");
  if (p->which_mc_model == 5 && p->sub_model != 1 && p->sub_model != 2) FAIL(SMC_ERR_PARAM, "MC-Glauber sub_model must be 1 or 2 (MCnucl.cpp:718-721)");
  if (p->which_mc_model == 1 && p->sub_model != 7 && p->sub_model != 100 && p->sub_model != 101) FAIL(SMC_ERR_PARAM, "MC-KLN sub_model must be 7 (KLN uGD), 100 or 101 (rcBK tables, src/ParamDefs.h)");
  if (p->collision_criterion == 4)
    FAIL(SMC_ERR_PARAM, "collision_criterion 4 (testCollisionFromDensity: a 0.02 fm Riemann sum per tested pair, GaussianNucleonsCal.cpp:99-117) is not built");
  if (p->ny < 1 || p->ny > 64) FAIL(SMC_ERR_PARAM, "ny (rapidity slices) must be 1..64");
  if (p->shape_of_nucleons < 1 || p->shape_of_nucleons > 4) FAIL(SMC_ERR_PARAM, "shape_of_nucleons must be 1, 2, 3 or 4");
  if (p->shape_of_nucleons == 3 && !(p->gaussian_lambda > 0)) FAIL(SMC_ERR_PARAM, "shape_of_nucleons 3 needs gaussian_lambda > 0");
  if (p->shape_of_entropy < 1 || p->shape_of_entropy > 3) FAIL(SMC_ERR_PARAM, "shape_of_entropy must be 1 (disk), 2 (gaussian) or 3 (valence quarks)");
  if (p->shape_of_entropy == 3 && p->which_mc_model == 1) FAIL(SMC_ERR_PARAM, "shape_of_entropy 3 only enters the MC-Glauber / sqrt(TA TB) densities (MCnucl.cpp:856)");
  if (p->shape_of_entropy == 3 && !(p->quark_width > 0)) FAIL(SMC_ERR_PARAM, "shape_of_entropy 3 needs quark_width > 0");
  if (p->aproj < 1 || p->atarg < 1 || p->aproj > 512 || p->atarg > 512) FAIL(SMC_ERR_PARAM, "Aproj/Atarg out of range");
  if (!(p->dx > 0) || !(p->dy > 0) || !(p->maxx > 0) || !(p->maxy > 0)) FAIL(SMC_ERR_PARAM, "bad grid");
  if (p->cc_fluctuation_model != 0 && p->cc_fluctuation_model != 1 && p->cc_fluctuation_model != 2 && p->cc_fluctuation_model != 6)
    FAIL(SMC_ERR_PARAM, "cc_fluctuation_model must be 0, 1, 2 or 6 (MCnucl.cpp:899-903)");
  if (p->cc_fluctuation_model == 1 && !(p->cc_fluctuation_k > 0)) FAIL(SMC_ERR_PARAM, "cc_fluctuation_model 1 needs cc_fluctuation_k > 0");

  smc_constants& k = ctx->k; smc::DevCfg& c = ctx->cfg;
  std::memset(&c, 0, sizeof c); std::memset(&ctx->st, 0, sizeof ctx->st);
  k.siginnn = smc_host::sigma_inel(p->ecm); k.siginnn200 = smc_host::sigma_inel(200.0);
  if (!smc_host::gaussian_nucleon(p->shape_of_nucleons, k.siginnn, p->gauss_nucl_width, p->gaussian_lambda, &k.width, &k.sigma_gg)) FAIL(SMC_ERR_PARAM, "unsupported shape_of_nucleons");
  k.dsq = 0.1 * k.siginnn / M_PI;
  k.maxx_cells = (int)((p->maxx - (-p->maxx)) / p->dx + 0.1) + 1; k.maxy_cells = (int)((p->maxy - (-p->maxy)) / p->dy + 0.1) + 1;
  k.kln_dt = 10.0 / k.siginnn; k.kln_tmax = p->tmax;
  if (p->shape_of_nucleons >= 2) { k.kln_tmax = p->tmax_subdivision * (p->tmax - 1) + 1; k.kln_dt /= p->tmax_subdivision; }   // MCnucl.cpp:915-919
  c.Maxx = k.maxx_cells; c.Maxy = k.maxy_cells; c.Xmin = -p->maxx; c.Ymin = -p->maxy; c.dx = p->dx; c.dy = p->dy;
  c.w = k.width; c.dsq = k.dsq; c.siginNN = k.siginnn; c.sigma_gg = k.sigma_gg; c.alpha = p->alpha;
  c.dmax = (p->shape_of_nucleons == 1) ? 2. * std::sqrt(k.dsq) : 5. * k.width;
  c.thrA = smc_host::sqrt_threshold(5 * k.width); c.thrB = 25. * (k.width * k.width);
  c.norm = 1 / (2 * M_PI * k.width * k.width); c.inv2w2 = 1.0 / (2 * k.width * k.width); c.areai = 10.0 / k.siginnn;
  c.recx = std::exp(-p->dx * p->dx / (k.width * k.width)); c.recy = std::exp(-p->dy * p->dy / (k.width * k.width));
  c.rclip_flat = std::sqrt(k.dsq) * (1.0 + 1e-12); c.kln_tmax_param = p->tmax;
  c.finalFactor = p->finalfactor;
  c.shape_of_nucleons = p->shape_of_nucleons; c.shape_of_entropy = p->shape_of_entropy;
  c.crit = (p->collision_criterion >= 1 && p->collision_criterion <= 3) ? p->collision_criterion
           : (p->shape_of_entropy == 2 ? 2 : p->shape_of_entropy == 3 ? 3 : 1);                       // MCnucl.cpp:364-384
  c.which_mc_model = p->which_mc_model; c.sub_model = p->sub_model; c.cc_fluct = p->cc_fluctuation_model; c.cc_k = p->cc_fluctuation_k;
  c.A[0] = p->aproj; c.A[1] = p->atarg; c.deformed[0] = p->proj_deformed; c.deformed[1] = p->targ_deformed;
  for (int s = 0; s < 2; s++) {
    smc_host::WoodsSaxon ws = smc_host::woods_saxon(c.A[s], c.deformed[s]);
    c.rad[s] = ws.rad; c.dr[s] = ws.dr; c.rmaxCut[s] = ws.rmaxCut; c.rwMax[s] = ws.rwMax; c.beta2[s] = ws.beta2; c.beta4[s] = ws.beta4;
    c.sampler[s] = sampler_mode(c.A[s], p->include_nn_correlation);
  }
  c.bmin = p->bmin; c.bmax = p->bmax; c.npmin = p->npmin; c.npmax = p->npmax;
  { const double eps = 1e-8, th = p->cc_fluctuation_gamma_theta > 0 ? p->cc_fluctuation_gamma_theta : 1.0, gk = 1. / th;   // MCnucl.cpp:1271-1301
    c.gam_k_part = (1 - p->alpha + eps) / 2. * gk; c.gam_th_part = 2. / (1 - p->alpha + eps) * th;
    c.gam_k_bin = (p->alpha + eps) * gk; c.gam_th_bin = 1. / (p->alpha + eps) * th; }
  c.quark_width = p->quark_width;
  c.quark_R = std::sqrt((3.0 / 2.0) * (k.width * k.width - p->quark_width * p->quark_width));   // Nucleus.cpp:31-32
  c.quark_rows = 0;
  { const double qw = p->quark_width > 0 ? p->quark_width : 0.3;
    c.q_inv2w2 = 1.0 / (2 * qw * qw); c.q_norm = 1 / (2 * M_PI * qw * qw); c.q_thr = 5 * qw;      // sic: squared distance against 5*width (Quark.cpp:19-20)
    c.q_recx = std::exp(-p->dx * p->dx / (qw * qw)); c.q_recy = std::exp(-p->dy * p->dy / (qw * qw)); c.q_reach = std::sqrt(5 * qw) * (1.0 + 1e-12); }
  ctx->need_quarks = (p->shape_of_entropy == 3 || c.crit == 3);
  c.seed_lo = (uint32_t)((uint64_t)p->randomseed); c.seed_hi = (uint32_t)(((uint64_t)p->randomseed) >> 32);
  c.Amax = std::max(p->aproj, p->atarg); c.Amax = (c.Amax + 7) & ~7;
  { long cap = p->ncoll_cap > 0 ? p->ncoll_cap : std::min<long>((long)p->aproj * p->atarg, 6144);
    cap = std::min<long>(cap, 60000); c.ncoll_cap = (int)std::max<long>(cap, 1); }
  c.ecc_from = p->ecc_from_order; c.ecc_to = p->ecc_to_order;
  c.kln_tmax = 0; c.kln_dT = k.kln_dt;
  { const double reach = std::max(c.dmax, std::sqrt(k.dsq)); c.wmax = (int)(2. * reach / std::min(p->dx, p->dy)) + 6; }
  c.inv_dx_f = (float)(1.0 / p->dx); c.inv_dy_f = (float)(1.0 / p->dy);
  ctx->G = (size_t)c.Maxx * c.Maxy;
  if (c.Maxy > 1024) FAIL(SMC_ERR_PARAM, "grids wider than 1024 cells are not supported");

  ctx->batch = p->max_batch > 0 ? p->max_batch : 1024;
  const int B = ctx->batch; smc::Store& st = ctx->st; int rc;
  if ((rc = dalloc(ctx, &st.nuc, (size_t)B * 2 * c.Amax * smc::NROW))) return rc;
  if ((rc = dalloc(ctx, &st.nuc_ncoll, (size_t)B * 2 * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &st.nuc_first, (size_t)B * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &st.coll, (size_t)B * c.ncoll_cap * smc::CROW))) return rc;
  if ((rc = dalloc(ctx, &st.coll_ij, (size_t)B * c.ncoll_cap))) return rc;
  if ((rc = dalloc(ctx, &st.part_idx, (size_t)B * 2 * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &st.spec_idx, (size_t)B * 2 * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &st.hdr_i, (size_t)B * smc::HDR_I))) return rc;
  if ((rc = dalloc(ctx, &st.hdr_d, (size_t)B * smc::HDR_D))) return rc;
  if ((rc = dalloc(ctx, &st.mom_out, (size_t)B * smc::MOM_OUT))) return rc;
  { uint64_t* ev; if ((rc = dalloc(ctx, &ev, (size_t)B))) return rc; st.event_id = ev; }
  if ((rc = dalloc(ctx, &st.try_start, (size_t)B))) return rc;
  if ((rc = dalloc(ctx, &ctx->d_redo, (size_t)B))) return rc;
  if ((rc = dalloc(ctx, &st.cm, (size_t)B * 4))) return rc;
  lap("device record buffers");
  CK(cudaMallocHost(&ctx->h_hdr_i, (size_t)B * smc::HDR_I * sizeof(int)));
  CK(cudaMallocHost(&ctx->h_hdr_d, (size_t)B * smc::HDR_D * sizeof(double)));
  CK(cudaMallocHost(&ctx->h_mom, (size_t)B * smc::MOM_OUT * sizeof(double)));
  CK(cudaMallocHost(&ctx->h_evid, (size_t)B * sizeof(uint64_t)));
  CK(cudaMallocHost(&ctx->h_try, (size_t)B * sizeof(int)));
  CK(cudaMallocHost(&ctx->h_nuc, (size_t)B * 2 * c.Amax * smc::NROW * sizeof(double)));
  for (int i = 0; i < 8; i++) st.kind_slot[i] = -1;
  lap("pinned host buffers");
  return SMC_OK;
}

extern "C" void smc_comm_finalize(smc_ctx* ctx);
extern "C" void smc_destroy(smc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  smc_comm_finalize(ctx);
  for (void* v : ctx->owned) cudaFree(v);
  if (ctx->d_grids) cudaFree(ctx->d_grids);
  if (ctx->d_srcrec) cudaFree(ctx->d_srcrec);
  if (ctx->d_cmpart) cudaFree(ctx->d_cmpart);
  if (ctx->d_pair_u) cudaFree(ctx->d_pair_u);
  if (ctx->d_coll_w) cudaFree(ctx->d_coll_w);
  if (ctx->d_quark) cudaFree(ctx->d_quark);
  if (ctx->d_cfgtab[0]) cudaFree(ctx->d_cfgtab[0]);
  if (ctx->d_cfgtab[1]) cudaFree(ctx->d_cfgtab[1]);
  if (ctx->d_kln) cudaFree(ctx->d_kln);
  if (ctx->d_rcbk) cudaFree(ctx->d_rcbk);
  if (ctx->d_avg) cudaFree(ctx->d_avg);
  if (ctx->d_avg_part) cudaFree(ctx->d_avg_part);
  cudaFree(ctx->sortbuf.k1); cudaFree(ctx->sortbuf.k2); cudaFree(ctx->sortbuf.v1); cudaFree(ctx->sortbuf.v2); cudaFree(ctx->sortbuf.tmp);
  bool any_slot = false;
  for (int q = 0; q < SMC_MAX_SLOTS; q++) any_slot |= ctx->slots[q].ready;
  if (any_slot) {     // return the active view to its slot, then release the others
    slot_store(ctx, ctx->slots[ctx->cur_slot]);
    for (int q = 0; q < SMC_MAX_SLOTS; q++) {
      smc_slot& o = ctx->slots[q];
      if (q == ctx->cur_slot || !o.ready || !o.stream) continue;
      if (o.d_grids) cudaFree(o.d_grids);
      if (o.d_srcrec) cudaFree(o.d_srcrec);
      if (o.d_cmpart) cudaFree(o.d_cmpart);
      cudaFreeHost(o.h_hdr_i); cudaFreeHost(o.h_hdr_d); cudaFreeHost(o.h_mom); cudaFreeHost(o.h_evid); cudaFreeHost(o.h_try);
      for (int i = 0; i < 8; i++) if (o.pev[i]) cudaEventDestroy(o.pev[i]);
      cudaStreamDestroy(o.stream);
    }
    for (int q = 0; q < SMC_MAX_SLOTS; q++) if (ctx->slots[q].ready && ctx->slots[q].done) cudaEventDestroy(ctx->slots[q].done);
  }
  if (ctx->h_hdr_i) cudaFreeHost(ctx->h_hdr_i);
  if (ctx->h_hdr_d) cudaFreeHost(ctx->h_hdr_d);
  if (ctx->h_mom) cudaFreeHost(ctx->h_mom);
  if (ctx->h_evid) cudaFreeHost(ctx->h_evid);
  if (ctx->h_try) cudaFreeHost(ctx->h_try);
  if (ctx->h_nuc) cudaFreeHost(ctx->h_nuc);
  for (int i = 0; i < 8; i++) if (ctx->pev[i]) cudaEventDestroy(ctx->pev[i]);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* smc_last_error(const smc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int smc_set_seed(smc_ctx* ctx, int64_t seed) {
  if (!ctx) return SMC_ERR_PARAM;
  ctx->p.randomseed = seed; ctx->cfg.seed_lo = (uint32_t)((uint64_t)seed); ctx->cfg.seed_hi = (uint32_t)(((uint64_t)seed) >> 32);
  return SMC_OK;
}
extern "C" int smc_max_batch(const smc_ctx* ctx) { return ctx ? ctx->batch : 0; }
extern "C" int smc_get_constants(const smc_ctx* ctx, smc_constants* c) { if (!ctx || !c) return SMC_ERR_PARAM; *c = ctx->k; return SMC_OK; }
extern "C" int smc_set_profiling(smc_ctx* ctx, int on) { if (!ctx) return SMC_ERR_PARAM; ctx->profile = on; for (int i = 0; i < 8; i++) ctx->stage_ms[i] = 0; return SMC_OK; }
extern "C" int smc_get_stage_ms(const smc_ctx* ctx, double* ms4) { if (!ctx || !ms4) return SMC_ERR_PARAM; for (int i = 0; i < 4; i++) ms4[i] = ctx->stage_ms[i]; return SMC_OK; }
extern "C" int64_t smc_kernel_launches(const smc_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" double smc_last_run_ms(const smc_ctx* ctx) { return ctx ? ctx->last_ms : 0.0; }

extern "C" int smc_load_quark_table(smc_ctx* ctx, const double* rows3, int n) {
  if (!ctx || !rows3 || n <= 0) return SMC_ERR_PARAM;
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_quark) { cudaFree(ctx->d_quark); ctx->d_quark = nullptr; }
  CK(cudaMalloc(&ctx->d_quark, (size_t)n * 3 * sizeof(double)));
  CK(cudaMemcpy(ctx->d_quark, rows3, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
  ctx->st.quark_table = ctx->d_quark; ctx->cfg.quark_rows = n;
  return SMC_OK;
}

extern "C" int smc_load_config_table(smc_ctx* ctx, int which, const double* xyz, int n_cfg, int a) {
  if (!ctx || which < 0 || which > 1 || !xyz || n_cfg <= 0) return SMC_ERR_PARAM;
  if (a != ctx->cfg.A[which]) FAIL(SMC_ERR_PARAM, "configuration table mass number does not match Aproj/Atarg");
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_cfgtab[which]) { cudaFree(ctx->d_cfgtab[which]); ctx->d_cfgtab[which] = nullptr; }
  const size_t n = (size_t)n_cfg * a * 3;
  CK(cudaMalloc(&ctx->d_cfgtab[which], n * sizeof(double)));
  CK(cudaMemcpy(ctx->d_cfgtab[which], xyz, n * sizeof(double), cudaMemcpyHostToDevice));
  ctx->st.cfg_table[which] = ctx->d_cfgtab[which]; ctx->cfg.ncfg[which] = n_cfg;
  return SMC_OK;
}

// ---- MC-KLN table --------------------------------------------------------------------------------
static void gauleg01(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0); w.assign(n, 0);
  for (int i = 0; i < (n + 1) / 2; i++) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1, z1;
    do {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1.0); z1 = z; z = z1 - p1 / pp;
    } while (std::fabs(z - z1) > 1e-15);
    x[i] = 0.5 * (1 - z); x[n - 1 - i] = 0.5 * (1 + z);
    w[i] = w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}

// rcBKfunc::rcBKfunc (src/rcBKfunc.cpp:15-210): the tabulated N_A(Y, kt) of every Q0^2 file + a natural cubic
// spline in kt per (file, Y-bin) (gsl_interp_cspline)
extern "C" int smc_load_rcbk_tables(smc_ctx* ctx, const double* kt, const double* na, int maxq0, int maxy, int maxkt) {
  if (!ctx || !kt || !na || maxq0 < 2 || maxy < 1 || maxkt < 3) return SMC_ERR_PARAM;
  if (ctx->p.sub_model != 100 && ctx->p.sub_model != 101) FAIL(SMC_ERR_STATE, "rcBK tables need sub_model 100 or 101");
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)maxq0 * maxy * maxkt;
  std::vector<double> y2(n, 0.0), u(maxkt);
  for (size_t t = 0; t < (size_t)maxq0 * maxy; t++) {
    const double* x = kt + t * maxkt; const double* y = na + t * maxkt; double* d2 = y2.data() + t * maxkt;
    d2[0] = 0.0; u[0] = 0.0;
    for (int i = 1; i + 1 < maxkt; i++) {
      const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]), pp = sig * d2[i - 1] + 2.0;
      d2[i] = (sig - 1.0) / pp;
      u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
      u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / pp;
    }
    d2[maxkt - 1] = 0.0;
    for (int k2 = maxkt - 2; k2 >= 0; k2--) d2[k2] = d2[k2] * d2[k2 + 1] + u[k2];
  }
  if (ctx->d_rcbk) { cudaFree(ctx->d_rcbk); ctx->d_rcbk = nullptr; }
  CK(cudaMalloc(&ctx->d_rcbk, 3 * n * sizeof(double)));
  CK(cudaMemcpy(ctx->d_rcbk, kt, n * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_rcbk + n, na, n * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_rcbk + 2 * n, y2.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  ctx->rcbk_q = maxq0; ctx->rcbk_y = maxy; ctx->rcbk_k = maxkt;
  return SMC_OK;
}

extern "C" int smc_set_kln_table(smc_ctx* ctx, const double* table, int tmax, double dt) {
  if (!ctx || !table || tmax < 3) return SMC_ERR_PARAM;
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_kln) { cudaFree(ctx->d_kln); ctx->d_kln = nullptr; }
  const size_t nt = (size_t)ctx->ny * tmax * tmax;            // one table per rapidity slice (dndyTable[iy], MCnucl.cpp:921-925)
  CK(cudaMalloc(&ctx->d_kln, nt * sizeof(double)));
  CK(cudaMemcpy(ctx->d_kln, table, nt * sizeof(double), cudaMemcpyHostToDevice));
  ctx->st.kln_table = ctx->d_kln; ctx->cfg.kln_tmax = tmax; ctx->cfg.kln_dT = dt;
  return SMC_OK;
}

extern "C" int smc_build_kln_table(smc_ctx* ctx, double* host_out) {
  if (!ctx) return SMC_ERR_PARAM;
  CK(cudaSetDevice(ctx->device));
  const int tmax = ctx->k.kln_tmax;
  if (ctx->p.sub_model >= 100 && !ctx->d_rcbk) FAIL(SMC_ERR_STATE, "rcBK uGD (sub_model 100/101) needs smc_load_rcbk_tables first (the javier/ table files, src/rcBKfunc.cpp:115-178)");
  const char* q = getenv("SMC_KLN_QUAD");     // "npt,nkt,nphi"
  int npt = 400, nkt = 200, nphi = 64;
  if (q) sscanf(q, "%d,%d,%d", &npt, &nkt, &nphi);
  std::vector<double> xp, wp, xk, wk, cp(nphi);
  gauleg01(npt, xp, wp); gauleg01(nkt, xk, wk);
  for (int i = 0; i < nphi; i++) cp[i] = std::cos(2 * M_PI * ((i + 0.5) / nphi));
  double* d = nullptr;
  const size_t nn = (size_t)2 * npt + 2 * nkt + nphi;
  CK(cudaMalloc(&d, nn * sizeof(double)));
  std::vector<double> h; h.insert(h.end(), xp.begin(), xp.end()); h.insert(h.end(), wp.begin(), wp.end());
  h.insert(h.end(), xk.begin(), xk.end()); h.insert(h.end(), wk.begin(), wk.end()); h.insert(h.end(), cp.begin(), cp.end());
  CK(cudaMemcpy(d, h.data(), nn * sizeof(double), cudaMemcpyHostToDevice));
  if (ctx->d_kln) { cudaFree(ctx->d_kln); ctx->d_kln = nullptr; }
  const size_t nt = (size_t)ctx->ny * tmax * tmax;
  CK(cudaMalloc(&ctx->d_kln, nt * sizeof(double)));
  smc::KlnCfg kc; kc.ecm = ctx->p.ecm; kc.lambda = ctx->p.lambda; kc.dT = ctx->k.kln_dt; kc.tmax = tmax; kc.pt_order = ctx->p.pt_order > 0 ? ctx->p.pt_order : 1;
  kc.model = ctx->p.sub_model; kc.maxQ0 = ctx->rcbk_q; kc.maxY = ctx->rcbk_y; kc.maxKt = ctx->rcbk_k; kc.dQ0 = ctx->p.sub_model == 100 ? 0.1 : 0.168;
  kc.siginNN200 = ctx->k.siginnn200;
  { const size_t nn2 = (size_t)ctx->rcbk_q * ctx->rcbk_y * ctx->rcbk_k; kc.rkt = ctx->d_rcbk; kc.rna = ctx->d_rcbk ? ctx->d_rcbk + nn2 : nullptr; kc.ry2 = ctx->d_rcbk ? ctx->d_rcbk + 2 * nn2 : nullptr; }
  kc.npt = npt; kc.nkt = nkt; kc.nphi = nphi; kc.xp = d; kc.wp = d + npt; kc.xk = d + 2 * npt; kc.wk = d + 2 * npt + nkt; kc.cphi = d + 2 * npt + 2 * nkt;
  for (int iy = 0; iy < ctx->ny; iy++) {
    // y = rapMin + (rapMax - rapMin) / binRapidity * iy with rapMin = -ymax, rapMax = ymax (MakeDensity.cpp:54-56, MCnucl.cpp:932)
    kc.y = -ctx->p.ymax + (ctx->p.ymax - (-ctx->p.ymax)) / ctx->ny * iy;
    CK(smc::launch_kln_table(kc, ctx->d_kln + (size_t)iy * tmax * tmax, ctx->stream)); ctx->launches++;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  ctx->st.kln_table = ctx->d_kln; ctx->cfg.kln_tmax = tmax; ctx->cfg.kln_dT = ctx->k.kln_dt;
  if (host_out) CK(cudaMemcpy(host_out, ctx->d_kln, nt * sizeof(double), cudaMemcpyDeviceToHost));
  return SMC_OK;
}

// ---- event batches -------------------------------------------------------------------------------
// per-nucleon state beyond the 8-double rows (stale base boxes, valence-quark offsets and weights): operation 3, the
// quark substructure options and the quarks.data writer need it; allocated for the active pipeline slot on first use
int smc_ensure_extra(smc_ctx* ctx) {
  if (ctx->st.nuc_extra) return SMC_OK;
  const size_t nx = (size_t)ctx->batch * 2 * ctx->cfg.Amax * smc::NEXTRA;
  CK(cudaMalloc(&ctx->st.nuc_extra, nx * sizeof(double))); ctx->owned.push_back(ctx->st.nuc_extra);
  CK(cudaMalloc(&ctx->st.nuc_extra_tmp, nx * sizeof(double))); ctx->owned.push_back(ctx->st.nuc_extra_tmp);
  CK(cudaMemset(ctx->st.nuc_extra, 0, nx * sizeof(double)));
  return SMC_OK;
}

int smc_plan_kinds(smc_ctx* ctx, unsigned flags, int* kinds, int* nk_dep) {
  smc::Store& st = ctx->st; const smc::DevCfg& c = ctx->cfg;
  for (int i = 0; i < 8; i++) st.kind_slot[i] = -1;
  int n = 0, nd = 0;
  auto add = [&](int kind, bool dep) { if (st.kind_slot[kind] < 0) { st.kind_slot[kind] = n++; if (dep) kinds[nd++] = kind; } };
  if (c.which_mc_model == 5) add(smc::GK_RHO, true);
  else if (c.which_mc_model == 7) { add(smc::GK_RHO, false); add(smc::GK_RHOA, true); add(smc::GK_RHOB, true); }
  else { add(smc::GK_RHO, false); add(smc::GK_TA1, true); add(smc::GK_TA2, true); }
  if ((flags & SMC_RUN_THICKNESS) || c.cc_fluct == 2) { add(smc::GK_TA1, true); add(smc::GK_TA2, true); }     // NBD model 2: k from min(TA, TB)
  if (flags & SMC_RUN_RHO_BINARY) add(smc::GK_RHO_BINARY, true);
  if (flags & SMC_RUN_SPECTATORS) { add(smc::GK_SPEC_A, true); add(smc::GK_SPEC_B, true); }
  st.nkinds = n; *nk_dep = nd;
  // Scan mode (moments only): deposit tiles, combine and moments all work on the event's bounding rectangle, so the
  // lattice outside it is never read on the device and the getters blank it on the host (smc_get_grid).  Profile
  // modes hand whole grids to the host and to the averaging kernels: those start from zeros.
  ctx->need_zero = (flags & ~(unsigned)SMC_RUN_MOMENTS) != 0;
  const size_t need = (size_t)ctx->batch * n * ctx->G * sizeof(double);
  if (need > ctx->grids_bytes) {
    if (ctx->d_grids) cudaFree(ctx->d_grids);
    ctx->d_grids = nullptr; ctx->grids_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->d_grids, need);
    if (e != cudaSuccess) { ctx->err = std::string("grid pool: ") + cudaGetErrorString(e); return SMC_ERR_NOMEM; }
    ctx->grids_bytes = need;
  }
  st.grids = ctx->d_grids;
  st.src_stride = (c.shape_of_entropy == 3 ? 6 : 2) * c.Amax + c.ncoll_cap;
  if (ctx->need_quarks || (flags & SMC_RUN_LISTS)) { int rc = smc_ensure_extra(ctx); if (rc) return rc; }
  st.work_off = (((size_t)ctx->batch * std::max(nd, 1) * st.src_stride * sizeof(smc::SrcRec)) + 15) & ~(size_t)15;
  st.work_cap = ctx->batch * std::max(nd, 1) * smc::deposit_cm_slots(c);
  const size_t need_rec = st.work_off + smc::deposit_work_bytes(c, ctx->batch, std::max(nd, 1));
  if (need_rec > ctx->srcrec_bytes) {
    if (ctx->d_srcrec) cudaFree(ctx->d_srcrec);
    ctx->d_srcrec = nullptr; ctx->srcrec_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->d_srcrec, need_rec);
    if (e != cudaSuccess) { ctx->err = std::string("source record pool: ") + cudaGetErrorString(e); return SMC_ERR_NOMEM; }
    ctx->srcrec_bytes = need_rec;
  }
  st.src_rec = (smc::SrcRec*)ctx->d_srcrec;
  st.cm_slots = smc::deposit_cm_slots(c);
  if (!ctx->d_cmpart) {
    cudaError_t e = cudaMalloc(&ctx->d_cmpart, (size_t)ctx->batch * st.cm_slots * 32 * 4 * sizeof(double));
    if (e != cudaSuccess) { ctx->err = std::string("cm partial sums: ") + cudaGetErrorString(e); return SMC_ERR_NOMEM; }
  }
  st.cm_part = ctx->d_cmpart;
  return SMC_OK;
}

int smc_run_grid_stages(smc_ctx* ctx, int m, const int* kinds, int nd);
static int run_grid_stages(smc_ctx* ctx, int m, const int* kinds, int nd) { return smc_run_grid_stages(ctx, m, kinds, nd); }
int smc_run_grid_stages(smc_ctx* ctx, int m, const int* kinds, int nd) {
  const smc::DevCfg& c = ctx->cfg;
  if (c.which_mc_model == 1 && !ctx->st.kln_table) FAIL(SMC_ERR_STATE, "MC-KLN density requires smc_build_kln_table / smc_set_kln_table first (MCnucl.cpp:636-640)");
  ctx->st.nbd_pass = ctx->slice;   // the first density of an event (one per rapidity slice); operation 3 counts its re-deposits from here
  if (ctx->d_kln) ctx->st.kln_table = ctx->d_kln + (size_t)ctx->slice * c.kln_tmax * c.kln_tmax;
  if (ctx->profile) CK(cudaEventRecord(ctx->pev[1], ctx->stream));
  // deposit CTAs only cover each event's bounding rectangle: whoever reads whole grids needs zeros elsewhere
  if (ctx->need_zero) {
    const size_t ev_bytes = (size_t)ctx->st.nkinds * ctx->G * sizeof(double);
    if (!ctx->st.redo) CK(cudaMemsetAsync(ctx->d_grids, 0, (size_t)m * ev_bytes, ctx->stream));
    else      // dS/dy-window re-runs touch only the flagged events (h_try holds the flags): the others keep their grids
      for (int e = 0; e < m; e++) if (ctx->h_try[e]) CK(cudaMemsetAsync(ctx->d_grids + (size_t)e * ctx->st.nkinds * ctx->G, 0, ev_bytes, ctx->stream));
  }
  // Sub-batches sized so that the density tiles a deposit launch writes are still in the 126 MB L2 when the
  // moments launch reads them back (profiling mode keeps whole-batch launches: one event pair per stage)
  static const int sub_env = getenv("SMC_SUBBATCH") ? atoi(getenv("SMC_SUBBATCH")) : 0;
  const int sub = (!ctx->profile && sub_env > 0) ? sub_env : m;
  for (int e0 = 0; e0 < m; e0 += sub) {
    const int mm = std::min(sub, m - e0);
    ctx->st.e0 = e0;
    CK(smc::launch_deposit(c, ctx->st, kinds, nd, mm, ctx->stream)); ctx->launches += 2;
    if (ctx->profile) CK(cudaEventRecord(ctx->pev[2], ctx->stream));
    if (c.which_mc_model != 5) { CK(smc::launch_combine(c, ctx->st, mm, ctx->stream)); ctx->launches++; }
    smc::Store stm = ctx->st;
    if (c.cc_fluct == 1 || c.cc_fluct == 2) {          // MCnucl::fluctuateCurrentDensity, MCnucl.cpp:819,868-905
      CK(smc::launch_fluctuate(c, ctx->st, mm, ctx->stream)); ctx->launches++;
      stm.cm_part = nullptr;                           // the partial sums of the deposit describe the smooth density
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->pev[3], ctx->stream));
    CK(smc::launch_moments(c, stm, mm, ctx->stream)); ctx->launches++;
    if (ctx->profile) CK(cudaEventRecord(ctx->pev[4], ctx->stream));
  }
  ctx->st.e0 = 0;
  return SMC_OK;
}

// per-stage device time of the batch that was just synchronised (profiling mode only)
static void collect_stage_ms(smc_ctx* ctx) {
  if (!ctx->profile) return;
  for (int i = 0; i < 4; i++) { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->pev[i], ctx->pev[i + 1]) == cudaSuccess) ctx->stage_ms[i] += ms; }
}

void smc_fill_out(smc_ctx* ctx, int m, smc_event_out* out) {
  for (int e = 0; e < m; e++) {
    const int* hi = ctx->h_hdr_i + (size_t)e * smc::HDR_I; const double* mo = ctx->h_mom + (size_t)e * smc::MOM_OUT;
    smc_event_out& o = out[e];
    o.b = ctx->h_hdr_d[(size_t)e * smc::HDR_D + smc::HD_B];
    o.npart1 = hi[smc::H_NP1]; o.npart2 = hi[smc::H_NP2]; o.ncoll = hi[smc::H_NCOLL]; o.tries = hi[smc::H_TRIES];
    o.nspec = hi[smc::H_NSPEC1] + hi[smc::H_NSPEC2]; o.status = hi[smc::H_STATUS];
    std::memcpy(o.mom, mo, 45 * sizeof(double));
    o.rn0 = mo[45]; o.total = mo[46]; o.xc = mo[47]; o.yc = mo[48]; o.dsdy = mo[49]; o.nonzero_cells = (int)mo[50];
  }
}

int smc_fetch_results(smc_ctx* ctx, int m) {
  CK(cudaMemcpyAsync(ctx->h_hdr_i, ctx->st.hdr_i, (size_t)m * smc::HDR_I * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_hdr_d, ctx->st.hdr_d, (size_t)m * smc::HDR_D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_mom, ctx->st.mom_out, (size_t)m * smc::MOM_OUT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SMC_OK;
}

// ---- pipeline slots ------------------------------------------------------------------------------
static void slot_store(smc_ctx* ctx, smc_slot& sl) {          // ctx (active view) -> slot
  smc::Store& st = ctx->st;
  sl.nuc = st.nuc; sl.nuc_ncoll = st.nuc_ncoll; sl.nuc_first = st.nuc_first; sl.coll = st.coll; sl.coll_ij = st.coll_ij;
  sl.part_idx = st.part_idx; sl.spec_idx = st.spec_idx; sl.hdr_i = st.hdr_i; sl.hdr_d = st.hdr_d; sl.mom_out = st.mom_out;
  sl.event_id = (uint64_t*)st.event_id; sl.try_start = st.try_start; sl.cm = st.cm; sl.d_redo = ctx->d_redo;
  sl.nuc_extra = st.nuc_extra; sl.nuc_extra_tmp = st.nuc_extra_tmp;
  sl.d_grids = ctx->d_grids; sl.grids_bytes = ctx->grids_bytes; sl.d_srcrec = ctx->d_srcrec; sl.srcrec_bytes = ctx->srcrec_bytes; sl.d_cmpart = ctx->d_cmpart; sl.stream = ctx->stream;
  for (int i = 0; i < 8; i++) sl.pev[i] = ctx->pev[i];
  sl.h_hdr_i = ctx->h_hdr_i; sl.h_hdr_d = ctx->h_hdr_d; sl.h_mom = ctx->h_mom; sl.h_evid = ctx->h_evid; sl.h_try = ctx->h_try;
}
static void slot_load(smc_ctx* ctx, const smc_slot& sl) {     // slot -> ctx (active view)
  smc::Store& st = ctx->st;
  st.nuc = sl.nuc; st.nuc_ncoll = sl.nuc_ncoll; st.nuc_first = sl.nuc_first; st.coll = sl.coll; st.coll_ij = sl.coll_ij;
  st.part_idx = sl.part_idx; st.spec_idx = sl.spec_idx; st.hdr_i = sl.hdr_i; st.hdr_d = sl.hdr_d; st.mom_out = sl.mom_out;
  st.event_id = sl.event_id; st.try_start = sl.try_start; st.cm = sl.cm; ctx->d_redo = sl.d_redo;
  st.nuc_extra = sl.nuc_extra; st.nuc_extra_tmp = sl.nuc_extra_tmp;
  ctx->d_grids = sl.d_grids; ctx->grids_bytes = sl.grids_bytes; st.grids = sl.d_grids; ctx->stream = sl.stream;
  ctx->d_srcrec = sl.d_srcrec; ctx->srcrec_bytes = sl.srcrec_bytes; st.src_rec = (smc::SrcRec*)sl.d_srcrec;
  ctx->d_cmpart = sl.d_cmpart; st.cm_part = sl.d_cmpart;
  for (int i = 0; i < 8; i++) ctx->pev[i] = sl.pev[i];
  ctx->h_hdr_i = sl.h_hdr_i; ctx->h_hdr_d = sl.h_hdr_d; ctx->h_mom = sl.h_mom; ctx->h_evid = sl.h_evid; ctx->h_try = sl.h_try;
}
static int slot_alloc(smc_ctx* ctx, smc_slot& sl) {            // the second slot: same sizes as the first
  const smc::DevCfg& c = ctx->cfg; const int B = ctx->batch; int rc;
  if ((rc = dalloc(ctx, &sl.nuc, (size_t)B * 2 * c.Amax * smc::NROW))) return rc;
  if ((rc = dalloc(ctx, &sl.nuc_ncoll, (size_t)B * 2 * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &sl.nuc_first, (size_t)B * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &sl.coll, (size_t)B * c.ncoll_cap * smc::CROW))) return rc;
  if ((rc = dalloc(ctx, &sl.coll_ij, (size_t)B * c.ncoll_cap))) return rc;
  if ((rc = dalloc(ctx, &sl.part_idx, (size_t)B * 2 * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &sl.spec_idx, (size_t)B * 2 * c.Amax))) return rc;
  if ((rc = dalloc(ctx, &sl.hdr_i, (size_t)B * smc::HDR_I))) return rc;
  if ((rc = dalloc(ctx, &sl.hdr_d, (size_t)B * smc::HDR_D))) return rc;
  if ((rc = dalloc(ctx, &sl.mom_out, (size_t)B * smc::MOM_OUT))) return rc;
  if ((rc = dalloc(ctx, &sl.event_id, (size_t)B))) return rc;
  if ((rc = dalloc(ctx, &sl.try_start, (size_t)B))) return rc;
  if ((rc = dalloc(ctx, &sl.cm, (size_t)B * 4))) return rc;
  if ((rc = dalloc(ctx, &sl.d_redo, (size_t)B))) return rc;
  sl.d_grids = nullptr; sl.grids_bytes = 0; sl.d_srcrec = nullptr; sl.srcrec_bytes = 0; sl.d_cmpart = nullptr;
  sl.nuc_extra = nullptr; sl.nuc_extra_tmp = nullptr;
  CK(cudaStreamCreate(&sl.stream));
  for (int i = 0; i < 8; i++) CK(cudaEventCreate(&sl.pev[i]));
  CK(cudaMallocHost(&sl.h_hdr_i, (size_t)B * smc::HDR_I * sizeof(int)));
  CK(cudaMallocHost(&sl.h_hdr_d, (size_t)B * smc::HDR_D * sizeof(double)));
  CK(cudaMallocHost(&sl.h_mom, (size_t)B * smc::MOM_OUT * sizeof(double)));
  CK(cudaMallocHost(&sl.h_evid, (size_t)B * sizeof(uint64_t)));
  CK(cudaMallocHost(&sl.h_try, (size_t)B * sizeof(int)));
  return SMC_OK;
}
int smc_activate_slot(smc_ctx* ctx, int s) {
  if (s == ctx->cur_slot && ctx->slots[s].ready) return SMC_OK;
  if (!ctx->slots[ctx->cur_slot].ready) { CK(cudaEventCreateWithFlags(&ctx->slots[ctx->cur_slot].done, cudaEventDisableTiming)); ctx->slots[ctx->cur_slot].ready = true; }
  slot_store(ctx, ctx->slots[ctx->cur_slot]);
  if (!ctx->slots[s].ready) {
    int rc = slot_alloc(ctx, ctx->slots[s]); if (rc) return rc;
    CK(cudaEventCreateWithFlags(&ctx->slots[s].done, cudaEventDisableTiming)); ctx->slots[s].ready = true;
  }
  slot_load(ctx, ctx->slots[s]); ctx->cur_slot = s;
  return SMC_OK;
}

// one batch of sampled events: ids -> device, K1+K2
int smc_sample_batch(smc_ctx* ctx, uint64_t first_event_id, int m) {
  ctx->epoch++;
  for (int e = 0; e < m; e++) { ctx->h_evid[e] = first_event_id + (uint64_t)e; ctx->h_try[e] = 0; }
  CK(cudaMemcpyAsync((void*)ctx->st.event_id, ctx->h_evid, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->st.try_start, ctx->h_try, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ctx->st.pair_u = nullptr; ctx->st.coll_w = nullptr; ctx->st.redo = nullptr;
  if (ctx->profile) CK(cudaEventRecord(ctx->pev[0], ctx->stream));
  CK(smc::launch_sample_collide(ctx->cfg, ctx->st, m, false, ctx->stream)); ctx->launches++;
  return SMC_OK;
}

// dS/dy window (MakeDensity.cpp:2173-2181): an event whose sum(rho) dx dy falls outside goes back to the
// rejection loop; its Philox stream simply continues at the next try index.
static int dsdy_cut_loop(smc_ctx* ctx, int m, const int* kinds, int nd) {
  if (ctx->p.cutdsdy != 1) return SMC_OK;
  int rc;
  for (int iter = 0; iter < 100000; iter++) {
    int nbad = 0;
    for (int e = 0; e < m; e++) {
      const double s = ctx->h_mom[(size_t)e * smc::MOM_OUT + 49];
      const bool bad = (ctx->h_hdr_i[(size_t)e * smc::HDR_I + smc::H_STATUS] == 0) && (s < ctx->p.cutdsdy_lowerbound || s > ctx->p.cutdsdy_upperbound);
      ctx->h_try[e] = bad ? 1 : 0; nbad += bad;
    }
    if (!nbad) break;
    CK(cudaMemcpyAsync(ctx->d_redo, ctx->h_try, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ctx->st.redo = ctx->d_redo;
    CK(smc::launch_sample_collide(ctx->cfg, ctx->st, m, false, ctx->stream)); ctx->launches++;
    if ((rc = run_grid_stages(ctx, m, kinds, nd))) { ctx->st.redo = nullptr; return rc; }
    ctx->st.redo = nullptr;
    if ((rc = smc_fetch_results(ctx, m))) return rc;
  }
  return SMC_OK;
}

// density + moments + results to the host + dS/dy window, for a batch whose records are on the device
int smc_events_first_pass(smc_ctx* ctx, int m, const int* kinds, int nd) {
  int rc;
  if ((rc = run_grid_stages(ctx, m, kinds, nd))) return rc;
  if ((rc = smc_fetch_results(ctx, m))) return rc;
  collect_stage_ms(ctx);
  return dsdy_cut_loop(ctx, m, kinds, nd);
}

// ny > 1 (MakeDensity.cpp:2170-2193): one more density + one more row per rapidity slice.  The event records and the
// deposit inputs do not depend on the slice; MC-KLN looks its density up in the table of slice iy, the NBD fluctuation
// draws afresh.  Rows go to out[(e * ny) + iy]; the grids left on the device are those of the last slice (the reference
// writes every slice to the same file name, so the last one is what its files hold).
static int run_slices(smc_ctx* ctx, int m, const int* kinds, int nd, smc_event_out* out) {
  const int ny = ctx->ny; int rc;
  std::vector<smc_event_out> tmp(m);
  smc_fill_out(ctx, m, tmp.data());
  for (int e = 0; e < m; e++) out[(size_t)e * ny] = tmp[e];
  for (int iy = 1; iy < ny; iy++) {
    ctx->slice = iy;
    if ((rc = smc_run_grid_stages(ctx, m, kinds, nd)) || (rc = smc_fetch_results(ctx, m))) { ctx->slice = 0; return rc; }
    smc_fill_out(ctx, m, tmp.data());
    for (int e = 0; e < m; e++) out[(size_t)e * ny + iy] = tmp[e];
  }
  ctx->slice = 0;
  return SMC_OK;
}

extern "C" int smc_run_events(smc_ctx* ctx, uint64_t first_event_id, int n, unsigned flags, smc_event_out* out) {
  if (!ctx || n < 0 || (!out && n > 0)) return SMC_ERR_PARAM;
  CK(cudaSetDevice(ctx->device));
  const smc::DevCfg& c = ctx->cfg;
  for (int s = 0; s < 2; s++) if ((c.sampler[s] == 2 || c.sampler[s] == 3) && !ctx->st.cfg_table[s]) FAIL(SMC_ERR_STATE, "this nucleus needs a configuration table: call smc_load_config_table (Nucleus.cpp:150-169)");
  if ((flags & ~(unsigned)SMC_RUN_MOMENTS) && n > ctx->batch)
    FAIL(SMC_ERR_PARAM, "grids and lists are kept for one device batch: with KEEP_RHO / THICKNESS / RHO_BINARY / SPECTATORS / LISTS call smc_run_events with n <= smc_max_batch()");
  int kinds[8], nd = 0, rc;
  if ((rc = smc_plan_kinds(ctx, flags, kinds, &nd))) return rc;
  const int nb = (n + ctx->batch - 1) / ctx->batch;
  if (nb >= 2 && ctx->p.cutdsdy != 1 && !ctx->profile && ctx->ny == 1 && !getenv("SMC_NO_PIPELINE")) {
    // software pipeline over NS slots: batch i runs on the stream of slot i % NS; its rows are read back when the slot
    // comes round again, so K1/K2 of one batch overlap K3/K4 and the copies of the others
    static const int ns_env = getenv("SMC_SLOTS") ? atoi(getenv("SMC_SLOTS")) : 4;
    const int NS = std::max(2, std::min(std::min(ns_env, SMC_MAX_SLOTS), nb));
    for (int sidx = 0; sidx < NS; sidx++) { if ((rc = smc_activate_slot(ctx, sidx))) return rc; if ((rc = smc_plan_kinds(ctx, flags, kinds, &nd))) return rc; }
    CK(cudaEventRecord(ctx->ev0, ctx->slots[0].stream));
    for (int i = 0; i < nb + NS; i++) {
      if ((rc = smc_activate_slot(ctx, i % NS))) return rc;
      if (i >= NS) {                                    // retire batch i-NS of this slot
        CK(cudaEventSynchronize(ctx->slots[i % NS].done));
        const int off = (i - NS) * ctx->batch, m = std::min(ctx->batch, n - off);
        smc_fill_out(ctx, m, out + off); ctx->last_n = m;
      }
      if (i < nb) {
        const int off = i * ctx->batch, m = std::min(ctx->batch, n - off);
        if ((rc = smc_sample_batch(ctx, first_event_id + (uint64_t)off, m))) return rc;
        if ((rc = run_grid_stages(ctx, m, kinds, nd))) return rc;
        CK(cudaMemcpyAsync(ctx->h_hdr_i, ctx->st.hdr_i, (size_t)m * smc::HDR_I * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_hdr_d, ctx->st.hdr_d, (size_t)m * smc::HDR_D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_mom, ctx->st.mom_out, (size_t)m * smc::MOM_OUT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->slots[i % NS].done, ctx->stream));
      }
    }
    if ((rc = smc_activate_slot(ctx, (nb - 1) % NS))) return rc;      // getters address the last batch
    ctx->last_n = std::min(ctx->batch, n - (nb - 1) * ctx->batch);
    // device time of the whole call: from the first operation of slot 0 to the end of all streams
    for (int q = 0; q < NS; q++) if (q != (nb - 1) % NS) CK(cudaStreamWaitEvent(ctx->stream, ctx->slots[q].done, 0));
    CK(cudaEventRecord(ctx->ev1, ctx->stream)); CK(cudaEventSynchronize(ctx->ev1));
    { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms; }
    ctx->last_flags = flags;
    return SMC_OK;
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int off = 0; off < n; off += ctx->batch) {
    const int m = std::min(ctx->batch, n - off);
    if ((rc = smc_sample_batch(ctx, first_event_id + (uint64_t)off, m))) return rc;
    if ((rc = run_grid_stages(ctx, m, kinds, nd))) return rc;
    if ((rc = smc_fetch_results(ctx, m))) return rc;
    collect_stage_ms(ctx);
    if ((rc = dsdy_cut_loop(ctx, m, kinds, nd))) return rc;
    if (ctx->ny > 1) { if ((rc = run_slices(ctx, m, kinds, nd, out + (size_t)off * ctx->ny))) return rc; }
    else smc_fill_out(ctx, m, out + off);
    ctx->last_n = m;
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream)); CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms; ctx->last_flags = flags;
  return SMC_OK;
}

// host nuclei -> device for events [off, off+m), then K2 only
int smc_stage_positions(smc_ctx* ctx, int off, int m, const smc_event_in* in, bool any_u, bool any_w) {
  const smc::DevCfg& c = ctx->cfg;
  ctx->epoch++;
  const int A = c.A[0], B = c.A[1], Amax = c.Amax;
  std::vector<double> hw_default;
  std::memset(ctx->h_nuc, 0, (size_t)m * 2 * Amax * smc::NROW * sizeof(double));
  std::memset(ctx->h_hdr_i, 0, (size_t)m * smc::HDR_I * sizeof(int));
  for (int e = 0; e < m; e++) {
    const smc_event_in& ev = in[off + e];
    std::memcpy(ctx->h_nuc + ((size_t)e * 2 + 0) * Amax * smc::NROW, ev.proj, (size_t)A * smc::NROW * sizeof(double));
    std::memcpy(ctx->h_nuc + ((size_t)e * 2 + 1) * Amax * smc::NROW, ev.targ, (size_t)B * smc::NROW * sizeof(double));
    ctx->h_hdr_d[(size_t)e * smc::HDR_D + smc::HD_B] = ev.b;
    ctx->h_hdr_i[(size_t)e * smc::HDR_I + smc::H_GIVENW] = ev.use_given_weights;
    ctx->h_evid[e] = (uint64_t)(off + e); ctx->h_try[e] = 0;
    if (any_u) {
      if (!ev.pair_uniform) FAIL(SMC_ERR_PARAM, "pair_uniform must be given for all events of a call or for none");
      CK(cudaMemcpyAsync(ctx->d_pair_u + (size_t)e * A * B, ev.pair_uniform, (size_t)A * B * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (any_w) {     // rows the caller does not supply (none at all, or fewer than the event's Ncoll) keep weight 1, additional_weight 0
      hw_default.assign((size_t)c.ncoll_cap * 2, 0.0);
      for (int q = 0; q < c.ncoll_cap; q++) hw_default[2 * q] = 1.0;
      if (ev.coll_weight) std::memcpy(hw_default.data(), ev.coll_weight, (size_t)std::max(0, std::min(ev.n_coll_weight, c.ncoll_cap)) * 2 * sizeof(double));
      CK(cudaMemcpy(ctx->d_coll_w + (size_t)e * c.ncoll_cap * 2, hw_default.data(), hw_default.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (ctx->st.nuc_extra) {           // operation-3 state (stale base boxes, quark offsets), else derived from the rows
      std::vector<double> ex((size_t)2 * Amax * smc::NEXTRA, 0.0);
      for (int s = 0; s < 2; s++) {
        const double* rows = s ? ev.targ : ev.proj; const double* given = s ? ev.targ_extra : ev.proj_extra; const int n = s ? B : A;
        for (int i = 0; i < n; i++) {
          double* x = ex.data() + ((size_t)s * Amax + i) * smc::NEXTRA;
          x[smc::XF] = x[smc::XF + 1] = x[smc::XF + 2] = 1.0 / 3.0;                       // Particle::resetFluctFactors
          if (given) std::memcpy(x, given + (size_t)i * smc::NEXTRA, smc::NEXTRA * sizeof(double));
          else { x[smc::XBXL] = rows[i * 8 + 3]; x[smc::XBXR] = rows[i * 8 + 4]; x[smc::XBYL] = rows[i * 8 + 5]; x[smc::XBYR] = rows[i * 8 + 6];
                 x[smc::XCX] = 0.5 * (rows[i * 8 + 3] + rows[i * 8 + 4]); x[smc::XCY] = 0.5 * (rows[i * 8 + 5] + rows[i * 8 + 6]); }
        }
      }
      CK(cudaMemcpy(ctx->st.nuc_extra + (size_t)e * 2 * Amax * smc::NEXTRA, ex.data(), ex.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  CK(cudaMemcpyAsync(ctx->st.nuc, ctx->h_nuc, (size_t)m * 2 * Amax * smc::NROW * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->st.hdr_d, ctx->h_hdr_d, (size_t)m * smc::HDR_D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->st.hdr_i, ctx->h_hdr_i, (size_t)m * smc::HDR_I * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync((void*)ctx->st.event_id, ctx->h_evid, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->st.try_start, ctx->h_try, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ctx->st.pair_u = any_u ? ctx->d_pair_u : nullptr; ctx->st.coll_w = any_w ? ctx->d_coll_w : nullptr; ctx->st.redo = nullptr;
  if (ctx->profile) CK(cudaEventRecord(ctx->pev[0], ctx->stream));
  CK(smc::launch_sample_collide(c, ctx->st, m, true, ctx->stream)); ctx->launches++;
  return SMC_OK;
}

int smc_check_positions(smc_ctx* ctx, int n, const smc_event_in* in, bool* any_u, bool* any_w) {
  const smc::DevCfg& c = ctx->cfg;
  const int A = c.A[0], B = c.A[1];
  *any_u = false; *any_w = false;
  for (int e = 0; e < n; e++) {
    if (in[e].na != A || in[e].nb != B) FAIL(SMC_ERR_PARAM, "smc_event_in.na/nb must equal Aproj/Atarg of the context");
    if (!in[e].proj || !in[e].targ) FAIL(SMC_ERR_PARAM, "smc_event_in.proj/targ is null");
    *any_u |= in[e].pair_uniform != nullptr; *any_w |= in[e].coll_weight != nullptr;
  }
  if (*any_u) {
    const size_t need = (size_t)ctx->batch * A * B * sizeof(double);
    if (need > ctx->pair_u_bytes) { if (ctx->d_pair_u) cudaFree(ctx->d_pair_u); ctx->d_pair_u = nullptr; ctx->pair_u_bytes = 0; CK(cudaMalloc(&ctx->d_pair_u, need)); ctx->pair_u_bytes = need; }
  }
  if (*any_w) {
    const size_t need = (size_t)ctx->batch * c.ncoll_cap * 2 * sizeof(double);
    if (need > ctx->coll_w_bytes) { if (ctx->d_coll_w) cudaFree(ctx->d_coll_w); ctx->d_coll_w = nullptr; ctx->coll_w_bytes = 0; CK(cudaMalloc(&ctx->d_coll_w, need)); ctx->coll_w_bytes = need; }
  }
  return SMC_OK;
}

extern "C" int smc_run_from_positions(smc_ctx* ctx, int n, const smc_event_in* in, unsigned flags, smc_event_out* out) {
  if (!ctx || n < 0 || (n > 0 && (!in || !out))) return SMC_ERR_PARAM;
  if ((flags & ~(unsigned)SMC_RUN_MOMENTS) && n > ctx->batch) FAIL(SMC_ERR_PARAM, "smc_run_from_positions: grids and lists are kept for one device batch, n exceeds smc_max_batch()");
  CK(cudaSetDevice(ctx->device));
  int kinds[8], nd = 0, rc;
  if ((rc = smc_plan_kinds(ctx, flags, kinds, &nd))) return rc;
  bool any_u, any_w;
  if ((rc = smc_check_positions(ctx, n, in, &any_u, &any_w))) return rc;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int off = 0; off < n; off += ctx->batch) {
    const int m = std::min(ctx->batch, n - off);
    if ((rc = smc_stage_positions(ctx, off, m, in, any_u, any_w))) return rc;
    if ((rc = run_grid_stages(ctx, m, kinds, nd))) return rc;
    if ((rc = smc_fetch_results(ctx, m))) return rc;
    collect_stage_ms(ctx);
    if (ctx->ny > 1) { if ((rc = run_slices(ctx, m, kinds, nd, out + (size_t)off * ctx->ny))) return rc; }
    else smc_fill_out(ctx, m, out + off);
    ctx->last_n = m;
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream)); CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms; ctx->last_flags = flags;
  return SMC_OK;
}

// ---- getters -------------------------------------------------------------------------------------
extern "C" int smc_get_grid(smc_ctx* ctx, int slot, int which, double* host) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!host || slot < 0 || slot >= ctx->last_n || which < 0 || which >= SMC_GRID_KINDS) FAIL(SMC_ERR_PARAM, "smc_get_grid: slot outside the last device batch, or bad grid kind");
  static const int map[SMC_GRID_KINDS] = {smc::GK_RHO, smc::GK_TA1, smc::GK_TA2, smc::GK_RHO_BINARY, smc::GK_SPEC_A, smc::GK_SPEC_B};
  const int ks = ctx->st.kind_slot[map[which]];
  if (ks < 0) FAIL(SMC_ERR_STATE, "that grid was not requested in the flags of the last run");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(host, ctx->d_grids + ((size_t)slot * ctx->st.nkinds + ks) * ctx->G, ctx->G * sizeof(double), cudaMemcpyDeviceToHost));
  if (!ctx->need_zero) {       // scan mode: only the event's bounding rectangle was written (all else is 0 by construction)
    int hi[smc::HDR_I];
    CK(cudaMemcpy(hi, ctx->st.hdr_i + (size_t)slot * smc::HDR_I, sizeof hi, cudaMemcpyDeviceToHost));
    const int My = ctx->cfg.Maxy;
    const bool spec = which == SMC_GRID_SPEC_A || which == SMC_GRID_SPEC_B;      // the spectator deposits have a rectangle of their own
    const int rl = hi[spec ? smc::H_SRLO : smc::H_RLO], rh = hi[spec ? smc::H_SRHI : smc::H_RHI], cl = hi[spec ? smc::H_SCLO : smc::H_CLO], ch = hi[spec ? smc::H_SCHI : smc::H_CHI];
    for (int i = 0; i < ctx->cfg.Maxx; i++)
      for (int j = 0; j < My; j++)
        if (i < rl || i >= rh || j < cl || j >= ch) host[(size_t)i * My + j] = 0.0;
  }
  return SMC_OK;
}

// n grids of one kind, slots first_slot .. first_slot+n-1 of the last batch, in ONE strided device->host copy (the
// event-by-event modes fetch a whole batch at once; `host` is best allocated with smc_pinned_alloc)
extern "C" int smc_get_grids(smc_ctx* ctx, int first_slot, int n, int which, double* host) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!host || n < 0 || first_slot < 0 || first_slot + n > ctx->last_n || which < 0 || which >= SMC_GRID_KINDS) FAIL(SMC_ERR_PARAM, "smc_get_grids: slots outside the last device batch, or bad grid kind");
  static const int map[SMC_GRID_KINDS] = {smc::GK_RHO, smc::GK_TA1, smc::GK_TA2, smc::GK_RHO_BINARY, smc::GK_SPEC_A, smc::GK_SPEC_B};
  const int ks = ctx->st.kind_slot[map[which]];
  if (ks < 0) FAIL(SMC_ERR_STATE, "that grid was not requested in the flags of the last run");
  if (!ctx->need_zero) FAIL(SMC_ERR_STATE, "smc_get_grids needs a profile run (SMC_RUN_KEEP_RHO ...): scan mode only writes each event's bounding rectangle");
  if (n == 0) return SMC_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy2DAsync(host, ctx->G * sizeof(double), ctx->d_grids + ((size_t)first_slot * ctx->st.nkinds + ks) * ctx->G, (size_t)ctx->st.nkinds * ctx->G * sizeof(double),
                       ctx->G * sizeof(double), (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SMC_OK;
}
extern "C" void* smc_pinned_alloc(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr; }
extern "C" void smc_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// The list getters serve single events, but callers walk whole batches (operations 1 and 2 write several lists per event):
// the first getter after a run mirrors the batch's records on the host in a few large copies, the others read the mirror.
static int cache_lists(smc_ctx* ctx) {
  smc_list_cache& lc = ctx->lists;
  if (lc.epoch == ctx->epoch && lc.n == ctx->last_n) return SMC_OK;
  const int Amax = ctx->cfg.Amax, n = ctx->last_n;
  CK(cudaSetDevice(ctx->device));
  lc.nuc.resize((size_t)n * 2 * Amax * smc::NROW); lc.ncoll.resize((size_t)n * 2 * Amax); lc.first.resize((size_t)n * Amax); lc.hdr.resize((size_t)n * smc::HDR_I);
  CK(cudaMemcpy(lc.nuc.data(), ctx->st.nuc, lc.nuc.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lc.ncoll.data(), ctx->st.nuc_ncoll, lc.ncoll.size() * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lc.first.data(), ctx->st.nuc_first, lc.first.size() * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lc.hdr.data(), ctx->st.hdr_i, lc.hdr.size() * sizeof(int), cudaMemcpyDeviceToHost));
  int mx = 0;
  for (int e = 0; e < n; e++) mx = std::max(mx, std::min(lc.hdr[(size_t)e * smc::HDR_I + smc::H_NCOLL], ctx->cfg.ncoll_cap));
  lc.coll_stride = mx; lc.coll.resize((size_t)n * mx * smc::CROW); lc.ij.resize((size_t)n * mx);
  if (mx > 0) {
    CK(cudaMemcpy2D(lc.coll.data(), (size_t)mx * smc::CROW * sizeof(double), ctx->st.coll, (size_t)ctx->cfg.ncoll_cap * smc::CROW * sizeof(double), (size_t)mx * smc::CROW * sizeof(double), n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy2D(lc.ij.data(), (size_t)mx * sizeof(int), ctx->st.coll_ij, (size_t)ctx->cfg.ncoll_cap * sizeof(int), (size_t)mx * sizeof(int), n, cudaMemcpyDeviceToHost));
  }
  lc.extra.clear();
  if (ctx->st.nuc_extra) {      // valence-quark state (smc_get_quarks)
    lc.extra.resize((size_t)n * 2 * Amax * smc::NEXTRA);
    CK(cudaMemcpy(lc.extra.data(), ctx->st.nuc_extra, lc.extra.size() * sizeof(double), cudaMemcpyDeviceToHost));
  }
  lc.epoch = ctx->epoch; lc.n = n;
  return SMC_OK;
}
static int fetch_event_lists(smc_ctx* ctx, int slot, std::vector<double>& nuc, std::vector<int>& ncoll, std::vector<int>& first, int hi[smc::HDR_I]) {
  const int Amax = ctx->cfg.Amax; int rc;
  if ((rc = cache_lists(ctx))) return rc;
  const smc_list_cache& lc = ctx->lists;
  nuc.assign(lc.nuc.begin() + (size_t)slot * 2 * Amax * smc::NROW, lc.nuc.begin() + (size_t)(slot + 1) * 2 * Amax * smc::NROW);
  ncoll.assign(lc.ncoll.begin() + (size_t)slot * 2 * Amax, lc.ncoll.begin() + (size_t)(slot + 1) * 2 * Amax);
  first.assign(lc.first.begin() + (size_t)slot * Amax, lc.first.begin() + (size_t)(slot + 1) * Amax);
  std::memcpy(hi, lc.hdr.data() + (size_t)slot * smc::HDR_I, smc::HDR_I * sizeof(int));
  return SMC_OK;
}

extern "C" int smc_get_nucleons(smc_ctx* ctx, int slot, int which, double* host8, int* n) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!n || slot < 0 || slot >= ctx->last_n || which < 0 || which > 1) FAIL(SMC_ERR_PARAM, "smc_get_nucleons: slot outside the last device batch");
  *n = ctx->cfg.A[which];
  if (!host8) return SMC_OK;
  std::vector<double> nuc; std::vector<int> nc, fi; int hi[smc::HDR_I]; int rc;
  if ((rc = fetch_event_lists(ctx, slot, nuc, nc, fi, hi))) return rc;
  for (int i = 0; i < *n; i++) {
    std::memcpy(host8 + (size_t)i * 8, nuc.data() + ((size_t)which * ctx->cfg.Amax + i) * smc::NROW, 8 * sizeof(double));
    host8[(size_t)i * 8 + 2] = (double)nc[(size_t)which * ctx->cfg.Amax + i];     // z slot carries the collision count
  }
  return SMC_OK;
}

// participants in the reference's order: projectile ascending i, target by first hit (Nucleus::markWounded)
extern "C" int smc_get_participants(smc_ctx* ctx, int slot, double* host8, int* n) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!n || slot < 0 || slot >= ctx->last_n) FAIL(SMC_ERR_PARAM, "smc_get_participants: slot outside the last device batch");
  std::vector<double> nuc; std::vector<int> nc, fi; int hi[smc::HDR_I]; int rc;
  if ((rc = fetch_event_lists(ctx, slot, nuc, nc, fi, hi))) return rc;
  *n = hi[smc::H_NP1] + hi[smc::H_NP2];
  if (!host8) return SMC_OK;
  const int Amax = ctx->cfg.Amax; int k = 0;
  auto put = [&](int s, int i) {
    const double* r = nuc.data() + ((size_t)s * Amax + i) * smc::NROW; double* o = host8 + (size_t)(k++) * 8;
    o[0] = r[smc::NX]; o[1] = r[smc::NY]; o[2] = s + 1; o[3] = r[smc::NW]; o[4] = r[smc::NXL]; o[5] = r[smc::NXR]; o[6] = r[smc::NYL]; o[7] = r[smc::NYR];
  };
  for (int i = 0; i < ctx->cfg.A[0]; i++) if (nc[i] > 0) put(0, i);
  std::vector<std::pair<long long, int>> ord;
  for (int j = 0; j < ctx->cfg.A[1]; j++) if (nc[Amax + j] > 0) ord.push_back(std::make_pair((long long)fi[j] * 65536 + j, j));
  std::sort(ord.begin(), ord.end());
  for (auto& pr : ord) put(1, pr.second);
  return SMC_OK;
}

extern "C" int smc_get_collisions(smc_ctx* ctx, int slot, double* host6, int* n) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!n || slot < 0 || slot >= ctx->last_n) FAIL(SMC_ERR_PARAM, "smc_get_collisions: slot outside the last device batch");
  { int rc; if ((rc = cache_lists(ctx))) return rc; }
  const smc_list_cache& lc = ctx->lists;
  const int nc = std::min(lc.hdr[(size_t)slot * smc::HDR_I + smc::H_NCOLL], ctx->cfg.ncoll_cap);
  *n = nc;
  if (!host6 || nc == 0) return SMC_OK;
  const double* c4 = lc.coll.data() + (size_t)slot * lc.coll_stride * smc::CROW; const int* ij = lc.ij.data() + (size_t)slot * lc.coll_stride;
  for (int k = 0; k < nc; k++) {
    double* o = host6 + (size_t)k * 6;
    o[0] = c4[(size_t)k * 4]; o[1] = c4[(size_t)k * 4 + 1]; o[2] = c4[(size_t)k * 4 + 2]; o[3] = c4[(size_t)k * 4 + 3]; o[4] = ij[k] >> 16; o[5] = ij[k] & 0xffff;
  }
  return SMC_OK;
}

// valence quarks of the wounded nucleons, participant order (Nucleus::dumpQuarks, Nucleus.cpp:780-797): x y xL xR yL yR per quark
extern "C" int smc_get_quarks(smc_ctx* ctx, int slot, double* host6, int* n) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!n || slot < 0 || slot >= ctx->last_n) FAIL(SMC_ERR_PARAM, "smc_get_quarks: slot outside the last device batch");
  if (!ctx->st.nuc_extra) FAIL(SMC_ERR_STATE, "smc_get_quarks: run with SMC_RUN_LISTS (or shape_of_entropy 3 / collision_criterion 3) so that the quark state is kept");
  std::vector<double> nuc; std::vector<int> nc, fi; int hi[smc::HDR_I]; int rc;
  if ((rc = fetch_event_lists(ctx, slot, nuc, nc, fi, hi))) return rc;
  *n = 3 * (hi[smc::H_NP1] + hi[smc::H_NP2]);
  if (!host6) return SMC_OK;
  const int Amax = ctx->cfg.Amax;
  if (ctx->lists.extra.size() < (size_t)(slot + 1) * 2 * Amax * smc::NEXTRA) FAIL(SMC_ERR_STATE, "smc_get_quarks: the quark state of this batch was not kept");
  const double* ex = ctx->lists.extra.data() + (size_t)slot * 2 * Amax * smc::NEXTRA;
  int k = 0; const double qw = ctx->p.quark_width;
  auto put = [&](int s, int i) {
    const double* r = nuc.data() + ((size_t)s * Amax + i) * smc::NROW; const double* x = ex + ((size_t)s * Amax + i) * smc::NEXTRA;
    for (int q = 0; q < 3; q++) {
      const double qx = x[smc::XQ + 3 * q], qy = x[smc::XQ + 3 * q + 1], X = qx + r[smc::NX], Y = qy + r[smc::NY];
      double* o = host6 + (size_t)(k++) * 6;
      // Quark::getBoundingBox (Quark.cpp:5-10): the construction-time box (centre = offset, side 8 * quark_width) moved to X, Y
      const double xl = qx - 8 * qw / 2, xr = qx + 8 * qw / 2, yl = qy - 8 * qw / 2, yr = qy + 8 * qw / 2;
      o[0] = X; o[1] = Y; o[2] = xl + (X - qx); o[3] = xr + (X - qx); o[4] = yl + (Y - qy); o[5] = yr + (Y - qy);
    }
  };
  for (int i = 0; i < ctx->cfg.A[0]; i++) if (nc[i] > 0) put(0, i);
  std::vector<std::pair<long long, int>> ord;
  for (int j = 0; j < ctx->cfg.A[1]; j++) if (nc[Amax + j] > 0) ord.push_back(std::make_pair((long long)fi[j] * 65536 + j, j));
  std::sort(ord.begin(), ord.end());
  for (auto& pr : ord) put(1, pr.second);
  return SMC_OK;
}

// spectators: projectile nucleons first (Y>0), then target (MCnucl.cpp:1223-1249)
extern "C" int smc_get_spectators(smc_ctx* ctx, int slot, double* host3, int* n) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!n || slot < 0 || slot >= ctx->last_n) FAIL(SMC_ERR_PARAM, "smc_get_spectators: slot outside the last device batch");
  std::vector<double> nuc; std::vector<int> nc, fi; int hi[smc::HDR_I]; int rc;
  if ((rc = fetch_event_lists(ctx, slot, nuc, nc, fi, hi))) return rc;
  *n = hi[smc::H_NSPEC1] + hi[smc::H_NSPEC2];
  if (!host3) return SMC_OK;
  const double ecm = ctx->p.ecm, vz = std::sqrt(1. - 1. / ((ecm / 2.) * (ecm / 2.)));
  const double Y = 0.5 * std::log((1. + vz) / (1. - vz + 1e-100));
  const int Amax = ctx->cfg.Amax; int k = 0;
  for (int s = 0; s < 2; s++) for (int i = 0; i < ctx->cfg.A[s]; i++) if (nc[(size_t)s * Amax + i] == 0) {
    const double* r = nuc.data() + ((size_t)s * Amax + i) * smc::NROW;
    host3[(size_t)k * 3] = r[smc::NX]; host3[(size_t)k * 3 + 1] = r[smc::NY]; host3[(size_t)k * 3 + 2] = s == 0 ? Y : -Y; k++;
  }
  return SMC_OK;
}

// ---- averaged profiles: filled in by smc_avg.cu ----------------------------------------------------
// ---- centrality sort (scripts/centrality_cut_h5.py:36-110: argsort(-key)) -------------------------
__global__ void iota_kernel(int64_t* p, int64_t n) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
extern "C" int smc_centrality_sort(smc_ctx* ctx, const double* key, int64_t n, int64_t* perm) {
  if (!ctx || !key || !perm || n <= 0 || n > 0x7fffffff) return SMC_ERR_PARAM;
  CK(cudaSetDevice(ctx->device));
  // the five work buffers are kept between calls (the per-centrality wrapper sorts once per key)
  smc_sort_buffers& sb = ctx->sortbuf;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (double*)nullptr, (double*)nullptr, (int64_t*)nullptr, (int64_t*)nullptr, (int)n, 0, 64, ctx->stream);
  if (sb.cap < n || sb.tmp_bytes < tb) {
    cudaFree(sb.k1); cudaFree(sb.k2); cudaFree(sb.v1); cudaFree(sb.v2); cudaFree(sb.tmp);
    sb.k1 = sb.k2 = nullptr; sb.v1 = sb.v2 = nullptr; sb.tmp = nullptr; sb.cap = 0; sb.tmp_bytes = 0;
    CK(cudaMalloc(&sb.k1, n * sizeof(double))); CK(cudaMalloc(&sb.k2, n * sizeof(double)));
    CK(cudaMalloc(&sb.v1, n * sizeof(int64_t))); CK(cudaMalloc(&sb.v2, n * sizeof(int64_t))); CK(cudaMalloc(&sb.tmp, tb));
    sb.cap = n; sb.tmp_bytes = tb;
  }
  CK(cudaMemcpyAsync(sb.k1, key, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sb.v1, n); ctx->launches++;
  size_t tb2 = sb.tmp_bytes;
  cub::DeviceRadixSort::SortPairsDescending(sb.tmp, tb2, sb.k1, sb.k2, sb.v1, sb.v2, (int)n, 0, 64, ctx->stream); ctx->launches++;
  CK(cudaMemcpyAsync(perm, sb.v2, n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SMC_OK;
}

// ---- micro-benchmarks for the roofline denominators -----------------------------------------------
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
    a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}
extern "C" double smc_measure_fp64_peak(smc_ctx* ctx) {
  if (!ctx) return 0.0;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return 0.0;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, ctx->device);
  double* d = nullptr; cudaMalloc(&d, 8);
  const int iters = 20000, blocks = prop.multiProcessorCount * 8, threads = 256;
  fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 100);
  double best = 0;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters);
    cudaEventRecord(ctx->ev1, ctx->stream); cudaEventSynchronize(ctx->ev1);
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    best = std::max(best, tf);
  }
  ctx->launches += 4;
  cudaFree(d);
  return best;
}
__global__ void hbm_write_kernel(double4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double4(1.0, 2.0, 3.0, 4.0);
}
extern "C" double smc_measure_hbm_write_peak(smc_ctx* ctx) {
  if (!ctx) return 0.0;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return 0.0;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, ctx->device);
  const size_t bytes = (size_t)2 << 30; double4* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return 0.0;
  double best = 0;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    hbm_write_kernel<<<prop.multiProcessorCount * 16, 512, 0, ctx->stream>>>(d, bytes / sizeof(double4));
    cudaEventRecord(ctx->ev1, ctx->stream); cudaEventSynchronize(ctx->ev1);
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    if (r) best = std::max(best, bytes / (ms * 1e-3) / 1e9);
  }
  ctx->launches += 4;
  cudaFree(d);
  return best;
}
