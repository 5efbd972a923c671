// smc_ctx.h -- the context object behind the C ABI (shared by smc_api.cu and smc_avg.cu)
#pragma once
#include <string>
#include <vector>
#include "../../include/supermc_b200.h"
#include "smc_common.cuh"

namespace smc {
enum { GK_RHO = 0, GK_TA1 = 1, GK_TA2 = 2, GK_RHO_BINARY = 3, GK_SPEC_A = 4, GK_SPEC_B = 5, GK_RHOA = 6, GK_RHOB = 7 };
cudaError_t launch_sample_collide(const DevCfg&, const Store&, int nev, bool given, cudaStream_t);
cudaError_t launch_deposit(const DevCfg&, const Store&, const int* kinds, int nk, int nev, cudaStream_t);
cudaError_t launch_combine(const DevCfg&, const Store&, int nev, cudaStream_t);
cudaError_t launch_fluctuate(const DevCfg&, const Store&, int nev, cudaStream_t);
cudaError_t launch_moments(const DevCfg&, const Store&, int nev, cudaStream_t);
int deposit_cm_slots(const DevCfg&);
size_t deposit_work_bytes(const DevCfg&, int batch, int nk);
struct KlnCfg { double ecm, lambda, y, dT; int tmax; int pt_order; int npt, nkt, nphi; const double *xp, *wp, *xk, *wk, *cphi;
                int model, maxQ0, maxY, maxKt; double dQ0, siginNN200; const double *rkt, *rna, *ry2; };
cudaError_t launch_kln_table(const KlnCfg&, double* table, cudaStream_t);
}  // namespace smc

// Per-batch device/host buffers of one pipeline slot.  Several slots let sample+collide of batch n+1 run on another
// stream while deposit/moments of batch n are still in flight (smc_run_events rotates over up to SMC_MAX_SLOTS).
struct smc_slot {
  bool ready;
  double* nuc; int* nuc_ncoll; int* nuc_first; double* coll; int* coll_ij; int* part_idx; int* spec_idx;
  int* hdr_i; double* hdr_d; double* mom_out; uint64_t* event_id; int* try_start; double* cm; int* d_redo;
  double* d_grids; size_t grids_bytes; void* d_srcrec; size_t srcrec_bytes; double* d_cmpart;
  double* nuc_extra; double* nuc_extra_tmp;
  cudaStream_t stream; cudaEvent_t done; cudaEvent_t pev[8];
  int* h_hdr_i; double* h_hdr_d; double* h_mom; uint64_t* h_evid; int* h_try;
};

// host mirror of the last batch's event records, filled by the first list getter after a run (smc_api.cu: cache_lists)
struct smc_list_cache { uint64_t epoch; int n; int coll_stride; std::vector<double> nuc, coll, extra; std::vector<int> ncoll, first, hdr, ij; };

struct smc_sort_buffers { double *k1, *k2; int64_t *v1, *v2; void* tmp; int64_t cap; size_t tmp_bytes; };

#define SMC_MAX_SLOTS 4
struct smc_ctx {
  smc_params p; smc_constants k; smc::DevCfg cfg; smc::Store st;
  int device; cudaStream_t stream; cudaEvent_t ev0, ev1;
  int batch; size_t G;
  std::vector<void*> owned;
  double* d_grids; size_t grids_bytes; bool need_zero;
  double* d_cmpart;                                // per deposit CTA partial sums for the centre of mass
  void* d_srcrec; size_t srcrec_bytes;             // expanded deposit sources (smc::SrcRec), [batch][deposit kinds][src_stride]
  double* d_pair_u; size_t pair_u_bytes; double* d_coll_w; size_t coll_w_bytes;
  double* d_quark; double* d_cfgtab[2]; double* d_kln; int* d_redo;
  double* d_rcbk; int rcbk_q, rcbk_y, rcbk_k;      // rcBK uGD tables: kt | N_A | y2, each [q][y][k]
  int* h_hdr_i; double* h_hdr_d; double* h_mom; uint64_t* h_evid; int* h_try; double* h_nuc;
  std::string err; int64_t launches; double last_ms; int last_n; unsigned last_flags;
  // averaged profiles (operation 3)
  int profile; double stage_ms[8]; cudaEvent_t pev[8];
  smc_slot slots[SMC_MAX_SLOTS]; int cur_slot;
  double* d_avg_part; size_t avg_part_bytes;      // per-slice partial sums of one accumulation (smc_avg.cu)
  double* d_avg; int64_t avg_doubles; int64_t avg_count; int avg_from, avg_to, avg_rp, avg_ed;
  void* comm;                                      // multi-GPU state (smc_comm.cu)
  uint64_t epoch; smc_list_cache lists;            // epoch: bumped whenever the device records change
  smc_sort_buffers sortbuf;                        // smc_centrality_sort work space, kept between calls
  bool need_quarks;                                // shape_of_entropy 3 / collision_criterion 3: valence-quark state is live
  int ny, slice;                                   // rapidity slices (MCnucl.cpp:115): slice = the one the grid stages compute next
};

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e); return SMC_ERR_CUDA; } } while (0)
#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)


// helpers of smc_api.cu used by the averaged-profile driver
int smc_plan_kinds(smc_ctx* ctx, unsigned flags, int* kinds, int* nk_dep);
int smc_ensure_extra(smc_ctx* ctx);
int smc_fetch_results(smc_ctx* ctx, int m);
void smc_fill_out(smc_ctx* ctx, int m, smc_event_out* out);
int smc_stage_positions(smc_ctx* ctx, int off, int m, const smc_event_in* in, bool any_u, bool any_w);
int smc_sample_batch(smc_ctx* ctx, uint64_t first_event_id, int m);
int smc_check_positions(smc_ctx* ctx, int n, const smc_event_in* in, bool* any_u, bool* any_w);
int smc_run_grid_stages(smc_ctx* ctx, int m, const int* kinds, int nd);
int smc_activate_slot(smc_ctx* ctx, int s);
int smc_events_first_pass(smc_ctx* ctx, int m, const int* kinds, int nd);
