/* supermc_b200.h -- C ABI of the B200-native superMC hot path.
 *
 * The reference (chunshen1987/superMC) has no FFI/plugin layer: the seam the per-event hot path
 * sits behind is the C++ class MCnucl as driven by MakeDensity (reference src/MCnucl.h:87-146,
 * src/MakeDensity.cpp:2143-2224).  This header is the batch-of-events C ABI that replaces that seam:
 * every entry point names the reference interface it stands in for.  Plain pointers and sizes only,
 * status-code returns (the reference prints to cerr and calls exit()), no exceptions cross the ABI.
 *
 * Threading: one context per GPU, used from one host thread at a time.  The caller owns every host
 * buffer; the context owns all device memory.  There is NO CPU fallback: every compute entry point
 * returns SMC_ERR_CUDA if no sm_100-class device is usable.
 */
#ifndef SUPERMC_B200_H
#define SUPERMC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SMC_ABI_VERSION 3
#define SMC_EXTRA_ROW 20

enum {
  SMC_OK = 0,
  SMC_ERR_PARAM = 1,      /* bad argument / unsupported option (reference: cerr + exit) */
  SMC_ERR_CUDA = 2,       /* CUDA runtime failure or no device */
  SMC_ERR_STATE = 3,      /* call sequence error (e.g. KLN density without a table) */
  SMC_ERR_OVERFLOW = 4,   /* a per-event capacity was exceeded (collision list, KLN table range) */
  SMC_ERR_NOMEM = 5
};

/* Flat copy of every parameters.dat key the path consumes (reference parameters.dat; consumers listed
 * in SURVEY.md appendix A).  Names are the reference's, lower-cased as ParameterReader stores them
 * (src/ParameterReader.cpp:61-69).  Integers are the reference's doubles truncated at use. */
typedef struct smc_params {
  int which_mc_model;            /* 1 MC-KLN, 5 MC-Glauber, 7 sqrt(TA TB)      MCnucl.cpp:86 */
  int sub_model;                 /* Glb: 1 classic, 2 "Uli"; KLN: 7, rcBK 100/101  MCnucl.cpp:87, ParamDefs.h */
  double lambda;                 /* KLN saturation-scale exponent              MakeDensity.cpp:97 */
  int tmax, tmax_subdivision;    /* KLN table size                             MCnucl.cpp:30,917-918 */
  double alpha;                  /* WN/BC mixing                               MCnucl.cpp:27 */
  int aproj, atarg;              /*                                            MCnucl.cpp:99-106 */
  int proj_deformed, targ_deformed;
  int include_nn_correlation;    /*                                            Nucleus.cpp:26 */
  int shape_of_nucleons;         /* 1 disk, 2 gaussian(sigma_NN), 3 gaussian(gaussian_lambda), 4 user width  GaussianNucleonsCal.cpp:28-55 */
  int collision_criterion;       /* 1 disk, 2 gaussian, 3 valence-quark overlap, else from shape_of_entropy   MCnucl.cpp:357-385 */
  int shape_of_entropy;          /* 1 disk, 2 gaussian, 3 valence quarks       MCnucl.cpp:109,856 */
  double quark_width;            /*                                            Nucleus.cpp:29 */
  double gauss_nucl_width;       /* shape_of_nucleons == 4 */
  double ecm, bmin, bmax;        /*                                            MakeDensity.cpp:34-35,61 */
  int npmin, npmax;              /* inclusive Npart window                     MCnucl.cpp:388-393 */
  int cutdsdy; double cutdsdy_lowerbound, cutdsdy_upperbound;  /*              MakeDensity.cpp:39-41 */
  int64_t randomseed;            /* Philox key (the reference seeds drand48/rand/mt19937 with it) */
  double finalfactor;            /*                                            MakeDensity.cpp:30 */
  int ecc_from_order, ecc_to_order;  /*                                        MakeDensity.cpp:2112 */
  double maxx, maxy, dx, dy;     /* grid                                       MCnucl.cpp:33-40 */
  int cc_fluctuation_model;      /* 0 none, 1/2 NBD per cell (constant k / k from TA,TB), 6 Gamma weights   MCnucl.cpp:67-83,868-905 */
  double cc_fluctuation_gamma_theta;
  int pt_order;                  /* KLN pT weight, 1 unless PT_flag<0          MCnucl.cpp:52-55 */
  double gaussian_lambda;        /* shape_of_nucleons == 3                     GaussianNucleonsCal.cpp:29,39-44 */
  double cc_fluctuation_k;       /* NBD k of cc_fluctuation_model == 1         MCnucl.cpp:68,872-880 */
  int ny; double ymax;           /* rapidity slices: rapMin = -ymax, rapMax = ymax  MakeDensity.cpp:54-58, MCnucl.cpp:33-38 */
  /* capacities of the device-side event records (not reference parameters) */
  int max_batch;                 /* events resident per launch wave; 0 = default */
  int ncoll_cap;                 /* collision-list capacity per event; 0 = default */
} smc_params;

/* Derived run constants (what the MCnucl / GaussianNucleonsCal constructors compute). */
typedef struct smc_constants {
  double siginnn;                /* sigma_in(ecm) [mb]            Regge96.cpp:27-50, MCnucl.cpp:58-64 */
  double siginnn200;             /* sigma_in(200 GeV)                                                  */
  double width;                  /* nucleon gaussian width w      GaussianNucleonsCal.cpp:33-54         */
  double sigma_gg;               /* Newton solution               GaussianNucleonsCal.cpp:130-163       */
  double dsq;                    /* 0.1 sigma_in / pi             MCnucl.cpp:120                        */
  int maxx_cells, maxy_cells;    /* Maxx, Maxy                    MCnucl.cpp:39-40                      */
  double kln_dt; int kln_tmax;   /* table step / size             MCnucl.cpp:915-919                    */
} smc_constants;

/* One event as the reference's MakeDensity loops see it after dumpEccentricities
 * (src/MakeDensity.cpp:2163-2193, 2244-2500). */
typedef struct smc_event_out {
  double b;                      /* impact parameter */
  int npart1, npart2, ncoll;     /* MCnucl::getNpart1/2, getNcoll    MCnucl.h:101-103 */
  int tries;                     /* (b, nuclei) draws consumed by the rejection loop MakeDensity.cpp:2147-2162 */
  int nspec;                     /* MCnucl::getSpectators count      MCnucl.cpp:1223-1249 */
  int status;                    /* SMC_OK or SMC_ERR_OVERFLOW for this event */
  double dsdy;                   /* sum(rho) dx dy, no finalFactor   MakeDensity.cpp:2557-2567 */
  double total;                  /* sum(rho*finalFactor) dx dy       column 48 of *_ecc_eccp_10.dat */
  double xc, yc;                 /* centre of mass of the profile    MakeDensity.cpp:2273-2282 */
  double mom[9][5];              /* n=1..9: Re eps_n, Im eps_n, Re eps'_n, Im eps'_n, <r^n>  :2389-2430 */
  double rn0;                    /* <r^0> (always 1; kept so rn[0..9] is complete) */
  int nonzero_cells;             /* cells with rho != 0 (work measure for the roofline; not a reference output) */
  int reserved;
} smc_event_out;

/* Parity / replay input: nuclei supplied by the caller instead of being sampled
 * (stands in for MCnucl::generateNuclei output, src/MCnucl.cpp:208-214). */
typedef struct smc_event_in {
  double b;
  int na, nb;
  const double* proj;            /* na rows of 8: x y z xL xR yL yR weight ; sorted by xL (Nucleus.cpp:314) */
  const double* targ;            /* nb rows of 8 */
  const double* pair_uniform;    /* na*nb uniforms consumed by GaussianNucleonsCal::testSmoothCollision
                                    (GaussianNucleonsCal.cpp:59-67), row-major (i,j); NULL = draw from Philox */
  const double* coll_weight;     /* optional: Gamma weight per collision in (i,j)-sorted order, ncoll rows of 2
                                    (fluctfactor, additional_weight); NULL = draw / derive on device */
  int n_coll_weight;
  int use_given_weights;         /* 1: nucleon weights come from proj/targ column 7 (reference: last draw wins) */
  /* per-nucleon state beyond the 8-double row, rows of SMC_EXTRA_ROW (20): [0..3] stale base box xL xR yL yR
   * (Particle::baseBox, quirk Q4), [4..12] three valence-quark offsets (x y z each), [13..14] AABB centre x y,
   * [15..17] per-quark multiplicity weights (Quark::fluctFactor; shape_of_entropy 3, used with use_given_weights),
   * [18..19] spare.  Read by the averaged profiles (smc_avg_run_from_positions) and by the quark-substructure options
   * (shape_of_entropy 3, collision_criterion 3); NULL = derive (no quark offsets, weights 1/3) */
  const double* proj_extra;
  const double* targ_extra;
} smc_event_in;

/* what a run materialises besides the smc_event_out rows */
enum {
  SMC_RUN_MOMENTS = 1u,          /* eccentricity table columns (operation 9) */
  SMC_RUN_KEEP_RHO = 2u,         /* keep rho (entropy) grids on the device for smc_get_grid (operations 1,2) */
  SMC_RUN_THICKNESS = 4u,        /* TA1/TA2 grids (always on for MC-KLN) */
  SMC_RUN_RHO_BINARY = 8u,       /* MCnucl::calculate_rho_binary           MCnucl.cpp:481-531 */
  SMC_RUN_SPECTATORS = 16u,      /* MCnucl::calculate_spectator_density    MCnucl.cpp:534-614 */
  SMC_RUN_LISTS = 32u            /* keep participant / collision / spectator lists for the dump* writers */
};
enum { SMC_GRID_RHO = 0, SMC_GRID_TA1 = 1, SMC_GRID_TA2 = 2, SMC_GRID_RHO_BINARY = 3,
       SMC_GRID_SPEC_A = 4, SMC_GRID_SPEC_B = 5, SMC_GRID_KINDS = 6 };

typedef struct smc_ctx smc_ctx;

int  smc_abi_version(void);
/* defaults of the reference's parameters.dat */
void smc_params_default(smc_params* p);
/* MCnucl::MCnucl / ~MCnucl (src/MCnucl.cpp:22-201) */
int  smc_create(const smc_params* p, int device, smc_ctx** out);
void smc_destroy(smc_ctx* ctx);
/* replaces `cerr << ...; exit()` */
const char* smc_last_error(const smc_ctx* ctx);
int  smc_get_constants(const smc_ctx* ctx, smc_constants* c);
/* events resident per device batch: the grid / list getters address the last batch of a run, so a caller that wants
 * them asks for at most this many events per smc_run_events call */
int  smc_max_batch(const smc_ctx* ctx);
/* srand(seed); srand48(seed) of src/main.cpp:32-33 after the fact: replaces the Philox key (all ranks of a run share one) */
int  smc_set_seed(smc_ctx* ctx, int64_t seed);

/* Nucleus::Nucleus reading tables/QuarkPos.txt into Particle::quark_pos (src/Nucleus.cpp:37-48):
 * n rows of (r1, r2, cos theta12).  Not loaded => r1 = r2 = 0 (every AABB is the +-4w base box). */
int  smc_load_quark_table(smc_ctx* ctx, const double* rows3, int n);
/* Nucleus::readin_helium3/4/carbon/oxygen_position and readin_nucleon_positions
 * (src/Nucleus.cpp:383-522): n_cfg configurations of A nucleons (x,y,z), which = 0 proj / 1 targ */
int  smc_load_config_table(smc_ctx* ctx, int which, const double* xyz, int n_cfg, int a);

/* MCnucl::makeTable (src/MCnucl.cpp:911-960): builds the tmax^2 dN/dy(TA,TB) table on the device with
 * a deterministic quadrature of KLNModel::func (src/KLNModel.cpp:219-277), one table per rapidity slice at
 * y = -ymax + 2 ymax / ny * iy; host_out (ny*tmax*tmax) optional */
int  smc_build_kln_table(smc_ctx* ctx, double* host_out);
/* rcBKfunc::rcBKfunc (src/rcBKfunc.cpp:15-210), sub_model 100 (59 files) / 101 (30 files): kt and N_A columns of the
 * javier/ft_rcbk_mv_qs02_*.dat files as [maxq0][maxy][maxkt]; needed before smc_build_kln_table for those sub-models */
int  smc_load_rcbk_tables(smc_ctx* ctx, const double* kt, const double* na, int maxq0, int maxy, int maxkt);
/* install a table computed elsewhere (e.g. the reference's data/dNdyTable.dat); ny tables of tmax*tmax, slice-major */
int  smc_set_kln_table(smc_ctx* ctx, const double* table, int tmax, double dt);

/* generateNuclei -> getBinaryCollision -> CentralityCut -> calculateThickness -> setDensity ->
 * dumpEccentricities for n accepted events with global ids first_event_id .. first_event_id+n-1
 * (src/MakeDensity.cpp:2143-2224).  Event k draws from Philox key=randomseed, counter=(k, try, ...),
 * so the result set does not depend on batch size or GPU count.
 * ny > 1: `out` receives n*ny rows, out[e*ny + iy] = event e at rapidity slice iy (the reference appends one table row
 * per slice, MakeDensity.cpp:2170-2193); the grids left for the getters are those of the last slice, which is what the
 * reference's files hold (every slice is written to the same file name). */
int  smc_run_events(smc_ctx* ctx, uint64_t first_event_id, int n, unsigned flags, smc_event_out* out);
/* same, nuclei supplied (parity entry; reference: the loop body after generateNuclei) */
int  smc_run_from_positions(smc_ctx* ctx, int n, const smc_event_in* in, unsigned flags, smc_event_out* out);

/* MCnucl::getRho/getTA1/getTA2/get_rho_binary/get_spectator_density (src/MCnucl.h:94-99,142) for the
 * event in batch slot `slot` of the last run; host receives Maxx*Maxy doubles, row index = ix */
int  smc_get_grid(smc_ctx* ctx, int slot, int which, double* host);
/* the same for n consecutive slots in one strided copy (operations 1 and 2 fetch a batch at once); host receives
 * n * Maxx*Maxy doubles.  smc_pinned_alloc returns page-locked host memory for it (device->host copies at full PCIe rate). */
int  smc_get_grids(smc_ctx* ctx, int first_slot, int n, int which, double* host);
void* smc_pinned_alloc(size_t bytes);
void smc_pinned_free(void* p);
/* MCnucl::dumpparticipantTable / dumpBinaryTable / dumpSpectatorsTable payloads (src/MCnucl.cpp:1177-1269)
 * rows: participants (x, y, nucleus id, weight, xL, xR, yL, yR) ; collisions (x, y, weight, addw, i, j) ;
 * spectators (x, y, rapidity).  Returns the row count through *n; host may be NULL to query. */
int  smc_get_participants(smc_ctx* ctx, int slot, double* host8, int* n);
int  smc_get_collisions(smc_ctx* ctx, int slot, double* host6, int* n);
int  smc_get_spectators(smc_ctx* ctx, int slot, double* host3, int* n);
int  smc_get_nucleons(smc_ctx* ctx, int slot, int which, double* host8, int* n);
/* Nucleus::dumpQuarks (src/Nucleus.cpp:780-797, the data/quarks.data rows): the three valence quarks of every wounded
 * nucleon, participant order, rows x y xL xR yL yR.  Needs SMC_RUN_LISTS, shape_of_entropy 3 or collision_criterion 3. */
int  smc_get_quarks(smc_ctx* ctx, int slot, double* host6, int* n);

/* MakeDensity::generate_profile_average accumulators (src/MakeDensity.cpp:1240-1577).
 * slots: for each order in [from,to] and each variant (0 rotated, 1 reaction-plane) and each quantity
 * (see SMC_AVG_*), a Maxx*Maxy sum on the device; smc_avg_run adds `n` accepted events. */
enum { SMC_AVG_SD = 0, SMC_AVG_TATB = 1, SMC_AVG_RHO_BINARY = 2, SMC_AVG_TA = 3, SMC_AVG_TB = 4,
       SMC_AVG_SPEC_A = 5, SMC_AVG_SPEC_B = 6, SMC_AVG_QUANTITIES = 7 };
/* branches: bit 0 = entropy (use_sd), bit 1 = energy (use_ed) accumulators */
int  smc_avg_begin(smc_ctx* ctx, int from_order, int to_order, int with_rp, int branches);
int  smc_avg_run(smc_ctx* ctx, uint64_t first_event_id, int n, smc_event_out* out);
int  smc_avg_run_from_positions(smc_ctx* ctx, int n, const smc_event_in* in, smc_event_out* out);
/* device address + element count of the accumulator block and of the accepted-event counter, so the
 * caller can sum them across GPUs (ncclAllReduce / torch.distributed.all_reduce on the raw pointer) */
int  smc_avg_device_buffer(smc_ctx* ctx, void** dev_ptr, int64_t* n_doubles);
int  smc_avg_count(smc_ctx* ctx, int64_t* count);
int  smc_avg_set_count(smc_ctx* ctx, int64_t count);
/* mean = sum / count for (order, variant, quantity, branch); host receives Maxx*Maxy doubles */
int  smc_avg_get(smc_ctx* ctx, int order, int variant, int quantity, int branch, double* host);

/* ---- several GPUs of one node, one process (rank) per GPU -------------------------------------------------------
 * The reference only knows "start 8 copies with different seeds" (CollectDataAccordingToSettings.py:110-115).  Here the
 * ranks own contiguous global event-id ranges of one run (smc_run_events takes the first id), and two things are combined:
 * the operation-3 accumulators (MakeDensity.cpp:1299 running means, kept as sums) and per-event rows for the centrality
 * sort.  smc_comm_init: rank 0 listens on addr:port (dotted IPv4, e.g. MASTER_ADDR / MASTER_PORT + 1), the others connect;
 * the all-reduce then runs as ncclAllReduce over NVLink (libnccl.so.2 resolved at run time) or, when ranks share a GPU or
 * NCCL is absent, as a peer-memory kernel over CUDA IPC.  SMC_COMM_BACKEND=nccl|ipc forces one.  All waits time out. */
int  smc_comm_init(smc_ctx* ctx, int rank, int world, const char* addr, int port);
void smc_comm_finalize(smc_ctx* ctx);                     /* also done by smc_destroy */
const char* smc_comm_backend(const smc_ctx* ctx);         /* "nccl", "ipc", "single", or "none" before smc_comm_init */
int  smc_comm_barrier(smc_ctx* ctx);
int  smc_comm_bcast_i64(smc_ctx* ctx, int64_t* v);        /* rank 0's value to all (seed of randomSeed < 0, src/main.cpp:28-31) */
/* rank 0 receives the rows of all ranks in rank order (`all` holds `cap` doubles); n_per_rank[world] optional */
int  smc_comm_gather_doubles(smc_ctx* ctx, const double* mine, int64_t n, double* all, int64_t cap, int64_t* n_per_rank);
/* in-place sum over all ranks of the accumulator block and of the accepted-event counter of smc_avg_* */
int  smc_avg_allreduce(smc_ctx* ctx);
double smc_comm_last_allreduce_ms(const smc_ctx* ctx);    /* device time of the last smc_avg_allreduce */

/* scripts/centrality_cut_h5.py:36-110: sort n events descending by key; perm receives the order */
int  smc_centrality_sort(smc_ctx* ctx, const double* key, int64_t n, int64_t* perm);

/* ---- the reference's 3-D extension (scripts/generate_3d_profiles/profile_3d.cpp, main.cpp) -----------------------
 * n sources (x, y, id: 1 projectile / 2 target, as in ParticipantTable_event_<k>.dat) become Gaussians in (eta_s, x, y):
 * rho_out[neta][nx][ny].  random_flag as profile_3d::set_variables (0: eta = +-2; 1: eta drawn from the tabulated
 * distribution; 2, 3: widths fluctuate too).  eta_in / sigma3_in (n and 3n doubles: sigma_x sigma_y sigma_eta) override the
 * draws (parity entry; the reference seeds with time()); eta_used / sigma3_used report what was used.  No context needed. */
typedef struct smc_profile3d_params { int nx, ny, neta; double dx, dy, deta; double ecm; int random_flag; int64_t seed; } smc_profile3d_params;
int  smc_profile3d(int device, const smc_profile3d_params* p, int n, const double* x, const double* y, const int* id,
                   const double* eta_in, const double* sigma3_in, double* rho_out, double* eta_used, double* sigma3_used);

/* diagnostics: kernels launched by this context so far, device time of the last run [ms];
 * with profiling on, CUDA-event time per stage {sample+collide, deposit, combine, moments} accumulates */
int     smc_set_profiling(smc_ctx* ctx, int on);
int     smc_get_stage_ms(const smc_ctx* ctx, double* ms4);
int64_t smc_kernel_launches(const smc_ctx* ctx);
double  smc_last_run_ms(const smc_ctx* ctx);
/* FP64 FMA throughput micro-benchmark on the context's device [TFLOP/s]; the roofline denominator
 * for the deposit / moment kernels (the bound is the FP64 pipe, not HBM) */
double  smc_measure_fp64_peak(smc_ctx* ctx);
double  smc_measure_hbm_write_peak(smc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
