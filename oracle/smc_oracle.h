/* smc_oracle.h -- CPU restatement ("port") of superMC's per-event hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under supermc_b200/ may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do, and only as the checker.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (or to 1e-13 where libm
 * orderings differ) against the unmodified reference built from /root/reference by
 * oracle/ref_build/Makefile (tests/test_oracle_vs_ref.py, runs where /root/reference exists) and
 * against the committed fixtures tests/golden/ that the same build produced
 * (tests/test_oracle_golden.py, runs everywhere).  Exception: the MC-KLN table integral -- the
 * reference integrates with BASES/VEGAS Monte Carlo (0.1 % stated accuracy); this file restates the
 * integrand exactly and integrates it with a deterministic product rule, so that one function is
 * pinned only to the reference's own Monte-Carlo error (see smc_o_kln_dndy).
 *
 * All arithmetic is IEEE double in the same expression order as the reference; compile with
 * -O2 -ffp-contract=off (no FMA contraction) so that truncating (int) casts and circle masks
 * reproduce the reference's decisions exactly.
 */
#ifndef SMC_ORACLE_H
#define SMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants of the run (Regge96.cpp:27-50, GaussianNucleonsCal.cpp:24-55,130-163) ---- */
double smc_o_sigma_inel(double ecm);
void smc_o_gauss_params(int shape_of_nucleons, double siginNN, double gaussian_lambda,
                        double gauss_nucl_width, double* width, double* sigma_gg);

typedef struct {
  int Maxx, Maxy;
  double Xmin, Ymin, dx, dy;
  double width;        /* nucleon Gaussian width w (== entropy width, quirk Q1) */
  double dsq;          /* 0.1*sigma_in/pi */
  double siginNN;
  double sigma_gg;
  double alpha;
  int shape_of_nucleons, shape_of_entropy, collision_criterion;
  int which_mc_model, sub_model, cc_fluct_model;
} smc_o_cfg;

/* ---- uniform streams ---- */
/* 48-bit LCG of drand48()/srand48() (POSIX), so the port can replay the reference's stream */
typedef struct { uint64_t x; } smc_o_rand48;
void smc_o_srand48(smc_o_rand48* s, long seed);
void smc_o_seed48(smc_o_rand48* s, unsigned short x0, unsigned short x1, unsigned short x2);
double smc_o_drand48(smc_o_rand48* s);
/* Philox4x32-10 (Salmon et al., SC'11) -- the counter-based stream of the CUDA path */
void smc_o_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* The event streams of the CUDA path, restated independently from the layout documented in
 * supermc_b200/csrc/smc_philox.h: key=(seed_lo,seed_hi), ctr=(event lo, event hi,
 * try<<8|kind<<1|nucleus, cand<<12|slot>>1), component slot&1, 53-bit mantissa */
typedef struct { uint32_t seed_lo, seed_hi; uint64_t event; uint32_t tr; int nuc; } smc_o_philox_stream;
double smc_o_uniform_philox(void* st, int kind, int cand, int slot);
/* A uniform source: kind/cand/slot address a counter-based stream; a sequential stream ignores them */
typedef double (*smc_o_uniform_fn)(void* st, int kind, int cand, int slot);
double smc_o_uniform_rand48(void* st, int kind, int cand, int slot);

/* ---- nucleus sampling (Nucleus.cpp:187-317,578-621; Particle.cpp:16-99; Box2D.cpp) ---- */
typedef struct {
  int A;
  double rad, dr, rmaxCut, rwMax;   /* Woods-Saxon parameters (Nucleus.cpp:65-119) */
  double beta2, beta4; int deformed;
  double width, quark_width, quark_R;
  const double* quark_table; int quark_rows;   /* tables/QuarkPos.txt, rows of (r1,r2,cos12) */
} smc_o_nucleus;
void smc_o_nucleus_init(smc_o_nucleus* n, int A, int deformed, double width, double quark_width,
                        const double* quark_table, int quark_rows);
/* out: A rows of 7 doubles (x,y,z,xL,xR,yL,yR), sorted by xL like Nucleus.cpp:314; returns #uniforms */
long smc_o_populate(const smc_o_nucleus* n, double xCenter, double yCenter,
                    smc_o_uniform_fn U, void* st, double* out7, double* cx_phi);
/* table-driven nuclei (Nucleus.cpp:555-574, 623-666): cfg = 3A coordinates of one configuration */
long smc_o_populate_table(const smc_o_nucleus* n, const double* cfg, int recentre, int redraw_rotation,
                          double xCenter, double yCenter, smc_o_uniform_fn U, void* st, double* out7);

/* deuteron (Nucleus.cpp:203-209,342-359; HulthenFunc.cpp:27-41) */
double smc_o_hulthen_inv_cdf(double y);
long smc_o_populate_deuteron(const smc_o_nucleus* n, double xCenter, double yCenter, smc_o_uniform_fn U, void* st, double* out7);

/* ---- collisions (MCnucl.cpp:217-308,357-385; GaussianNucleonsCal.cpp:59-67) ---- */
/* proj7/targ7: rows (x,y,z,xL,xR,yL,yR) sorted by xL.  u_dense (A*B, may be NULL): receives the
 * uniform consumed by pair (i,j) in sweep order, -1 where the sweep never tested the pair.
 * pairs: out (i,j) per hit in sweep order, capacity max_pairs.  returns Ncoll */
int smc_o_collide(const smc_o_cfg* c, int A, const double* proj7, int B, const double* targ7,
                  smc_o_uniform_fn U, void* st, const double* u_in_dense, double* u_dense,
                  int* ncollA, int* ncollB, int* firsthitB, int* pairs, int max_pairs, long* n_tested);

/* ---- deposits ---- */
/* sources: rows of 8 doubles (x, y, xL, xR, yL, yR, weight, extra) */
void smc_o_thickness(const smc_o_cfg* c, int n, const double* src8, double* TA);            /* MCnucl.cpp:432-478 */
void smc_o_add_density(const smc_o_cfg* c, int n, const double* src8, double* dens);        /* MCnucl.cpp:822-866 */
/* shape_of_entropy = 3 / collision_criterion = 3 (valence-quark substructure) */
void smc_o_quark_out(double* qout6);          /* the populate functions leave qx0 qy0 qx1 qy1 qx2 qy2 per nucleon (sorted order) here; NULL = off */
void smc_o_quark_collide(const double* qA6, const double* qB6, double quark_width);      /* inputs of the quark-overlap hit test */
void smc_o_add_density_quarks(const smc_o_cfg* c, int n, const double* src8, const double* q6, const double* f3, double qw, double* dens);
double smc_o_density_quarks(const smc_o_cfg* c, int np, const double* proj8, const double* qP, const double* fP, int nt, const double* targ8,
                            const double* qT, const double* fT, int nc, const double* coll8, double qw, double* rho);
void smc_o_binary_term(const smc_o_cfg* c, int n, const double* coll8, double* tab);         /* MCnucl.cpp:724-759 */
void smc_o_unit_gauss(const smc_o_cfg* c, int n, const double* src8, double* grid);          /* MCnucl.cpp:481-531,534-614 */
/* rho for which_mc_model 5 / 7 (MCnucl.cpp:688-811); returns dndy (sum over cells) */
double smc_o_density(const smc_o_cfg* c, int np, const double* proj8, int nt, const double* targ8,
                     int nc, const double* coll8, double* rho);
/* which_mc_model 1 (MCnucl.cpp:654-687, arsenal.cpp:33-54); returns dndy or -1 on table overflow */
double smc_o_six_point(double x, double y, double v00, double v01, double v02, double v10, double v11, double v20);
double smc_o_density_kln(const smc_o_cfg* c, const double* TA1, const double* TA2, const double* table,
                         int tmax, double dT, double* rho);

/* ---- moments ---- */
/* GlueDensity.cpp:87-144: out = {xcm, ycm, angle, weight} */
void smc_o_cm_angle(const smc_o_cfg* c, const double* dens, int n, double* out4);
/* MakeDensity.cpp:2244-2511; boxes: rows (xL,xR,yL,yR) in getHotSpots order (MCnucl.cpp:1303-1323).
 * out: mom_real[10], mom_imag[10], momp_real[10], momp_imag[10], rn[10], then total, xc, yc  (53) */
void smc_o_eccentricities(const smc_o_cfg* c, const double* dens, int nbox, const double* boxes4,
                          int from_order, int to_order, double* out53);

/* ---- MC-KLN (KLNModel.cpp:97-125,219-277,360-399; KLNfunc.h:14-17) ---- */
typedef struct { double ecm, lambda, siginNN200; int model; int pt_order; } smc_o_kln;
double smc_o_kln_integrand(const smc_o_kln* k, double y, double ta, double tb, const double x[3]);
double smc_o_kln_dndy(const smc_o_kln* k, double y, double ta, double tb, int npt, int nkt, int nphi);
/* rcBK tabulated uGD (rcBKfunc.h:65-121); set = 100 (59 tables, dQ0 0.1) or 101 (30 tables, dQ0 0.168) */
typedef struct { int set, maxQ0, maxY, maxKt; double dQ0; const double *kt, *na, *y2; } smc_o_rcbk;
void smc_o_spline_natural(const double* x, const double* y, int n, double* y2);
double smc_o_rcbk_func(const smc_o_rcbk* t, double qs0_2, double x, double kt2, double alp);
double smc_o_rcbk_integrand(const smc_o_kln* k, const smc_o_rcbk* t, double y, double ta, double tb, const double x[3]);
double smc_o_rcbk_dndy(const smc_o_kln* k, const smc_o_rcbk* t, double y, double ta, double tb, int npt, int nkt, int nphi);

/* ---- NBD multiplicity fluctuations (MCnucl.cpp:868-905, NBD.cpp, RandomVariable.cpp:190-286) ---- */
double smc_o_nbd_pdf(double p, double r, double k_in);
int smc_o_nbd_envelope(double p, double r, double* edge, double* height);
long smc_o_nbd_rand(double p, double r, smc_o_rand48* st);
int smc_o_nbd_law(double p, double r, long* k0, double* weights, int cap);
long smc_o_nbd_quantile(double p, double r, double u);
void smc_o_fluctuate_density(const smc_o_cfg* c, int model, double cc_k, const double* TA1, const double* TA2,
                             const double* u, double* rho);

/* 3-D extension (scripts/generate_3d_profiles/profile_3d.cpp:274-325); src7 rows x y id eta sigma_x sigma_y sigma_eta */
void smc_o_profile3d(int nx, int ny, int neta, double dx, double dy, double deta, int n, const double* src7, double* rho);

#ifdef __cplusplus
}
#endif
#endif
