/* smc_oracle.c -- CPU restatement of superMC's per-event hot path.  See smc_oracle.h.
 * TEST INFRASTRUCTURE ONLY: the product (supermc_b200/) never links or calls this file.
 * Each function cites the reference lines (/root/reference/src/...) it follows.
 */
#include "smc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ======================================================================================
 * sigma_NN: PDG-1996 Regge fit, pp channel (Regge96.cpp:27-50; call sites MCnucl.cpp:58-64)
 * ====================================================================================== */
double smc_o_sigma_inel(double ecm) {
  double s = ecm * ecm;
  double sig = 22.0 * pow(s, 0.079) + 56.1 * pow(s, -0.46);
  double bel = 2.0 * 2.3 + 2.0 * 2.3 + 4.0 * pow(s, 0.0808) - 4.2;
  double sigel = 0.0511 * sig * sig / bel;
  return sig - sigel;
}

/* 38-point Gauss-Legendre rule on [a,b] (Nucleus.cpp:697-751) */
static void gauss38(double a, double b, double* xn, double* wn) {
  static const double xh[19] = {4.078514790458e-2, 1.220840253379e-1, 2.025704538921e-1, 2.817088097902e-1,
    3.589724404794e-1, 4.338471694324e-1, 5.058347179279e-1, 5.744560210478e-1, 6.392544158297e-1,
    6.997986803792e-1, 7.556859037540e-1, 8.065441676053e-1, 8.520350219324e-1, 8.918557390046e-1,
    9.257413320486e-1, 9.534663309335e-1, 9.748463285902e-1, 9.897394542664e-1, 9.980499305357e-1};
  static const double wh[19] = {8.152502928039e-2, 8.098249377060e-2, 7.990103324353e-2, 7.828784465821e-2,
    7.615366354845e-2, 7.351269258474e-2, 7.038250706690e-2, 6.678393797914e-2, 6.274093339213e-2,
    5.828039914700e-2, 5.343201991033e-2, 4.822806186076e-2, 4.270315850467e-2, 3.689408159400e-2,
    3.083950054518e-2, 2.457973973823e-2, 1.815657770961e-2, 1.161344471647e-2, 5.002880749632e-3};
  double x[38], w[38];
  for (int k = 0; k < 19; k++) { x[19 + k] = xh[k]; w[19 + k] = wh[k]; }
  for (int i = 0; i < 19; i++) { x[i] = -x[37 - i]; w[i] = w[37 - i]; }
  for (int i = 0; i < 38; i++) { xn[i] = (b - a) * x[i] / 2.0 + (a + b) / 2.0; wn[i] = (b - a) * w[i] / 2.0; }
}

/* sigma_gg from sigma_in = int d^2b [1-exp(-sigma_gg Tpp(b))] by Newton (GaussianNucleonsCal.cpp:130-163) */
static double sig_eff(double siginNN, double width) {
  double xg[38], wg[38];
  gauss38(0.0, 1.0, xg, wg);
  double sigin = siginNN * 0.1, Bmax = 5.0 * width, sigeff = 10.0, sigeff0;
  do {
    sigeff0 = sigeff;
    double sum = 0.0, dN = 0.0;
    for (int ib = 0; ib < 38; ib++) {
      double b = xg[ib] * Bmax, db = wg[ib] * Bmax;
      double Tpp = exp(-b * b / (4. * width * width)) / (M_PI * (4. * width * width));
      sum += 2 * M_PI * b * db * (1.0 - exp(-sigeff * Tpp));
      dN += 2 * M_PI * b * db * Tpp * exp(-sigeff * Tpp);
    }
    sigeff -= (sum - sigin) / dN;
  } while (fabs(sigeff - sigeff0) > 1e-4);
  return sigeff;
}

/* arsenal.cpp:531-571 qiu_simpsons applied to Gamma0Integrand (GaussianNucleonsCal.cpp:19-22): Simpson sums on
 * 1, 2, 4, ... panels until two successive values differ by less than epsilon (depth cap 50) */
static double gamma0_integrand(double t) { return 1. / t * exp(-t); }
static double qiu_simpsons_gamma0(double a, double b, double epsilon) {
  double f_1 = gamma0_integrand(a) + gamma0_integrand(b), f_2 = 0., f_4 = 0., sum_previous = 0., sum_current = 0.;
  long count = 1, i;
  double length = (b - a), step = length / count;
  int currentRecursionDepth = 1;
  f_4 = gamma0_integrand(a + 0.5 * step);
  sum_current = (length / 6) * (f_1 + f_2 * 2. + f_4 * 4.);
  do {
    sum_previous = sum_current;
    f_2 += f_4;
    count *= 2;
    step /= 2.0;
    f_4 = 0.;
    for (i = 0; i < count; i++) f_4 += gamma0_integrand(a + step * (i + 0.5));
    sum_current = (length / 6 / count) * (f_1 + f_2 * 2. + f_4 * 4.);
    if (currentRecursionDepth > 50) break;
    else currentRecursionDepth++;
  } while (fabs(sum_current - sum_previous) > epsilon);
  return sum_current;
}

/* GaussianNucleonsCal ctor (GaussianNucleonsCal.cpp:24-55) */
void smc_o_gauss_params(int shape, double siginNN, double gaussian_lambda, double gauss_nucl_width,
                        double* width, double* sigma_gg) {
  double w = 0.0, sg = 0.0;
  if (shape == 1) { w = sqrt(0.1 * siginNN / (M_PI)) / 2.0; sg = sig_eff(siginNN, w); }
  if (shape == 3) {                                                                      /* :39-44 */
    double ratio = (0.5772156649 + qiu_simpsons_gamma0(gaussian_lambda, gaussian_lambda + 100., 1e-10) + log(gaussian_lambda)) / gaussian_lambda;
    w = sqrt(siginNN * 0.1 / (4 * M_PI * gaussian_lambda * ratio));
    sg = siginNN * 0.1 / ratio;
  }
  else if (shape == 2) { w = sqrt(0.1 * siginNN / M_PI) / sqrt(8); sg = sig_eff(siginNN, w); }
  else if (shape == 4) { w = gauss_nucl_width; sg = sig_eff(siginNN, w); }
  *width = w; *sigma_gg = sg;
}

/* ======================================================================================
 * uniform streams
 * ====================================================================================== */
void smc_o_srand48(smc_o_rand48* s, long seed) { s->x = (((uint64_t)(uint32_t)seed) << 16) | 0x330EULL; }
void smc_o_seed48(smc_o_rand48* s, unsigned short x0, unsigned short x1, unsigned short x2) {
  s->x = (uint64_t)x0 | ((uint64_t)x1 << 16) | ((uint64_t)x2 << 32);
}
double smc_o_drand48(smc_o_rand48* s) {
  s->x = (0x5DEECE66DULL * s->x + 0xBULL) & 0xFFFFFFFFFFFFULL;
  return ldexp((double)s->x, -48);
}
double smc_o_uniform_rand48(void* st, int kind, int cand, int slot) {
  (void)kind; (void)cand; (void)slot;
  return smc_o_drand48((smc_o_rand48*)st);
}

void smc_o_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double smc_o_uniform_philox(void* st, int kind, int cand, int slot) {
  const smc_o_philox_stream* s = (const smc_o_philox_stream*)st;
  uint32_t ctr[4] = {(uint32_t)s->event, (uint32_t)(s->event >> 32),
                     (s->tr << 8) | ((uint32_t)kind << 1) | (uint32_t)s->nuc, ((uint32_t)cand << 12) | ((uint32_t)slot >> 1)};
  uint32_t key[2] = {s->seed_lo, s->seed_hi}, o[4];
  smc_o_philox4x32_10(ctr, key, o);
  uint32_t a = o[2 * (slot & 1)], b = o[2 * (slot & 1) + 1];
  uint64_t m = ((uint64_t)a << 21) | ((uint64_t)b >> 11);
  return (double)m * (1.0 / 9007199254740992.0);
}

/* ======================================================================================
 * Box2D semantics (Box2D.cpp:14-41, Box2D.h:49-63) on a 6-double box {xL,xR,yL,yR,xC,yC}
 * ====================================================================================== */
typedef struct { double xL, xR, yL, yR, xC, yC; } box_t;
static void box_zero(box_t* b) { b->xL = b->xR = b->yL = b->yR = b->xC = b->yC = 0; }
static void box_set_center(box_t* b, double x, double y) {
  b->xL += x - b->xC; b->xR += x - b->xC; b->yL += y - b->yC; b->yR += y - b->yC; b->xC = x; b->yC = y;
}
static void box_set_square(box_t* b, double size) {
  b->xL = b->xC - size / 2; b->xR = b->xC + size / 2; b->yL = b->yC - size / 2; b->yR = b->yC + size / 2;
}
static void box_union(box_t* b, const box_t* o) {
  b->xL = o->xL < b->xL ? o->xL : b->xL;  b->xR = o->xR > b->xR ? o->xR : b->xR;
  b->yL = o->yL < b->yL ? o->yL : b->yL;  b->yR = o->yR > b->yR ? o->yR : b->yR;
  b->xC = (b->xL + b->xR) / 2.0; b->yC = (b->yL + b->yR) / 2.0;
}

/* Point3D::rotate (MathBasics.cpp:41-50) */
static void rot3(double cth, double phi, double* x, double* y, double* z) {
  double x0 = *x, y0 = *y, z0 = *z, cphi = cos(phi), sth = sqrt(1. - cth * cth), sphi = sin(phi);
  *x = cth * cphi * x0 - sphi * y0 + sth * cphi * z0;
  *y = cth * sphi * x0 + cphi * y0 + sth * sphi * z0;
  *z = -sth * x0 + 0. * y0 + cth * z0;
}

/* Particle ctor: base box + three valence quarks -> AABB (Particle.cpp:16-24,32-99; Quark.h:28-35;
 * Quark.cpp:5-10).  Consumes 4 uniforms (kind 4, slots 0..3). */
/* valence-quark state for shape_of_entropy = 3 / collision_criterion = 3 (set by the test, not thread-safe):
 * qout  -- emit_sorted() leaves qx0 qy0 qx1 qy1 qx2 qy2 per nucleon (sorted order) here
 * qA/qB -- offsets of the projectile / target nucleons for the quark-overlap hit test, qw = quark_width */
static struct { double* qout; const double* qA; const double* qB; double qw; } g_q;
void smc_o_quark_out(double* qout6) { g_q.qout = qout6; }
void smc_o_quark_collide(const double* qA6, const double* qB6, double quark_width) { g_q.qA = qA6; g_q.qB = qB6; g_q.qw = quark_width; }
static double g_last_q[6];

static void particle_box(const smc_o_nucleus* n, double x0, double y0, smc_o_uniform_fn U, void* st,
                         int cand, box_t* out) {
  box_t base; box_zero(&base); box_set_center(&base, x0, y0); box_set_square(&base, 8 * n->width);
  double u0 = U(st, 4, cand, 0);
  int index = (int)(250000 * u0);
  double r1 = 0, r2 = 0, z12 = 0;
  if (n->quark_table && index < n->quark_rows) {
    r1 = n->quark_table[3 * index]; r2 = n->quark_table[3 * index + 1]; z12 = n->quark_table[3 * index + 2];
  }
  r1 = r1 * n->quark_R; r2 = r2 * n->quark_R;
  double Theta12 = acos(z12);
  double z1 = 2. * U(st, 4, cand, 1) - 1.;
  double Theta1 = acos(z1);
  double phi1 = 2 * M_PI * U(st, 4, cand, 2);
  double phi2 = 2 * M_PI * U(st, 4, cand, 3);
  double ux = sin(Theta1) * cos(phi1), uy = sin(Theta1) * sin(phi1), uz = z1;
  double vx = sin(Theta1 + Theta12) * cos(phi1), vy = sin(Theta1 + Theta12) * sin(phi1), vz = cos(Theta1 + Theta12);
  double c = cos(phi2), s = sin(phi2);
  double r1x = r1 * ux, r1y = r1 * uy;
  double r2x = vx * (c + ux * ux * (1 - c)) + vy * (ux * uy * (1 - c) - uz * s) + vz * (ux * uz * (1 - c) + uy * s);
  double r2y = vx * (ux * uy * (1 - c) + uz * s) + vy * (c + uy * uy * (1 - c)) + vz * (uy * uz * (1 - c) - ux * s);
  r2x = r2x * r2; r2y = r2y * r2;
  double qx[3] = {r1x, r2x, -r1x - r2x}, qy[3] = {r1y, r2y, -r1y - r2y};
  *out = base;
  for (int q = 0; q < 3; q++) {
    box_t b; box_zero(&b); box_set_center(&b, qx[q], qy[q]); box_set_square(&b, 8 * n->quark_width);
    box_set_center(&b, x0 + qx[q], y0 + qy[q]);
    box_union(out, &b);
    g_last_q[2 * q] = qx[q]; g_last_q[2 * q + 1] = qy[q];
  }
}

void smc_o_nucleus_init(smc_o_nucleus* n, int A, int deformed, double width, double quark_width,
                        const double* quark_table, int quark_rows) {
  memset(n, 0, sizeof *n);
  n->A = A; n->deformed = deformed; n->width = width; n->quark_width = quark_width;
  n->quark_R = sqrt((3.0 / 2.0) * (width * width - quark_width * quark_width));   /* Nucleus.cpp:31-32 */
  n->quark_table = quark_table; n->quark_rows = quark_rows;
  if (A == 1) return;
  double a = (double)A;
  n->rad = 1.12 * pow(a, 0.333333) - 0.86 / pow(a, 0.333333); n->dr = 0.54;       /* Nucleus.cpp:65-66 */
  if (A == 197) { n->rad = 6.42; n->dr = 0.45; }
  else if (A == 63) { n->rad = 4.28; n->dr = 0.5; }
  else if (A == 238) { n->rad = 6.86; n->dr = 0.44; }
  else if (A == 208) { n->rad = 6.67; n->dr = 0.44; }
  else if (A == 129) { n->rad = 5.36; n->dr = 0.590; }
  n->rmaxCut = n->rad + 2.5;
  n->rwMax = 1.0 / (1.0 + exp(-n->rad / n->dr));
  if (deformed) {                                                                   /* Nucleus.cpp:122-144 */
    if (A == 197) { n->beta2 = -0.13; n->beta4 = -0.03; }
    else if (A == 63) { n->beta2 = 0.162; n->beta4 = 0.006; }
    else if (A == 129) { n->beta2 = 0.162; n->beta4 = -0.003; }
    else if (A == 238) { n->beta2 = 0.28; n->beta4 = 0.093; }
  }
}

static double sph_harm(int l, double ct) {                                          /* Nucleus.cpp:669-694 */
  if (l == 2) return (3.0 * ct * ct - 1.0) * 0.31539156525252005;
  double y = 35.0 * ct * ct * ct * ct; y -= 30.0 * ct * ct; y += 3.0;
  return y * 0.10578554691520431;
}

typedef struct { double x, y, z; box_t box; int idx; double q[6]; } part_t;
static int cmp_xl(const void* a, const void* b) {
  const part_t* p = (const part_t*)a; const part_t* q = (const part_t*)b;
  if (p->box.xL < q->box.xL) return -1;
  if (p->box.xL > q->box.xL) return 1;
  return p->idx - q->idx;
}
static void part_setxy(part_t* p, double x, double y) {                              /* Particle.cpp:176-185 */
  p->x = x; box_set_center(&p->box, x, p->box.yC);
  p->y = y; box_set_center(&p->box, p->box.xC, y);
}
static void emit_sorted(part_t* P, int A, double* out7) {
  qsort(P, A, sizeof(part_t), cmp_xl);                                              /* Nucleus.cpp:314 */
  for (int i = 0; i < A; i++) {
    double* o = out7 + 7 * i;
    o[0] = P[i].x; o[1] = P[i].y; o[2] = P[i].z; o[3] = P[i].box.xL; o[4] = P[i].box.xR; o[5] = P[i].box.yL; o[6] = P[i].box.yR;
    if (g_q.qout) memcpy(g_q.qout + 6 * i, P[i].q, sizeof P[i].q);
  }
}

/* Nucleus::populate for A==1 and the Woods-Saxon branch (Nucleus.cpp:187-202,272-317,578-621) */
long smc_o_populate(const smc_o_nucleus* n, double xCenter, double yCenter,
                    smc_o_uniform_fn U, void* st, double* out7, double* cx_phi) {
  long nu = 0;
  const int A = n->A;
  const double rmin = 0.9 * 0.9;
  double ctr = 1.0 - 2.0 * U(st, 1, 0, 0);
  double phir = 2 * M_PI * U(st, 1, 0, 1);
  nu += 2;
  if (cx_phi) { cx_phi[0] = ctr; cx_phi[1] = phir; }
  part_t* P = (part_t*)malloc(sizeof(part_t) * (A > 0 ? A : 1));
  int cand = 0;
  int nws = 0;   /* flat index of the Woods-Saxon rejection draws: the address the Philox callback uses (drand48 ignores it) */
  if (A == 1) {
    P[0].x = xCenter; P[0].y = yCenter; P[0].z = 0.0; P[0].idx = 0;
    particle_box(n, xCenter, yCenter, U, st, cand, &P[0].box); memcpy(P[0].q, g_last_q, sizeof g_last_q); nu += 4;
    emit_sorted(P, 1, out7); free(P); return nu;
  }
  double xcm = 0.0, ycm = 0.0, zcm = 0.0;
  for (int ia = 0; ia < A; ia++) {
    double x, y, z; int icon;
    do {
      double r = 0.0, cx = 1.0;
      if (n->deformed) {
        double rad1, rwMax1;
        do {
          r = n->rmaxCut * pow(U(st, 2, nws, 0), 1.0 / 3.0);
          cx = 1.0 - 2.0 * U(st, 2, nws, 1);
          double y20 = sph_harm(2, cx), y40 = sph_harm(4, cx);
          rad1 = n->rad * (1.0 + n->beta2 * y20 + n->beta4 * y40);
          rwMax1 = 1.0 / (1.0 + exp(-rad1 / n->dr));
          nu += 3;
        } while (U(st, 2, nws++, 2) * rwMax1 > 1.0 / (1.0 + exp((r - rad1) / n->dr)));
        double sx = sqrt(1.0 - cx * cx);
        double phi = 2 * M_PI * U(st, 3, cand, 1); nu += 1;
        x = r * sx * cos(phi); y = r * sx * sin(phi); z = r * cx;
        rot3(ctr, phir, &x, &y, &z);
      } else {
        do {
          r = n->rmaxCut * pow(U(st, 2, nws, 0), 1.0 / 3.0);
          nu += 2;
        } while (U(st, 2, nws++, 1) * n->rwMax > 1.0 / (1.0 + exp((r - n->rad) / n->dr)));
        cx = 1.0 - 2.0 * U(st, 3, cand, 0);
        double sx = sqrt(1.0 - cx * cx);
        double phi = 2 * M_PI * U(st, 3, cand, 1); nu += 2;
        x = r * sx * cos(phi); y = r * sx * sin(phi); z = r * cx;
      }
      icon = 0;
      for (int i = 0; i < ia; i++) {
        double r2 = (x - P[i].x) * (x - P[i].x) + (y - P[i].y) * (y - P[i].y) + (z - P[i].z) * (z - P[i].z);
        if (r2 < rmin) { icon = 1; break; }
      }
      if (icon) cand++;
    } while (icon == 1);
    xcm += x; ycm += y; zcm += z;
    P[ia].x = x; P[ia].y = y; P[ia].z = z; P[ia].idx = ia;
    particle_box(n, x, y, U, st, cand, &P[ia].box); memcpy(P[ia].q, g_last_q, sizeof g_last_q); nu += 4;
    cand++;
  }
  for (int ia = 0; ia < A; ia++) {
    double x = P[ia].x - xcm / A + xCenter, y = P[ia].y - ycm / A + yCenter, z = P[ia].z - zcm / A;
    part_setxy(&P[ia], x, y); P[ia].z = z;
  }
  emit_sorted(P, A, out7);
  free(P);
  return nu;
}

/* table-driven nuclei: He3/He4/C/O (GetNucleonPosition, Nucleus.cpp:555-574: rotate by the populate()
 * orientation, no recentring) and NN-correlated Au/Pb (Nucleus.cpp:239-271,623-666: recentre, draw a
 * fresh rotation, recentre again in populate). */
long smc_o_populate_table(const smc_o_nucleus* n, const double* cfg, int recentre, int redraw_rotation,
                          double xCenter, double yCenter, smc_o_uniform_fn U, void* st, double* out7) {
  long nu = 0; const int A = n->A;
  double ctr = 1.0 - 2.0 * U(st, 1, 0, 0);
  double phir = 2 * M_PI * U(st, 1, 0, 1); nu += 2;
  part_t* P = (part_t*)malloc(sizeof(part_t) * A);
  double* t = (double*)malloc(sizeof(double) * 3 * A);
  memcpy(t, cfg, sizeof(double) * 3 * A);
  if (recentre) {
    double xcm = 0, ycm = 0, zcm = 0;
    for (int ia = 0; ia < A; ia++) { xcm += t[3 * ia]; ycm += t[3 * ia + 1]; zcm += t[3 * ia + 2]; }
    for (int ia = 0; ia < A; ia++) { t[3 * ia] -= xcm / A; t[3 * ia + 1] -= ycm / A; t[3 * ia + 2] -= zcm / A; }
  }
  if (redraw_rotation) { ctr = 1.0 - 2.0 * U(st, 1, 0, 2); phir = 2 * M_PI * U(st, 1, 0, 3); nu += 2; }
  for (int ia = 0; ia < A; ia++) rot3(ctr, phir, &t[3 * ia], &t[3 * ia + 1], &t[3 * ia + 2]);
  if (recentre) {
    double xcm = 0, ycm = 0, zcm = 0;
    for (int ia = 0; ia < A; ia++) {
      xcm += t[3 * ia]; ycm += t[3 * ia + 1]; zcm += t[3 * ia + 2];
      P[ia].x = t[3 * ia]; P[ia].y = t[3 * ia + 1]; P[ia].z = t[3 * ia + 2]; P[ia].idx = ia;
      particle_box(n, P[ia].x, P[ia].y, U, st, ia, &P[ia].box); memcpy(P[ia].q, g_last_q, sizeof g_last_q); nu += 4;
    }
    for (int ia = 0; ia < A; ia++) {
      double x = P[ia].x - xcm / A + xCenter, y = P[ia].y - ycm / A + yCenter, z = P[ia].z - zcm / A;
      part_setxy(&P[ia], x, y); P[ia].z = z;
    }
  } else {
    for (int ia = 0; ia < A; ia++) {
      P[ia].x = t[3 * ia] + xCenter; P[ia].y = t[3 * ia + 1] + yCenter; P[ia].z = t[3 * ia + 2]; P[ia].idx = ia;
      particle_box(n, P[ia].x, P[ia].y, U, st, ia, &P[ia].box); memcpy(P[ia].q, g_last_q, sizeof g_last_q); nu += 4;
    }
  }
  emit_sorted(P, A, out7);
  free(P); free(t);
  return nu;
}

/* deuteron: Hulthen wave function, inverse CDF by the reference's Newton iteration with a numeric
 * derivative (HulthenFunc.cpp:27-41, arsenal.cpp invertFunc :318-380, RandomVariable.cpp:70-76,140-145;
 * Nucleus.cpp:203-209,342-359) */
static double hulthen_cdf(double r) {
  const double alpha = .228, beta = 1.18;
  if (r <= 0) return 0;
  const double c = (alpha * beta * (alpha + beta)) / ((alpha - beta) * (alpha - beta));
  return 2 * c * (2 * (exp(-r * (alpha + beta)) / (alpha + beta)) - .5 * exp(-2 * alpha * r) / alpha - .5 * exp(-2 * beta * r) / beta + .5 / alpha + .5 / beta - 2 / (alpha + beta));
}
double smc_o_hulthen_inv_cdf(double y) {
  const double xL = 0, xR = 100.0, dx = 0.001, accuracy = dx * 0.001;
  double XX2 = 1.0, XX1 = XX2 - 10 * accuracy; int impatience = 0;
  while (fabs(XX2 - XX1) > accuracy) {
    XX1 = XX2;
    double F0 = hulthen_cdf(XX1) - y;
    double X1 = (XX1 > xL + dx) ? XX1 - dx : xL, X2 = (XX1 < xR - dx) ? XX1 + dx : xR;
    double F3 = (hulthen_cdf(X1) - hulthen_cdf(X2)) / (X1 - X2);
    XX2 = XX1 - F0 / F3;
    if (++impatience > 60) break;
  }
  return XX2;
}
long smc_o_populate_deuteron(const smc_o_nucleus* n, double xCenter, double yCenter, smc_o_uniform_fn U, void* st, double* out7) {
  double ctr = 1.0 - 2.0 * U(st, 1, 0, 0), phir = 2 * M_PI * U(st, 1, 0, 1);
  double u = U(st, 9, 0, 0);
  double d = smc_o_hulthen_inv_cdf(0.0 + 1e-30 + (1.0 - 2 * 1e-30) * u);
  double x1 = d / 2.0, y1 = 0, z1 = 0;
  rot3(ctr, phir, &x1, &y1, &z1);
  part_t P[2];
  P[0].x = x1 + xCenter; P[0].y = y1 + yCenter; P[0].z = z1; P[0].idx = 0;
  P[1].x = -x1 + xCenter; P[1].y = -y1 + yCenter; P[1].z = -z1; P[1].idx = 1;
  particle_box(n, P[0].x, P[0].y, U, st, 0, &P[0].box); memcpy(P[0].q, g_last_q, sizeof g_last_q);
  particle_box(n, P[1].x, P[1].y, U, st, 1, &P[1].box); memcpy(P[1].q, g_last_q, sizeof g_last_q);
  emit_sorted(P, 2, out7);
  return 11;
}

/* ======================================================================================
 * collisions: AABB sweep + hit test (MCnucl.cpp:242-295,357-385; GaussianNucleonsCal.cpp:59-67)
 * ====================================================================================== */
int smc_o_collide(const smc_o_cfg* c, int A, const double* P, int B, const double* T,
                  smc_o_uniform_fn U, void* st, const double* u_in, double* u_dense,
                  int* ncollA, int* ncollB, int* firsthitB, int* pairs, int max_pairs, long* n_tested) {
  int crit = c->collision_criterion;
  if (crit < 1 || crit > 4) crit = (c->shape_of_entropy == 2) ? 2 : (c->shape_of_entropy == 3) ? 3 : 1;   /* MCnucl.cpp:376-383 */
  if (crit == 4 || (crit == 3 && !(g_q.qA && g_q.qB))) return -1;                    /* 4 (numeric overlap) is not restated */
  const double w = c->width;
  for (int i = 0; i < A; i++) ncollA[i] = 0;
  for (int j = 0; j < B; j++) { ncollB[j] = 0; if (firsthitB) firsthitB[j] = -1; }
  if (u_dense) for (long k = 0; k < (long)A * B; k++) u_dense[k] = -1.0;
  int ncoll = 0, start = 0; long tested = 0;
  for (int ip = 0; ip < A; ip++) {
    const double* p = P + 7 * ip;
    double tXL = 0, tXR = 0;
    while (start < B) {
      tXL = T[7 * start + 3]; tXR = T[7 * start + 4];
      if (tXR >= p[3]) break;
      start++;
    }
    int i = start;
    while (i < B && p[4] >= tXL) {
      const double* t = T + 7 * i;
      tXL = t[3];
      if (p[5] <= t[6] && p[6] >= t[5]) {
        double b = sqrt((t[0] - p[0]) * (t[0] - p[0]) + (t[1] - p[1]) * (t[1] - p[1]));
        int hit;
        tested++;
        if (crit == 1) hit = (b * b <= c->dsq) ? 1 : 0;
        else if (crit == 3) {                                                            /* GaussianNucleonsCal.cpp:70-97 */
          const double gw2 = g_q.qw * g_q.qw; double overlap = 0;
          for (int a = 0; a < 3; a++) for (int q = 0; q < 3; q++) {
            const double mx = g_q.qA[6 * ip + 2 * a] + p[0], my = g_q.qA[6 * ip + 2 * a + 1] + p[1];       /* Quark::getX = x + parent */
            const double yx = g_q.qB[6 * i + 2 * q] + t[0], yy = g_q.qB[6 * i + 2 * q + 1] + t[1];
            const double d = (mx - yx) * (mx - yx) + (my - yy) * (my - yy);
            overlap += (1 / (4 * M_PI * gw2)) * exp(-d / (4 * gw2)) / 9;
          }
          double u = u_in ? u_in[(long)ip * B + i] : U(st, 5, ip, i);
          if (u_dense) u_dense[(long)ip * B + i] = u;
          hit = (u < 1. - exp(-c->sigma_gg * overlap)) ? 1 : 0;
        } else {
          double u = u_in ? u_in[(long)ip * B + i] : U(st, 5, ip, i);
          if (u_dense) u_dense[(long)ip * B + i] = u;
          hit = (u < 1. - exp(-c->sigma_gg * exp(-b * b / (4. * w * w)) / (4. * M_PI * w * w))) ? 1 : 0;
        }
        if (hit) {
          if (ncoll < max_pairs) { pairs[2 * ncoll] = ip; pairs[2 * ncoll + 1] = i; }
          if (firsthitB && ncollB[i] == 0) firsthitB[i] = ncoll;
          ncollA[ip]++; ncollB[i]++; ncoll++;
        }
      }
      i++;
    }
  }
  if (n_tested) *n_tested = tested;
  return ncoll;
}

/* ======================================================================================
 * deposits.  Window rule (quirk Q7): left=(int)((x-d-Xmin)/dx), right=(int)((x+d-Xmin)/dx), clip to
 * [0,Max], loop left <= i < right.
 * ====================================================================================== */
static double d_max_of(const smc_o_cfg* c) {
  return (c->shape_of_nucleons == 1) ? 2. * sqrt(c->dsq) : 5. * c->width;
}
static int imax2(int a, int b) { return a > b ? a : b; }
static int imin2(int a, int b) { return a < b ? a : b; }

/* Particle::getSmoothTn (Particle.cpp:122-132) */
static double smooth_tn(double w, double x, double y, double xg, double yg) {
  double r = sqrt((xg - x) * (xg - x) + (yg - y) * (yg - y));
  if (r > 5 * w) return 0;
  return (1 / (2 * M_PI * w * w)) * exp(-r * r / (2 * w * w));
}

void smc_o_thickness(const smc_o_cfg* c, int n, const double* S, double* TA) {     /* MCnucl.cpp:432-478 */
  const double d_max = d_max_of(c);
  for (int k = 0; k < n; k++) {
    double x = S[8 * k], y = S[8 * k + 1];
    int xl = imax2(0, (int)((x - d_max - c->Xmin) / c->dx)), xr = imin2(c->Maxx, (int)((x + d_max - c->Xmin) / c->dx));
    int yl = imax2(0, (int)((y - d_max - c->Ymin) / c->dy)), yr = imin2(c->Maxy, (int)((y + d_max - c->Ymin) / c->dy));
    for (int ix = xl; ix < xr; ix++) {
      double xg = c->Xmin + ix * c->dx;
      for (int iy = yl; iy < yr; iy++) {
        double yg = c->Ymin + iy * c->dy;
        double dc = (x - xg) * (x - xg) + (y - yg) * (y - yg);
        if (c->shape_of_nucleons == 1) { if (dc > c->dsq) continue; TA[(long)ix * c->Maxy + iy] += 10.0 / c->siginNN; }
        else TA[(long)ix * c->Maxy + iy] += smooth_tn(c->width, x, y, xg, yg);
      }
    }
  }
}

void smc_o_add_density(const smc_o_cfg* c, int n, const double* S, double* dens) { /* MCnucl.cpp:822-866 */
  for (int k = 0; k < n; k++) {
    const double* s = S + 8 * k;
    double x = s[0], y = s[1], f = s[6];
    int xl = imax2(0, (int)((s[2] - c->Xmin) / c->dx)), xr = imin2(c->Maxx, (int)((s[3] - c->Xmin) / c->dx));
    int yl = imax2(0, (int)((s[4] - c->Ymin) / c->dy)), yr = imin2(c->Maxy, (int)((s[5] - c->Ymin) / c->dy));
    for (int ir = xl; ir < xr; ir++) {
      double xg = c->Xmin + ir * c->dx;
      for (int jr = yl; jr < yr; jr++) {
        double yg = c->Ymin + jr * c->dy;
        double dc = (x - xg) * (x - xg) + (y - yg) * (y - yg);
        if (c->shape_of_entropy == 1) { if (dc > c->dsq) continue; double areai = 10.0 / c->siginNN; dens[(long)ir * c->Maxy + jr] += areai * f; }
        else dens[(long)ir * c->Maxy + jr] += smooth_tn(c->width, x, y, xg, yg) * f;   /* Particle.cpp:168-174 */
      }
    }
  }
}

/* addDensity with shape_of_entropy == 3 (MCnucl.cpp:856-858 -> Particle::getFluctuatedDensity, Particle.cpp:149-163 ->
 * Quark::getSmoothDensity / getSmoothTn, Quark.h:52-55, Quark.cpp:14-22): three Gaussians of width quark_width at the
 * valence quarks, weights f3, cut at d > 5*quark_width where d is the SQUARED distance (the reference compares fm^2 with
 * fm -- kept); window = the nucleon's AABB */
void smc_o_add_density_quarks(const smc_o_cfg* c, int n, const double* S, const double* q6, const double* f3, double qw, double* dens) {
  for (int k = 0; k < n; k++) {
    const double* s = S + 8 * k;
    double X = s[0], Y = s[1];
    int xl = imax2(0, (int)((s[2] - c->Xmin) / c->dx)), xr = imin2(c->Maxx, (int)((s[3] - c->Xmin) / c->dx));
    int yl = imax2(0, (int)((s[4] - c->Ymin) / c->dy)), yr = imin2(c->Maxy, (int)((s[5] - c->Ymin) / c->dy));
    for (int ir = xl; ir < xr; ir++) {
      double xg = c->Xmin + ir * c->dx;
      for (int jr = yl; jr < yr; jr++) {
        double yg = c->Ymin + jr * c->dy;
        double density = 0;
        for (int q = 0; q < 3; q++) {
          double xRel = xg - X, yRel = yg - Y, x = q6[6 * k + 2 * q], y = q6[6 * k + 2 * q + 1];
          double d = (xRel - x) * (xRel - x) + (yRel - y) * (yRel - y);
          double tn = (d > 5 * qw) ? 0 : (1 / (2 * M_PI * qw * qw)) * exp(-d / (2 * qw * qw));
          density += f3[3 * k + q] * tn;
        }
        dens[(long)ir * c->Maxy + jr] += density;
      }
    }
  }
}

double smc_o_density_quarks(const smc_o_cfg* c, int np, const double* proj8, const double* qP, const double* fP, int nt, const double* targ8,
                            const double* qT, const double* fT, int nc, const double* coll8, double qw, double* rho) {
  const long G = (long)c->Maxx * c->Maxy;
  double dndy = 0.0;
  double* a = (double*)calloc(G, sizeof(double));
  double* b = (double*)calloc(G, sizeof(double));
  if (c->which_mc_model == 5) {
    if (c->sub_model == 1) {
      smc_o_add_density_quarks(c, np, proj8, qP, fP, qw, a); smc_o_add_density_quarks(c, nt, targ8, qT, fT, qw, a);
      double prefactor = (1.0 - c->alpha) / 2.;
      for (long k = 0; k < G; k++) a[k] = a[k] * prefactor;
    }
    if (c->alpha > 1e-8) smc_o_binary_term(c, nc, coll8, b);
    for (long k = 0; k < G; k++) { double d = a[k] + b[k]; rho[k] = d; dndy += d; }
  } else {
    smc_o_add_density_quarks(c, np, proj8, qP, fP, qw, a); smc_o_add_density_quarks(c, nt, targ8, qT, fT, qw, b);
    for (long k = 0; k < G; k++) { double d = sqrt(a[k] * b[k]); rho[k] = d; dndy += d; }
  }
  free(a); free(b);
  return dndy;
}

void smc_o_binary_term(const smc_o_cfg* c, int n, const double* S, double* tab) {  /* MCnucl.cpp:724-759 */
  const double d_max = d_max_of(c), wsq = c->width * c->width, dcmax = 25. * wsq, Alpha = c->alpha;
  for (int k = 0; k < n; k++) {
    const double* s = S + 8 * k;
    double x = s[0], y = s[1], fluct = (c->cc_fluct_model > 5) ? s[6] : 1.0, addw = s[7];
    int xl = imax2(0, (int)((x - d_max - c->Xmin) / c->dx)), xr = imin2(c->Maxx, (int)((x + d_max - c->Xmin) / c->dx));
    int yl = imax2(0, (int)((y - d_max - c->Ymin) / c->dy)), yr = imin2(c->Maxy, (int)((y + d_max - c->Ymin) / c->dy));
    for (int ir = xl; ir < xr; ir++) {
      double xg = c->Xmin + ir * c->dx;
      for (int jr = yl; jr < yr; jr++) {
        double yg = c->Ymin + jr * c->dy;
        double dc = (x - xg) * (x - xg) + (y - yg) * (y - yg);
        if (c->shape_of_entropy == 1) {
          if (dc <= c->dsq) tab[(long)ir * c->Maxy + jr] += fluct * (10.0 / c->siginNN) * (Alpha + (1. - Alpha) * addw);
        } else if (dc <= dcmax) {
          tab[(long)ir * c->Maxy + jr] += fluct * (1 / (2 * M_PI * c->width * c->width)) * exp(-dc / (2 * wsq)) * (Alpha + (1. - Alpha) * addw);
        }
      }
    }
  }
}

/* rho_binary and the spectator densities: unit-weight deposits (MCnucl.cpp:481-531,534-614) */
void smc_o_unit_gauss(const smc_o_cfg* c, int n, const double* S, double* grid) {
  const double d_max = d_max_of(c), wsq = c->width * c->width, dcmax = 25. * wsq;
  for (int k = 0; k < n; k++) {
    double x = S[8 * k], y = S[8 * k + 1];
    int xl = imax2(0, (int)((x - d_max - c->Xmin) / c->dx)), xr = imin2(c->Maxx, (int)((x + d_max - c->Xmin) / c->dx));
    int yl = imax2(0, (int)((y - d_max - c->Ymin) / c->dy)), yr = imin2(c->Maxy, (int)((y + d_max - c->Ymin) / c->dy));
    for (int ir = xl; ir < xr; ir++) {
      double xg = c->Xmin + ir * c->dx;
      for (int jr = yl; jr < yr; jr++) {
        double yg = c->Ymin + jr * c->dy;
        double dc = (x - xg) * (x - xg) + (y - yg) * (y - yg);
        if (c->shape_of_nucleons == 1) { if (dc <= c->dsq) grid[(long)ir * c->Maxy + jr] += (10.0 / c->siginNN); }
        else { if (dc > dcmax) continue; grid[(long)ir * c->Maxy + jr] += (1 / (2 * M_PI * c->width * c->width)) * exp(-dc / (2 * wsq)); }
      }
    }
  }
}

double smc_o_density(const smc_o_cfg* c, int np, const double* proj8, int nt, const double* targ8,
                     int nc, const double* coll8, double* rho) {
  const long G = (long)c->Maxx * c->Maxy;
  double dndy = 0.0;
  double* a = (double*)calloc(G, sizeof(double));
  double* b = (double*)calloc(G, sizeof(double));
  if (c->which_mc_model == 5) {                                                      /* MCnucl.cpp:688-779 */
    if (c->sub_model == 1) {
      smc_o_add_density(c, np, proj8, a); smc_o_add_density(c, nt, targ8, a);
      double prefactor = (1.0 - c->alpha) / 2.;
      for (long k = 0; k < G; k++) a[k] = a[k] * prefactor;
    }
    if (c->alpha > 1e-8) smc_o_binary_term(c, nc, coll8, b);
    for (long k = 0; k < G; k++) { double d = a[k] + b[k]; rho[k] = d; dndy += d; }
  } else {                                                                           /* model 7, MCnucl.cpp:780-811 */
    smc_o_add_density(c, np, proj8, a); smc_o_add_density(c, nt, targ8, b);
    for (long k = 0; k < G; k++) { double d = sqrt(a[k] * b[k]); rho[k] = d; dndy += d; }
  }
  free(a); free(b);
  return dndy;
}

double smc_o_six_point(double x, double y, double v00, double v01, double v02, double v10, double v11, double v20) {
  double axx = 1.0 / 2.0 * (v00 - 2 * v10 + v20);                                    /* arsenal.cpp:33-54 */
  double axy = v00 - v01 - v10 + v11;
  double ayy = 1.0 / 2.0 * (v00 - 2 * v01 + v02);
  double bx = 1.0 / 2.0 * (-3.0 * v00 + 4 * v10 - v20);
  double by = 1.0 / 2.0 * (-3.0 * v00 + 4 * v01 - v02);
  double cc = v00;
  return axx * x * x + axy * x * y + ayy * y * y + bx * x + by * y + cc;
}

double smc_o_density_kln(const smc_o_cfg* c, const double* TA1, const double* TA2, const double* tb,
                         int tmax, double dT, double* rho) {                         /* MCnucl.cpp:654-687 */
  double dndy = 0.0;
  for (int ir = 0; ir < c->Maxx; ir++) for (int jr = 0; jr < c->Maxy; jr++) {
    long k = (long)ir * c->Maxy + jr;
    double di = TA1[k] / dT, dj = TA2[k] / dT;
    if ((di < 0 || di >= tmax - 2) || (dj < 0 || dj >= tmax - 2)) return -1.0;
    int i = (int)floor(di), j = (int)floor(dj);
#define TB(a, b) tb[(long)(a) * tmax + (b)]
    double r = smc_o_six_point(di - i, dj - j, TB(i, j), TB(i, j + 1), TB(i, j + 2), TB(i + 1, j), TB(i + 1, j + 1), TB(i + 2, j));
#undef TB
    rho[k] = r; dndy += r;
  }
  return dndy;
}

/* ======================================================================================
 * moments
 * ====================================================================================== */
void smc_o_cm_angle(const smc_o_cfg* c, const double* dens, int n, double* out4) {   /* GlueDensity.cpp:87-144 */
  double Xcm = 0, Ycm = 0, weight = 0;
  for (int i = 0; i < c->Maxx; i++) for (int j = 0; j < c->Maxy; j++) {
    double x = c->Xmin + i * c->dx, y = c->Ymin + j * c->dy;
    double wei = dens[(long)i * c->Maxy + j] * c->dx * c->dy;
    weight += wei; Xcm += x * wei; Ycm += y * wei;
  }
  Xcm /= weight; Ycm /= weight;
  double Nr = 0, Ni = 0;
  for (int i = 0; i < c->Maxx; i++) for (int j = 0; j < c->Maxy; j++) {
    double x = c->Xmin + i * c->dx - Xcm, y = c->Ymin + j * c->dy - Ycm;
    double th = atan2(y, x), wei = dens[(long)i * c->Maxy + j];
    double rwei = pow(sqrt(x * x + y * y), n);
    Nr += rwei * cos(n * th) * wei; Ni += rwei * sin(n * th) * wei;
  }
  out4[0] = Xcm; out4[1] = Ycm; out4[2] = -atan2(-Ni, -Nr) / n; out4[3] = weight;
}

void smc_o_eccentricities(const smc_o_cfg* c, const double* dens, int nbox, const double* B4,
                          int from_order, int to_order, double* out) {               /* MakeDensity.cpp:2244-2511 */
  const int MO = 10; const double eps = 1e-15;
  double *mr = out, *mi = out + 10, *pr = out + 20, *pi_ = out + 30, *rn = out + 40;
  double norm[10], normp[10];
  for (int i = 0; i < 53; i++) out[i] = 0.0;
  for (int i = 0; i < MO; i++) { norm[i] = 0; normp[i] = 0; }
  const int Maxx = c->Maxx, Maxy = c->Maxy; const double Xmin = c->Xmin, Ymin = c->Ymin, dx = c->dx, dy = c->dy;
  double xc = 0, yc = 0, total = 0;
  for (int i = 0; i < Maxx; i++) for (int j = 0; j < Maxy; j++) {                    /* :2273-2282 */
    double x = Xmin + i * dx, y = Ymin + j * dy, d = dens[(long)i * Maxy + j];
    xc += x * d; yc += y * d; total += d;
  }
  xc /= total; yc /= total;
  double density_norm = 0.;
  for (int ix = 0; ix < Maxx; ix++) {                                                /* :2285-2298 */
    double x_o = Xmin + ix * dx - xc;
    for (int j = 0; j < Maxy; j++) {
      double y_o = Ymin + j * dy - yc, r_o = sqrt(x_o * x_o + y_o * y_o), d = dens[(long)ix * Maxy + j];
      for (int io = 0; io < MO; io++) rn[io] += pow(r_o, io) * d;
      density_norm += d;
    }
  }
  for (int io = 0; io < MO; io++) rn[io] /= density_norm;
  /* hot-spot region = union of all boxes (MCnucl.cpp:1303-1323), then the boolean mask (:2341-2358) */
  double rXL = B4[0], rXR = B4[1], rYL = B4[2], rYR = B4[3];
  for (int k = 1; k < nbox; k++) {
    const double* b = B4 + 4 * k;
    rXL = b[0] < rXL ? b[0] : rXL; rXR = b[1] > rXR ? b[1] : rXR;
    rYL = b[2] < rYL ? b[2] : rYL; rYR = b[3] > rYR ? b[3] : rYR;
  }
  int nY = (int)((rYR - rYL) / dy), nX = (int)((rXR - rXL) / dx);
  if (nX < 0) nX = 0; if (nY < 0) nY = 0;
  unsigned char* mask = (unsigned char*)calloc((size_t)nX * nY + 1, 1);
  for (int k = 0; k < nbox; k++) {
    const double* b = B4 + 4 * k;
    int x0 = (int)((b[0] - rXL) / dx), x1 = (int)((b[1] - rXL) / dx);
    int y0 = (int)((b[2] - rYL) / dy), y1 = (int)((b[3] - rYL) / dx);             /* sic: dx (quirk Q5) */
    for (int xi = x0; xi < x1; xi++) for (int yi = y0; yi < y1; yi++)
      if (xi >= 0 && xi < nX && yi >= 0 && yi < nY) mask[(size_t)xi * nY + yi] = 1;
  }
  const int i0 = (int)((rXL - Xmin) / dx), j0 = (int)((rYL - Ymin) / dy);
  for (int order = from_order; order <= to_order && order < MO; order++) {          /* :2389-2430 */
    for (int iR = 0; iR < nX; iR++) for (int jR = 0; jR < nY; jR++) {
      if (!mask[(size_t)iR * nY + jR]) continue;
      int i = iR + i0, j = jR + j0;
      double x = Xmin + i * dx - xc, y = Ymin + j * dy - yc;
      double r = sqrt(x * x + y * y), theta = atan2(y, x);
      if (i < Maxx && i >= 0 && j < Maxy && j >= 0) {
        double d = dens[(long)i * Maxy + j];
        mr[order] += r * r * cos(order * theta) * d;
        mi[order] += r * r * sin(order * theta) * d;
        norm[order] += r * r * d;
        int m = (order == 1) ? 3 : order;
        pr[order] += pow(r, m) * cos(order * theta) * d;
        pi_[order] += pow(r, m) * sin(order * theta) * d;
        normp[order] += pow(r, m) * d;
      }
    }
    mr[order] = -mr[order] / (norm[order] + eps); mi[order] = -mi[order] / (norm[order] + eps);
    pr[order] = -pr[order] / (normp[order] + eps); pi_[order] = -pi_[order] / (normp[order] + eps);
  }
  free(mask);
  out[50] = total; out[51] = xc; out[52] = yc;
}

/* ======================================================================================
 * MC-KLN kT-factorisation integrand (KLNModel.cpp:219-277,360-399; KLNfunc.h:14-17).
 * Normalisation of getdNdy (KLNModel.cpp:97-125): g2hfac(=2) * Norm * I * 9/32.
 * ====================================================================================== */
static double kln_alpha_s(double q2) {                                               /* KLNModel.h:90-95 */
  const double alphaS = 0.5, lqcd2 = 0.2 * 0.2, Beta0 = (33.0 - 2.0 * 3.0) / (12 * M_PI);
  if (q2 <= lqcd2) return alphaS;
  double a = 1.0 / (Beta0 * log(q2 / lqcd2));
  return alphaS < a ? alphaS : a;
}
double smc_o_kln_integrand(const smc_o_kln* k, double rapidity, double ta, double tb, const double x[3]) {
  const double Ptmin = 0.1, Ptmax = 12.0, CF = (3.0 * 3.0 - 1.0) / (2 * 3.0);
  const double fac = CF * 2. / (3. * M_PI * M_PI);
  double pt = Ptmin + x[0] * (Ptmax - Ptmin), ktmax = pt, kt = ktmax * x[1], phi = 2 * M_PI * x[2];
  double ktsq1 = 0.25 * (pt * pt + kt * kt + 2 * kt * pt * cos(phi));
  double ktsq2 = 0.25 * (pt * pt + kt * kt - 2 * pt * kt * cos(phi));
  double mt = pt;
  double x1 = mt / k->ecm * exp(rapidity), x2 = mt / k->ecm / exp(rapidity);
  if (x1 > 1.0 || x2 > 1.0) return 0.0;
  double qs2a = ta * 2. / 1.53 * pow(0.01 / x1, k->lambda), qs2b = tb * 2. / 1.53 * pow(0.01 / x2, k->lambda);
  double alpa = kln_alpha_s(qs2a), alpb = kln_alpha_s(qs2b);
  double f1 = ((ktsq1 <= qs2a) ? fac / alpa : fac * qs2a / ktsq1 / alpa) * pow(1.0 - x1, 4.);
  double f2 = ((ktsq2 <= qs2b) ? fac / alpb : fac * qs2b / ktsq2 / alpb) * pow(1.0 - x2, 4.);
  double scale = ktsq1 > ktsq2 ? ktsq1 : ktsq2, m2 = mt * mt;
  double alp = kln_alpha_s(scale > m2 ? scale : m2);
  double result = alp * f1 * f2;
  if (k->pt_order == 2) result *= (Ptmax - Ptmin) * pt * pt; else result *= 2.0 * M_PI * (Ptmax - Ptmin) * pt;
  result /= mt * mt;
  result *= 2.0 * M_PI * kt * ktmax;
  result /= 4.0;
  return result;
}

/* ---- rcBK tabulated uGD (rcBKfunc.h:65-121, rcBKfunc.cpp:115-210; KLNModel.cpp:360-399) -------------------
 * tables: kt[iq][iy][ik], N_A[iq][iy][ik]; natural cubic spline in kt (gsl_interp_cspline), nearest bin in
 * Y = ln(x0/x) (dY = 0.1), linear in Q0^2 between neighbouring tables.  The table files are absent upstream:
 * parity of this function is pinned only against the reference run on SYNTHETIC tables with the GSL stand-in. */
void smc_o_spline_natural(const double* x, const double* y, int n, double* y2) {
  double* u = (double*)calloc(n, sizeof(double));
  y2[0] = 0.0; y2[n - 1] = 0.0;
  for (int i = 1; i + 1 < n; i++) {
    double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    double p = sig * y2[i - 1] + 2.0;
    y2[i] = (sig - 1.0) / p;
    u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
  }
  for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
  free(u);
}
static double spline_eval(const double* xa, const double* ya, const double* y2, int n, double x) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) { int k = (hi + lo) >> 1; if (xa[k] > x) hi = k; else lo = k; }
  double h = xa[hi] - xa[lo], a = (xa[hi] - x) / h, b = (x - xa[lo]) / h;
  return a * ya[lo] + b * ya[hi] + ((a * a * a - a) * y2[lo] + (b * b * b - b) * y2[hi]) * (h * h) / 6.0;
}
double smc_o_rcbk_func(const smc_o_rcbk* t, double qs0_2, double x, double kt2, double alp) {
  if (x < 0. || x > 1. || qs0_2 < 0) return 0.;
  const double dY = 0.1, x0 = 0.01, lgXlambda = 0.3;
  double Y = log(x0 / x);
  if (Y < 0.0) { qs0_2 *= exp(lgXlambda * Y); Y = 0.; }
  int iy = (int)(Y / dY + .5);
  if (iy >= t->maxY) iy = t->maxY - 1;
  double Q02 = 4. / 8. * qs0_2;
  int iq = (int)(Q02 / t->dQ0); iq -= 1;
  int iqoffset = 1;
  if (t->set == 100) { iq -= 1; iqoffset++; }
  if (iq == t->maxQ0 - 1) iq--; else if (iq > t->maxQ0 - 1) iq = t->maxQ0 - 2;
  double fac = kt2 / (6. * M_PI * M_PI * M_PI) / alp, k = sqrt(kt2);
  const int n = t->maxKt;
#define TAB(arr, q) ((arr) + ((size_t)(q) * t->maxY + iy) * n)
  if (iq >= 0) {
    double val1 = spline_eval(TAB(t->kt, iq), TAB(t->na, iq), TAB(t->y2, iq), n, k);
    double val2 = spline_eval(TAB(t->kt, iq + 1), TAB(t->na, iq + 1), TAB(t->y2, iq + 1), n, k);
    return (val1 + (val2 - val1) * (Q02 - (iq + iqoffset) * t->dQ0) / t->dQ0) * fac;
  }
  double val2 = spline_eval(TAB(t->kt, 0), TAB(t->na, 0), TAB(t->y2, 0), n, k);
  return (val2 * Q02 / (iqoffset * t->dQ0)) * fac;
#undef TAB
}
/* KLNModel::func with model = rcBKalbacete / Set2 (KLNModel.cpp:219-277,360-399) */
double smc_o_rcbk_integrand(const smc_o_kln* k, const smc_o_rcbk* t, double rapidity, double ta, double tb, const double x[3]) {
  const double Ptmin = 0.1, Ptmax = 12.0;
  double pt = Ptmin + x[0] * (Ptmax - Ptmin), ktmax = pt, kt = ktmax * x[1], phi = 2 * M_PI * x[2];
  double ktsq1 = 0.25 * (pt * pt + kt * kt + 2 * kt * pt * cos(phi));
  double ktsq2 = 0.25 * (pt * pt + kt * kt - 2 * pt * kt * cos(phi));
  double mt = pt, x1 = mt / k->ecm * exp(rapidity), x2 = mt / k->ecm / exp(rapidity);
  if (x1 > 1.0 || x2 > 1.0) return 0.0;
  const double q0 = (t->set == 100) ? 0.399 : 0.336;
  double qs2a = ta * k->siginNN200 / 10. * q0, qs2b = tb * k->siginNN200 / 10. * q0;
  double f1 = smc_o_rcbk_func(t, qs2a, x1, ktsq1, kln_alpha_s(ktsq1)) * pow(1.0 - x1, 4.);
  double f2 = smc_o_rcbk_func(t, qs2b, x2, ktsq2, kln_alpha_s(ktsq2)) * pow(1.0 - x2, 4.);
  double scale = ktsq1 > ktsq2 ? ktsq1 : ktsq2, m2 = mt * mt;
  double result = kln_alpha_s(scale > m2 ? scale : m2) * f1 * f2;
  if (k->pt_order == 2) result *= (Ptmax - Ptmin) * pt * pt; else result *= 2.0 * M_PI * (Ptmax - Ptmin) * pt;
  result /= mt * mt; result *= 2.0 * M_PI * kt * ktmax; result /= 4.0;
  return result;
}

/* Gauss-Legendre nodes on [0,1] by Newton on P_n */
static void gauleg01(int n, double* x, double* w) {
  for (int i = 0; i < (n + 1) / 2; i++) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1, z1;
    do {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1.0); z1 = z; z = z1 - p1 / pp;
    } while (fabs(z - z1) > 1e-15);
    x[i] = 0.5 * (1 - z); x[n - 1 - i] = 0.5 * (1 + z);
    w[i] = w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}
/* deterministic stand-in for KLNModel::ktF_MCintegral (BASES, KLNModel.cpp:177-213): product rule */
static double kln_dndy_any(const smc_o_kln* k, const smc_o_rcbk* t, double y, double ta, double tb, int npt, int nkt, int nphi);
double smc_o_kln_dndy(const smc_o_kln* k, double y, double ta, double tb, int npt, int nkt, int nphi) { return kln_dndy_any(k, 0, y, ta, tb, npt, nkt, nphi); }
double smc_o_rcbk_dndy(const smc_o_kln* k, const smc_o_rcbk* t, double y, double ta, double tb, int npt, int nkt, int nphi) { return kln_dndy_any(k, t, y, ta, tb, npt, nkt, nphi); }
static double kln_dndy_any(const smc_o_kln* k, const smc_o_rcbk* t, double y, double ta, double tb, int npt, int nkt, int nphi) {
  const double hbarC = 0.197327053, CF = (3.0 * 3.0 - 1.0) / (2 * 3.0), Norm = 2. / CF / (hbarC * hbarC);
  double *xp = malloc(sizeof(double) * npt), *wp = malloc(sizeof(double) * npt);
  double *xk = malloc(sizeof(double) * nkt), *wk = malloc(sizeof(double) * nkt);
  gauleg01(npt, xp, wp); gauleg01(nkt, xk, wk);
  double sum = 0.0;
  for (int a = 0; a < npt; a++) for (int b = 0; b < nkt; b++) {
    double s = 0.0;
    for (int p = 0; p < nphi; p++) { double x[3] = {xp[a], xk[b], (p + 0.5) / nphi}; s += t ? smc_o_rcbk_integrand(k, t, y, ta, tb, x) : smc_o_kln_integrand(k, y, ta, tb, x); }
    sum += wp[a] * wk[b] * s / nphi;
  }
  free(xp); free(wp); free(xk); free(wk);
  return 2.0 * Norm * sum * 9. / 32.;
}

/* ======================================================================================
 * NBD multiplicity fluctuations (cc_fluctuation_model 1, 2): MCnucl::fluctuateCurrentDensity
 * (MCnucl.cpp:868-905) -> NBD::rand(p, r) (NBD.cpp:31-90) -> RandomVariable's step-function envelope
 * (RandomVariable.cpp:190-216, 243-286).  log Gamma comes from libm instead of the reference's own
 * rational approximation (arsenal.cpp log_gamma_function): pmf values agree to ~1e-15, which no sample sees.
 * ====================================================================================== */
/* NBD::pdf (NBD.cpp:31-40): pmf of floor(k_in) for success probability p and r failures */
double smc_o_nbd_pdf(double p, double r, double k_in) {
  if (k_in < 0) return 0;
  int k = (int)floor(k_in);
  double prefactor = exp(lgamma(k + r) - lgamma(k + 1.0) - lgamma(r));      /* binomial_coefficient(k+r-1, k), arsenal.cpp:886-891 */
  return prefactor * pow(1 - p, r) * pow(p, k);
}

/* NBD::recalculateMode + RandomVariable::constructEnvelopTab: M = step_left + 6 intervals of width std starting at
 * mode - step_left*std; edge[0..M], height[m] = larger pmf of the two ends.  Returns M. */
int smc_o_nbd_envelope(double p, double r, double* edge, double* height) {
  double mode = (r <= 1) ? 1e-30 : p * (r - 1) / (1 - p);
  double std = sqrt(p * r) / (1 - p);
  int step_left;
  for (step_left = 6; step_left > 0; step_left--) if (mode - std * step_left >= 0) break;
  double LB = mode - step_left * std, RB = LB + std;
  double pdfLB = smc_o_nbd_pdf(p, r, LB), pdfRB = smc_o_nbd_pdf(p, r, RB);
  int M = step_left + 6;
  edge[0] = LB;
  for (int ii = 0; ii < M; ii++) {
    height[ii] = pdfLB > pdfRB ? pdfLB : pdfRB;
    edge[ii + 1] = RB;
    LB = RB; RB += std;
    pdfLB = smc_o_nbd_pdf(p, r, LB); pdfRB = smc_o_nbd_pdf(p, r, RB);
  }
  return M;
}

/* NBD::rand(p, r) literally: envelope tables, inverse-CDF draw, acceptance test; `next` yields the drand48 sequence */
long smc_o_nbd_rand(double p, double r, smc_o_rand48* st) {
  const double ZERO = 1e-15;
  if (p < ZERO) return 0;
  if (p + ZERO > 1.0) return 0;
  double edge[16], height[16], cum[16], cen[16], hv[16];
  int M = smc_o_nbd_envelope(p, r, edge, height);
  double std = sqrt(p * r) / (1 - p), sum = 0;
  cum[0] = 0;                                               /* envelopInvCDFTab: (sum, right edge) */
  for (int m = 0; m < M; m++) { sum += std * height[m]; cum[m + 1] = sum; }
  /* envelopPdfTab: centres edge[m] + std/2 with a zero cap on either side (nearest-neighbour lookup) */
  cen[0] = edge[0] - std / 2; hv[0] = 0;
  for (int m = 0; m < M; m++) { cen[m + 1] = edge[m] + std / 2; hv[m + 1] = height[m]; }
  cen[M + 1] = edge[M] + std / 2; hv[M + 1] = 0;
  double x = 0;
  for (long it = 0; ; it++) {
    /* RandomVariable::drand(LB, RB) on [0, sum] */
    double width = cum[M] - cum[0], dw = width * 1e-30;
    double y = cum[0] + dw + (width - 2 * dw) * smc_o_drand48(st);
    /* interpLinearMono (arsenal.cpp:258-283) with binarySearch (:644-676) */
    if (fabs(y - cum[0]) < (cum[1] - cum[0]) * 1e-30) x = edge[0];
    else {
      int i0 = 0, i1 = M, idx = (int)floor((i1 + i0) / 2.);
      while (i1 - i0 > 1) { if (cum[idx] < y) i0 = idx; else i1 = idx; idx = (int)floor((i1 + i0) / 2.); }
      x = edge[i0] + (edge[i0 + 1] - edge[i0]) / (cum[i0 + 1] - cum[i0]) * (y - cum[i0]);
    }
    /* interpNearestDirect (arsenal.cpp:141-166) on the equally spaced centres */
    double env;
    {
      double dx = cen[1] - cen[0];
      if (fabs(x - cen[0]) < dx * 1e-30) env = hv[0];
      else { long idx = (long)floor((x - cen[0]) / dx); env = (x - cen[idx] > dx / 2) ? hv[idx + 1] : hv[idx]; }
    }
    if (smc_o_drand48(st) < smc_o_nbd_pdf(p, r, x) / (1.0 * env + 1e-60)) return (long)x;
    if (it + 1 > 1000) return (long)x;                      /* MAXITER */
  }
}

/* The law of that sampler in closed form: the envelope draw is uniform on interval m with probability ~ std*height[m]
 * and survives with probability min(1, pmf/height[m]), so
 *     P(k) ~ sum_m |[edge_m, edge_m+1) n [k, k+1)| * min(height[m], pmf(k)),   k = floor(edge_0) ... floor(edge_M)
 * -- a NBD truncated to [mode - s*std, mode + 6*std) with fractional end cells (cells with 6 std < 1 always give 0).
 * weights[i] for k = *k0 + i, i < n (not normalised); returns n (<= cap). */
int smc_o_nbd_law(double p, double r, long* k0, double* weights, int cap) {
  const double ZERO = 1e-15;
  if (p < ZERO || p + ZERO > 1.0) { *k0 = 0; weights[0] = 1.0; return 1; }
  double edge[16], height[16];
  int M = smc_o_nbd_envelope(p, r, edge, height);
  long ka = (long)floor(edge[0]), kb = (long)floor(edge[M]);
  int n = 0;
  *k0 = ka;
  for (long k = ka; k <= kb && n < cap; k++, n++) {
    double pk = smc_o_nbd_pdf(p, r, (double)k), w = 0;
    for (int m = 0; m < M; m++) {
      double lo = edge[m] > (double)k ? edge[m] : (double)k, hi = edge[m + 1] < (double)(k + 1) ? edge[m + 1] : (double)(k + 1);
      if (hi > lo) w += (hi - lo) * (height[m] < pk ? height[m] : pk);
    }
    weights[n] = w;
  }
  return n;
}

/* inverse CDF of the closed-form law at u in [0,1): what the CUDA path evaluates per cell (one Philox uniform) */
long smc_o_nbd_quantile(double p, double r, double u) {
  const double ZERO = 1e-15;
  if (p < ZERO || p + ZERO > 1.0) return 0;
  double edge[16], height[16];
  int M = smc_o_nbd_envelope(p, r, edge, height);
  if (edge[M] <= 1.0) return 0;                              /* the whole envelope lies inside the cell k = 0 */
  long ka = (long)floor(edge[0]), kb = (long)floor(edge[M]);
  double tot = 0;
  for (int pass = 0; pass < 2; pass++) {
    double acc = 0, target = u * tot;
    for (long k = ka; k <= kb; k++) {
      double pk = smc_o_nbd_pdf(p, r, (double)k), w = 0;
      for (int m = 0; m < M; m++) {
        double lo = edge[m] > (double)k ? edge[m] : (double)k, hi = edge[m + 1] < (double)(k + 1) ? edge[m + 1] : (double)(k + 1);
        if (hi > lo) w += (hi - lo) * (height[m] < pk ? height[m] : pk);
      }
      acc += w;
      if (pass == 1 && acc > target) return k;
    }
    tot = acc;
  }
  return kb;
}

/* MCnucl::fluctuateCurrentDensity (MCnucl.cpp:868-905) with the per-cell uniforms supplied (u[cell] in [0,1)):
 * model 1: k = cc_fluctuation_k; model 2: k = kpp * min(TA1, TA2) * siginNN / 10 */
void smc_o_fluctuate_density(const smc_o_cfg* c, int model, double cc_k, const double* TA1, const double* TA2,
                             const double* u, double* rho) {
  const double HBARC = 0.197327053;
  const double kpp = 1.0 / M_PI * c->dx * c->dy * 1.0 * (0.25 * 0.25 / HBARC / HBARC);
  for (int ir = 0; ir < c->Maxx; ir++)
    for (int jr = 0; jr < c->Maxy; jr++) {
      const size_t q = (size_t)ir * c->Maxy + jr;
      double nb = rho[q] * c->dx * c->dy, n;
      if (model == 1) n = (double)smc_o_nbd_quantile(nb / (nb + cc_k), cc_k, u[q]);
      else {
        double k = kpp * (TA1[q] < TA2[q] ? TA1[q] : TA2[q]) * c->siginNN / 10;
        if (nb < 1e-10) n = nb; else n = (double)smc_o_nbd_quantile(nb / (nb + k), k, u[q]);
      }
      rho[q] = n / (c->dx * c->dy);
    }
}


/* ======================================================================================
 * 3-D extension: profile_3d::generate_3d_profile (scripts/generate_3d_profiles/profile_3d.cpp:274-325) with the
 * rapidities and widths given: src7 rows x y id eta sigma_x sigma_y sigma_eta; rho[neta][nx][ny]
 * ====================================================================================== */
void smc_o_profile3d(int nx, int ny, int neta, double dx, double dy, double deta, int n, const double* src7, double* rho) {
  for (long q = 0; q < (long)neta * nx * ny; q++) rho[q] = 0.0;
  for (int i = 0; i < n; i++) {
    const double* s = src7 + 7 * i;
    double part_x = s[0], part_y = s[1], part_eta = s[3];
    int idx_x0 = (int)(part_x / dx + (nx - 1) / 2), idx_y0 = (int)(part_y / dy + (ny - 1) / 2), idx_eta0 = (int)(part_eta / deta + (neta - 1) / 2);
    double sx = s[4], sy = s[5], se = s[6];
    int rx = (int)(6 * sx / dx), ry = (int)(6 * sy / dy), re = (int)(6 * se / deta);
    int xl = imax2(idx_x0 - rx, 0), xr = imin2(idx_x0 + rx, nx), yl = imax2(idx_y0 - ry, 0), yr = imin2(idx_y0 + ry, ny);
    int el = imax2(idx_eta0 - re, 0), er = imin2(idx_eta0 + re, neta);
    for (int j = el; j < er; j++) {
      double eg = (j - (neta - 1) / 2.) * deta;
      double dis_eta = (eg - part_eta) * (eg - part_eta) / (2. * se * se), norm_eta = 1. / sqrt(2. * M_PI * se * se);
      for (int k = xl; k < xr; k++) {
        double xg = (k - (nx - 1) / 2.) * dx;
        double dis_x = (xg - part_x) * (xg - part_x) / (2. * sx * sx), norm_x = 1. / sqrt(2. * M_PI * sx * sx);
        for (int l = yl; l < yr; l++) {
          double yg = (l - (ny - 1) / 2.) * dy;
          double dis_y = (yg - part_y) * (yg - part_y) / (2. * sy * sy), norm_y = 1. / sqrt(2. * M_PI * sy * sy);
          rho[((long)j * nx + k) * ny + l] += exp(-dis_eta - dis_x - dis_y) * norm_eta * norm_x * norm_y;
        }
      }
    }
  }
}
