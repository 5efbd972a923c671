// ref_dump: harness around the UNMODIFIED reference classes (compiled from /root/reference/src where
// they lie). TEST INFRASTRUCTURE ONLY -- produces full-precision records that pin oracle/smc_oracle.c
// and the committed golden fixtures under tests/golden/.  Nothing in the product links this.
//
// It replaces only the reference's main.cpp (src/main.cpp:19-73): same ParameterReader start-up and
// seeding, then the event loop of MakeDensity::generateEccTable (src/MakeDensity.cpp:2143-2224) is
// driven through the public MCnucl API, and protected state is read through probe subclasses
// (MakeDensity.h:14-19, MCnucl.h:24-36 are `protected`, so no reference source is patched).
//
// usage: ref_dump <out.bin> <n_accepted_events> [name=value ...]
//   extra (harness-only) names, parsed by ParameterReader like any other:
//     dump_grids=1      write TA1, TA2, rho per accepted event
//     dump_extra=1      also rho_binary, spectator densities, spectator list
//     dump_rotate=1     also the recenter/rotate sequence of generate_profile_average for orders 2,3
//     dump_tries=1      write rejected tries too (collision parity on Ncoll==0 cases)
//
// record stream: [i64 name_len][name][i64 ndim][i64 dims...][f64 data...], little endian.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <sys/time.h>
#include <iostream>
#include <sstream>
#include <fstream>
#include <iomanip>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <time.h>
// Quark::fluctFactor (the per-quark multiplicity weight of shape_of_entropy = 3) has no getter: this harness -- and only
// this translation unit -- reads it by seeing the reference's class bodies with `class` spelt `struct` and `private` spelt
// `public` (the object layout is the same, no reference source is touched; the standard headers those files include are
// already in, so only the reference's own declarations are affected)
#define private public
#define class struct
#include "Particle.h"
#undef class
#undef private
#include "MakeDensity.h"
#include "ParamDefs.h"
#include "ParameterReader.h"

using namespace std;

struct McProbe : public MCnucl {
  using MCnucl::proj; using MCnucl::targ; using MCnucl::binaryCollision; using MCnucl::spectators;
  using MCnucl::TA1; using MCnucl::TA2; using MCnucl::rho_binary; using MCnucl::spectator_1;
  using MCnucl::spectator_2; using MCnucl::rho; using MCnucl::Maxx; using MCnucl::Maxy;
  using MCnucl::dndy; using MCnucl::gaussCal; using MCnucl::siginNN; using MCnucl::dsq;
  using MCnucl::dndyTable; using MCnucl::tmax; using MCnucl::dT; using MCnucl::binRapidity;
};
struct GdProbe : public GlueDensity { using GlueDensity::density; };
struct MdProbe : public MakeDensity {
  MdProbe(ParameterReader* p) : MakeDensity(p) {}
  using MakeDensity::mc; using MakeDensity::Maxx; using MakeDensity::Maxy;
  using MakeDensity::bmin; using MakeDensity::bmax; using MakeDensity::finalFactor; using MakeDensity::Npart; using MakeDensity::wf;
};

struct PartProbe : public Particle { using Particle::baseBox; };
static FILE* out;
static void wr(const string& name, const vector<long>& dims, const double* d) {
  long nl = name.size(); fwrite(&nl, 8, 1, out); fwrite(name.data(), 1, nl, out);
  long nd = dims.size(); fwrite(&nd, 8, 1, out);
  long n = 1; for (size_t i = 0; i < dims.size(); i++) { fwrite(&dims[i], 8, 1, out); n *= dims[i]; }
  if (n) fwrite(d, 8, n, out);
}
static void wr1(const string& name, const vector<double>& v) { vector<long> d(1, (long)v.size()); wr(name, d, v.data()); }
static void wr2(const string& name, long r, long c, const vector<double>& v) { vector<long> d; d.push_back(r); d.push_back(c); wr(name, d, v.data()); }
static void wrgrid(const string& name, double** g, int nx, int ny) {
  vector<double> v((size_t)nx * ny);
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) v[(size_t)i * ny + j] = g[i][j];
  wr2(name, nx, ny, v);
}

static void dump_nucleus(const string& name, Nucleus* nuc) {
  vector<Particle*>& n = nuc->getNucleons();
  vector<double> v;
  for (size_t i = 0; i < n.size(); i++) {
    Box2D b = n[i]->getBoundingBox();
    v.push_back(n[i]->getX()); v.push_back(n[i]->getY()); v.push_back(n[i]->getZ());
    v.push_back(b.getXL()); v.push_back(b.getXR()); v.push_back(b.getYL()); v.push_back(b.getYR());
    v.push_back((double)n[i]->getNumberOfCollision()); v.push_back(n[i]->getFluctfactor());
  }
  wr2(name, (long)n.size(), 9, v);
  // state the averaged-profile path depends on (quirk Q4): stale base box, quark offsets, AABB centre
  vector<double> x;
  for (size_t i = 0; i < n.size(); i++) {
    Box2D bb = static_cast<PartProbe*>(n[i])->baseBox, cb = n[i]->getBoundingBox();
    x.push_back(bb.getXL()); x.push_back(bb.getXR()); x.push_back(bb.getYL()); x.push_back(bb.getYR());
    vector<Quark>& q = n[i]->getQuarks();
    for (int k = 0; k < 3; k++) { x.push_back(q[k].getLocalX()); x.push_back(q[k].getLocalY()); x.push_back(q[k].getLocalZ()); }
    x.push_back(cb.getX()); x.push_back(cb.getY()); x.push_back(0.0);
  }
  wr2(name + "_x", (long)n.size(), 16, x);
  vector<double> qf;
  for (size_t i = 0; i < n.size(); i++) { vector<Quark>& q = n[i]->getQuarks(); for (int k = 0; k < 3; k++) qf.push_back(q[k].fluctFactor); }
  wr2(name + "_qf", (long)n.size(), 3, qf);
}

static void dump_grids(const string& pfx, McProbe* mc) {
  wrgrid(pfx + "TA1", mc->TA1, mc->Maxx, mc->Maxy);
  wrgrid(pfx + "TA2", mc->TA2, mc->Maxx, mc->Maxy);
  GdProbe* gd = static_cast<GdProbe*>(mc->rho);
  wrgrid(pfx + "rho", gd->density[0], mc->Maxx, mc->Maxy);
}

int main(int argc, char* argv[]) {
  if (argc < 3) { fprintf(stderr, "usage: ref_dump out.bin nev [name=value...]\n"); return 2; }
  out = fopen(argv[1], "wb");
  int nev = atoi(argv[2]);
  ParameterReader paraRdr;
  paraRdr.readFromFile("parameters.dat");
  paraRdr.setVal("dump_grids", 0); paraRdr.setVal("dump_extra", 0);
  paraRdr.setVal("dump_rotate", 0); paraRdr.setVal("dump_tries", 0); paraRdr.setVal("dump_text", 0); paraRdr.setVal("dump_ugd", 0);
  paraRdr.readFromArguments(argc, argv, "#", 3);
  int randomSeed = paraRdr.getVal("randomSeed");
  if (randomSeed < 0) randomSeed = 1;
  srand(randomSeed); srand48(randomSeed);                       // src/main.cpp:32-33
  MdProbe* dens = new MdProbe(&paraRdr);
  McProbe* mc = static_cast<McProbe*>(dens->mc);
  const int dgr = paraRdr.getVal("dump_grids"), dex = paraRdr.getVal("dump_extra");
  const int drot = paraRdr.getVal("dump_rotate"), dtr = paraRdr.getVal("dump_tries"), dtext = paraRdr.getVal("dump_text");
  const int from_order = paraRdr.getVal("ecc_from_order"), to_order = paraRdr.getVal("ecc_to_order");

  { vector<double> c;
    c.push_back(mc->siginNN); c.push_back(mc->gaussCal->width); c.push_back(mc->gaussCal->sigma_gg);
    c.push_back(mc->dsq); c.push_back(mc->Maxx); c.push_back(mc->Maxy); c.push_back(dens->finalFactor);
    wr1("consts", c); }
  if (mc->dndyTable) {          // MC-KLN look-up table (MCnucl.cpp:911-960)
    vector<double> t((size_t)mc->tmax * mc->tmax);
    for (int i = 0; i < mc->tmax; i++) for (int j = 0; j < mc->tmax; j++) t[(size_t)i * mc->tmax + j] = mc->dndyTable[0][i][j];
    wr2("kln_table", mc->tmax, mc->tmax, t);
    for (int iy = 1; iy < mc->binRapidity; iy++) {      // one table per rapidity slice (ny > 1)
      for (int i = 0; i < mc->tmax; i++) for (int j = 0; j < mc->tmax; j++) t[(size_t)i * mc->tmax + j] = mc->dndyTable[iy][i][j];
      char nm[32]; snprintf(nm, sizeof nm, "kln_table_y%d", iy); wr2(nm, mc->tmax, mc->tmax, t);
    }
    vector<double> c; c.push_back(mc->dT); c.push_back(mc->tmax); wr1("kln_consts", c);
  }

  if ((int)paraRdr.getVal("dump_ugd") && dens->wf) {      // UnintegPartonDist::getFunc on a fixed lattice of arguments
    const double qs[] = {0.05, 0.3, 0.41, 1.7, 5.5, 11.9, 14.0}, xs[] = {1e-5, 1e-3, 0.01, 0.05, 0.3}, ks[] = {0.01, 0.3, 2.5, 30., 120.};
    vector<double> v;
    for (int a = 0; a < 7; a++) for (int b2 = 0; b2 < 5; b2++) for (int c = 0; c < 5; c++) {
      v.push_back(qs[a]); v.push_back(xs[b2]); v.push_back(ks[c]); v.push_back(0.3);
      v.push_back(dens->wf->getFunc(qs[a], xs[b2], ks[c], 0.3));
    }
    wr2("ugd_samples", (long)(v.size() / 5), 5, v);
  }
  const int nyb = mc->binRapidity;
  double*** d1 = new double**[nyb];
  for (int iy = 0; iy < nyb; iy++) { d1[iy] = new double*[dens->Maxx]; for (int i = 0; i < dens->Maxx; i++) d1[iy][i] = new double[dens->Maxy](); }
  char eccfile[] = "data/h_ecc_%d.dat";

  int event = 0, tryid = 0;
  while (event < nev) {
    int binary = 0, accepted = 0;
    double b = sqrt((dens->bmax * dens->bmax - dens->bmin * dens->bmin) * drand48() + dens->bmin * dens->bmin);
    mc->generateNuclei(b);
    char pfx[64]; snprintf(pfx, sizeof pfx, "t%d/", tryid);
    // snapshot before collisions
    vector<Particle*> pn = mc->proj->getNucleons(), tn = mc->targ->getNucleons();
    map<Particle*, int> pidx, tidx;
    for (size_t i = 0; i < pn.size(); i++) pidx[pn[i]] = i;
    for (size_t i = 0; i < tn.size(); i++) tidx[tn[i]] = i;
    unsigned short tmp[3] = {0, 0, 0}, st[3];
    unsigned short* old = seed48(tmp); memcpy(st, old, 6); seed48(st);   // read the drand48 state, restore it
    binary = mc->getBinaryCollision();
    accepted = (binary != 0 && mc->CentralityCut() != 0);
    if (accepted || dtr) {
      string P(pfx);
      vector<double> hdr; hdr.push_back(b); hdr.push_back(binary); hdr.push_back(mc->getNpart1());
      hdr.push_back(mc->getNpart2()); hdr.push_back(accepted); hdr.push_back(st[0]); hdr.push_back(st[1]); hdr.push_back(st[2]);
      wr1(P + "hdr", hdr);
      dump_nucleus(P + "proj", mc->proj); dump_nucleus(P + "targ", mc->targ);
      vector<double> pp, tp;
      vector<Particle*>& wp = mc->proj->getParticipants(); vector<Particle*>& wt = mc->targ->getParticipants();
      for (size_t i = 0; i < wp.size(); i++) pp.push_back(pidx[wp[i]]);
      for (size_t i = 0; i < wt.size(); i++) tp.push_back(tidx[wt[i]]);
      wr1(P + "proj_part", pp); wr1(P + "targ_part", tp);
      // collisions in createBinaryCollisions order (MCnucl.cpp:326-352)
      vector<double> cv; size_t ic = 0;
      for (size_t i = 0; i < wp.size(); i++) {
        vector<Particle*>& cl = wp[i]->getCollidingParticles();
        for (size_t j = 0; j < cl.size(); j++, ic++) {
          CollisionPair* c = mc->binaryCollision[ic];
          cv.push_back(c->getX()); cv.push_back(c->getY()); cv.push_back(c->getfluctfactor());
          cv.push_back(c->additional_weight); cv.push_back(pidx[wp[i]]); cv.push_back(tidx[cl[j]]);
        }
      }
      wr2(P + "coll", (long)mc->binaryCollision.size(), 6, cv);
    }
    if (accepted) {
      string P(pfx);
      mc->calculateThickness();
      mc->setDensity(0, -1);
      vector<double> s; s.push_back(mc->dndy); wr1(P + "dndy", s);
      if (dgr) dump_grids(P, mc);
      { vector<Box2D> hs; Box2D r = mc->getHotSpots(hs);
        vector<double> v; v.push_back(r.getXL()); v.push_back(r.getXR()); v.push_back(r.getYL()); v.push_back(r.getYR());
        wr1(P + "region", v); }
      dens->setSd(d1, 0);
      dens->dumpEccentricities(eccfile, d1, 0, from_order, to_order, mc->getNpart1() + mc->getNpart2(), mc->getNcoll(), b);
      for (int iy = 1; iy < nyb; iy++) {      // the other rapidity slices (generateEccTable, MakeDensity.cpp:2170-2193): one more row each
        mc->setDensity(iy, -1);
        dens->setSd(d1, iy);
        dens->dumpEccentricities(eccfile, d1, iy, from_order, to_order, mc->getNpart1() + mc->getNpart2(), mc->getNcoll(), b);
        if (dgr) { char q[96]; snprintf(q, sizeof q, "%srho_y%d", pfx, iy); GdProbe* gd = static_cast<GdProbe*>(mc->rho); wrgrid(q, gd->density[iy], mc->Maxx, mc->Maxy); }
      }
      if (nyb > 1) mc->setDensity(0, -1);
      if (dtext) {      // the reference's own text writers on this event (format fixtures)
        char f1[] = "data/ref_block.dat", f2[] = "data/ref_4col.dat";
        dens->Npart = mc->getNpart1() + mc->getNpart2();
        dens->dumpDensityBlock(f1, d1, 0); dens->dumpDensity4Col(f2, d1, 0);
        char fp[64]; snprintf(fp, sizeof fp, "data/ref_participants_%d.dat", event); mc->dumpparticipantTable(fp);
        snprintf(fp, sizeof fp, "data/ref_binary_%d.dat", event); mc->dumpBinaryTable(fp);
      }
      if (dex) {
        mc->calculate_rho_binary();
        mc->getSpectators();
        if (dtext) mc->dumpSpectatorsTable(1000 + event);
        mc->calculate_spectator_density();
        wrgrid(P + "rho_binary", mc->rho_binary, mc->Maxx, mc->Maxy);
        wrgrid(P + "spec1", mc->spectator_1, mc->Maxx, mc->Maxy);
        wrgrid(P + "spec2", mc->spectator_2, mc->Maxx, mc->Maxy);
        vector<double> sv;
        for (size_t i = 0; i < mc->spectators.size(); i++) {
          sv.push_back(mc->spectators[i]->getX()); sv.push_back(mc->spectators[i]->getY());
          sv.push_back(mc->spectators[i]->getRapidity_Y());
        }
        wr2(P + "spectators", (long)mc->spectators.size(), 3, sv);
      }
      if (drot) {
        // the per-order sequence of generate_profile_average (MakeDensity.cpp:1289-1340), sd branch
        if (!dex) mc->getSpectators();
        for (int order = 2; order <= 3; order++) {
          char q[96];
          mc->setDensity(0, -1);
          mc->recenterGrid(0, order); mc->calculateThickness(); mc->setDensity(0, -1);
          snprintf(q, sizeof q, "%srp%d/", pfx, order);
          { double xc, yc; mc->rho->getCM(xc, yc, 0); vector<double> v; v.push_back(xc); v.push_back(yc); v.push_back(mc->rho->getCMAngle(0)); wr1(string(q) + "cm", v); }
          dump_grids(q, mc); dump_nucleus(string(q) + "proj", mc->proj); dump_nucleus(string(q) + "targ", mc->targ);
          mc->rotateGrid(0, order); mc->calculateThickness(); mc->setDensity(0, -1);
          snprintf(q, sizeof q, "%srot%d/", pfx, order);
          { double xc, yc; mc->rho->getCM(xc, yc, 0); vector<double> v; v.push_back(xc); v.push_back(yc); v.push_back(mc->rho->getCMAngle(0)); wr1(string(q) + "cm", v); }
          dump_grids(q, mc); dump_nucleus(string(q) + "proj", mc->proj); dump_nucleus(string(q) + "targ", mc->targ);
          mc->calculate_rho_binary(); mc->calculate_spectator_density();
          wrgrid(string(q) + "rho_binary", mc->rho_binary, mc->Maxx, mc->Maxy);
          wrgrid(string(q) + "spec1", mc->spectator_1, mc->Maxx, mc->Maxy);
          wrgrid(string(q) + "spec2", mc->spectator_2, mc->Maxx, mc->Maxy);
          vector<double> cv;
          for (size_t i = 0; i < mc->binaryCollision.size(); i++) { cv.push_back(mc->binaryCollision[i]->getX()); cv.push_back(mc->binaryCollision[i]->getY()); }
          wr2(string(q) + "coll_xy", (long)mc->binaryCollision.size(), 2, cv);
        }
      }
      event++;
    }
    mc->deleteNucleus();
    tryid++;
  }
  fclose(out);
  return 0;
}
