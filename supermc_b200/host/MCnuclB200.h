// MCnuclB200 -- the binding a maintainer of the reference would add (option B): a class with the calls
// MakeDensity makes on MCnucl (reference src/MCnucl.h:87-146) that forwards to the C ABI in batches.  Header-only; the only
// dependencies are include/supermc_b200.h and a parameter reader with `double getVal(const std::string&)` (the reference's
// ParameterReader or this repo's).  MakeDensity's event loops keep their shape:
//
//     mc->generateNuclei(b); binary = mc->getBinaryCollision(); if (binary == 0 || mc->CentralityCut() == 0) { mc->deleteNucleus(); continue; }
//     mc->calculateThickness(); mc->setDensity(iy, -1); ... mc->getRho(iy, i, j) ... mc->getTA1(i, j) ...
//
// with two differences a maintainer has to know: (1) the rejection loop (b draw, Ncoll > 0, Npart window, dS/dy window)
// runs on the device -- every generateNuclei() call lands on the next ACCEPTED event, the `b` argument is ignored and
// lastB() returns the impact parameter that was drawn; (2) dumpEccentricities' grid loops are not needed: moments(n)
// holds the five columns of order n, totalEntropy() the sixth-last column (the grids stay available for code that
// wants them).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/supermc_b200.h"

template <class Reader>
class MCnuclB200T {
 public:
  // grids: which lattices every event keeps on the host side (SMC_RUN_* flags); events_per_batch <= smc_max_batch()
  explicit MCnuclB200T(Reader* p, unsigned grids = SMC_RUN_KEEP_RHO | SMC_RUN_THICKNESS, int events_per_batch = 256, int device = 0)
      : flags(SMC_RUN_MOMENTS | grids), nbatch(events_per_batch) {
    smc_params q; smc_params_default(&q);
    auto I = [&](const char* n) { return (int)p->getVal(n); };
    q.which_mc_model = I("which_mc_model"); q.sub_model = I("sub_model"); q.lambda = p->getVal("lambda");
    q.tmax = I("tmax"); q.tmax_subdivision = I("tmax_subdivision"); q.alpha = p->getVal("alpha");
    q.aproj = I("Aproj"); q.atarg = I("Atarg"); q.proj_deformed = I("proj_deformed"); q.targ_deformed = I("targ_deformed");
    q.include_nn_correlation = I("include_NN_correlation"); q.shape_of_nucleons = I("shape_of_nucleons");
    q.collision_criterion = I("collision_criterion"); q.shape_of_entropy = I("shape_of_entropy"); q.quark_width = p->getVal("quark_width");
    q.gauss_nucl_width = p->getVal("gauss_nucl_width"); q.gaussian_lambda = p->getVal("gaussian_lambda"); q.ecm = p->getVal("ecm");
    q.bmin = p->getVal("bmin"); q.bmax = p->getVal("bmax"); q.npmin = I("Npmin"); q.npmax = I("Npmax");
    q.cutdsdy = I("cutdSdy"); q.cutdsdy_lowerbound = p->getVal("cutdSdy_lowerBound"); q.cutdsdy_upperbound = p->getVal("cutdSdy_upperBound");
    q.randomseed = (int64_t)p->getVal("randomSeed"); q.finalfactor = p->getVal("finalFactor");
    q.ecc_from_order = I("ecc_from_order"); q.ecc_to_order = I("ecc_to_order");
    q.maxx = p->getVal("maxx"); q.maxy = p->getVal("maxy"); q.dx = p->getVal("dx"); q.dy = p->getVal("dy");
    q.cc_fluctuation_model = I("cc_fluctuation_model"); q.cc_fluctuation_gamma_theta = p->getVal("cc_fluctuation_Gamma_theta");
    q.cc_fluctuation_k = p->getVal("cc_fluctuation_k"); q.ny = 1; q.ymax = p->getVal("ymax"); q.max_batch = events_per_batch;
    params = q;
    if (smc_create(&q, device, &ctx) != SMC_OK) { const std::string m = ctx ? smc_last_error(ctx) : "smc_create failed"; if (ctx) smc_destroy(ctx); ctx = nullptr; throw std::runtime_error(m); }
    smc_constants k; smc_get_constants(ctx, &k); Maxx = k.maxx_cells; Maxy = k.maxy_cells;
    nbatch = std::min(nbatch, smc_max_batch(ctx));
    ev.resize((size_t)nbatch);
    if (q.which_mc_model == 1) makeTable();
  }
  ~MCnuclB200T() { if (ctx) smc_destroy(ctx); }
  MCnuclB200T(const MCnuclB200T&) = delete;
  MCnuclB200T& operator=(const MCnuclB200T&) = delete;

  void makeTable() { std::vector<double> t((size_t)1); if (smc_build_kln_table(ctx, nullptr) != SMC_OK) fail(); }     // MCnucl.cpp:911-960
  // MCnucl.h:123 -- the Npart window is a construction parameter of the engine (Npmin / Npmax): only a no-op change is accepted
  void setCentralityCut(int Nmin, int Nmax) { if (Nmin != params.npmin || Nmax != params.npmax) throw std::runtime_error("MCnuclB200: set Npmin / Npmax in the parameters"); }

  void generateNuclei(double /*b: drawn on the device*/) { next(); }                                                  // :112
  int getBinaryCollision() { return cur().ncoll; }                                                                    // :119
  int CentralityCut() { return 1; }                                                                                   // :122 (the device only returns accepted events)
  void deleteNucleus() {}                                                                                             // :113
  void calculateThickness() {}                                                                                        // :116 (part of the batch)
  void setDensity(int /*iy*/, int /*ipt*/) {}                                                                         // :114
  void calculate_rho_binary() { need(SMC_RUN_RHO_BINARY, "rho_binary"); }                                             // :118
  void calculate_spectator_density() { need(SMC_RUN_SPECTATORS, "spectator densities"); }                             // :141

  int getNpart1() { return cur().npart1; }
  int getNpart2() { return cur().npart2; }
  int getNcoll() { return cur().ncoll; }
  int getSpectators() { return cur().nspec; }                                                                         // :140
  double lastB() { return cur().b; }
  double getdNdy() { return cur().dsdy; }                                                                             // :104-108 (sum(rho) dx dy)
  double totalEntropy() { return cur().total; }
  const double* moments(int order) { return cur().mom[order - 1]; }        // Re eps_n, Im eps_n, Re eps'_n, Im eps'_n, <r^n>

  double getRho(int /*iy*/, int x, int y) { return grid(SMC_GRID_RHO, SMC_RUN_KEEP_RHO)[(size_t)i_cur * G() + (size_t)x * Maxy + y]; }   // :97
  double getTA1(int x, int y) { return grid(SMC_GRID_TA1, SMC_RUN_THICKNESS)[(size_t)i_cur * G() + (size_t)x * Maxy + y]; }           // :94
  double getTA2(int x, int y) { return grid(SMC_GRID_TA2, SMC_RUN_THICKNESS)[(size_t)i_cur * G() + (size_t)x * Maxy + y]; }           // :95
  double get_rho_binary(int x, int y) { return grid(SMC_GRID_RHO_BINARY, SMC_RUN_RHO_BINARY)[(size_t)i_cur * G() + (size_t)x * Maxy + y]; }
  double get_spectator_density(int id, int x, int y) { return grid(id == 1 ? SMC_GRID_SPEC_A : SMC_GRID_SPEC_B, SMC_RUN_SPECTATORS)[(size_t)i_cur * G() + (size_t)x * Maxy + y]; }
  int getMaxx() const { return Maxx; }
  int getMaxy() const { return Maxy; }
  smc_ctx* context() { return ctx; }

 private:
  size_t G() const { return (size_t)Maxx * Maxy; }
  [[noreturn]] void fail() { throw std::runtime_error(smc_last_error(ctx)); }
  void need(unsigned f, const char* what) { if (!(flags & f)) throw std::runtime_error(std::string("MCnuclB200: construct with the flag for ") + what); }
  smc_event_out& cur() { if (i_cur < 0) throw std::runtime_error("MCnuclB200: generateNuclei first"); return ev[(size_t)i_cur]; }
  void next() {                      // the next accepted event; a new device batch when the current one is used up
    if (i_cur >= 0 && i_cur + 1 < n_cur) { i_cur++; return; }
    if (smc_run_events(ctx, next_id, nbatch, flags, ev.data()) != SMC_OK) fail();
    next_id += (uint64_t)nbatch; n_cur = nbatch; i_cur = 0;
    for (auto& g : grids) g.clear();
  }
  const std::vector<double>& grid(int which, unsigned f) {
    need(f, "that lattice"); cur();
    std::vector<double>& g = grids[which];
    if (g.empty()) { g.resize((size_t)n_cur * G()); if (smc_get_grids(ctx, 0, n_cur, which, g.data()) != SMC_OK) fail(); }    // one strided copy per batch and kind
    return g;
  }
  smc_ctx* ctx = nullptr; smc_params params; unsigned flags; int nbatch, n_cur = 0, i_cur = -1, Maxx = 0, Maxy = 0; uint64_t next_id = 0;
  std::vector<smc_event_out> ev; std::vector<double> grids[SMC_GRID_KINDS];
};
