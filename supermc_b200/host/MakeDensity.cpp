#include "MakeDensity.h"
#include <algorithm>
#include <cmath>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <dirent.h>
#include <unistd.h>
#include <fstream>
#include <functional>
#include <iostream>
#include <mutex>
#include <sstream>
#include <sys/stat.h>
#include <thread>

// ---- a small pool of writer threads: text formatting caps operations 1/2 in the reference ----------
namespace {
class WriterPool {
 public:
  explicit WriterPool(int n) : stop(false), pending(0) { for (int i = 0; i < n; i++) th.emplace_back([this] { loop(); }); }
  ~WriterPool() { wait(); { std::lock_guard<std::mutex> l(m); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); }
  void submit(std::function<void()> f) {
    std::unique_lock<std::mutex> l(m);
    cv_room.wait(l, [this] { return q.size() < 64; });        // bound the memory held by queued grids
    q.push_back(std::move(f)); pending++; cv.notify_one();
  }
  void wait() { std::unique_lock<std::mutex> l(m); cv_done.wait(l, [this] { return pending == 0; }); }
 private:
  void loop() {
    for (;;) {
      std::function<void()> f;
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [this] { return stop || !q.empty(); }); if (stop && q.empty()) return; f = std::move(q.front()); q.pop_front(); cv_room.notify_one(); }
      f();
      { std::lock_guard<std::mutex> l(m); pending--; if (pending == 0) cv_done.notify_all(); }
    }
  }
  std::vector<std::thread> th; std::deque<std::function<void()>> q; std::mutex m; std::condition_variable cv, cv_room, cv_done; bool stop; int pending;
};

void write_file(const std::string& name, const std::string& text, bool append) {
  FILE* f = std::fopen(name.c_str(), append ? "ab" : "wb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", name.c_str()); return; }
  std::fwrite(text.data(), 1, text.size(), f); std::fclose(f);
}
inline void put(std::string& s, const char* fmt, double v) { char b[64]; int n = std::snprintf(b, sizeof b, fmt, v); s.append(b, n); }
// "%<width>.<prec>g" without printf: std::to_chars(general, prec) is specified to produce what printf("%.<prec>g") produces in
// the C locale (shortest of %e / %f at that precision, trailing zeros removed) and is several times faster; the text
// formatting is what bounds operations 1, 2 and 9 once the physics runs on the GPU
inline void put_g(std::string& s, double v, int width, int prec) {
  char b[64];
  const auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::general, prec);
  const int n = (int)(r.ptr - b);
  if (n < width) s.append((size_t)(width - n), ' ');
  s.append(b, (size_t)n);
}
int ival(ParameterReader* p, const char* n) { return (int)p->getVal(n); }
}  // namespace

// ---- formatting (printf %g == iostream default floatfield with the same precision) -----------------
namespace {
// one "%<width>.<prec>g" cell, right-aligned; returns its length (>= width)
inline int fmt_cell(char* dst, double v, int width, int prec) {
  char b[40];
  const auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::general, prec);
  const int n = (int)(r.ptr - b), pad = n < width ? width - n : 0;
  std::memset(dst, ' ', (size_t)pad); std::memcpy(dst + pad, b, (size_t)n);
  return pad + n;
}
struct RowCells { char mom[45][24]; int mlen[45]; char tail[4 * 24 + 8 * 24 + 2]; int tlen; };
// the 49 (+4 for deformed nuclei) cells of one event: 45 moments, then Npart Ncoll total b [0 0 0 0] "\n" as one tail
inline void format_cells(const smc_event_out& ev, bool deformed, RowCells& rc) {
  for (int n = 0; n < 9; n++) for (int k = 0; k < 5; k++) rc.mlen[n * 5 + k] = fmt_cell(rc.mom[n * 5 + k], ev.mom[n][k], 16, 8);
  int t = 0;
  t += fmt_cell(rc.tail + t, (double)(ev.npart1 + ev.npart2), 10, 5); t += fmt_cell(rc.tail + t, (double)ev.ncoll, 10, 5);
  t += fmt_cell(rc.tail + t, ev.total, 16, 8); t += fmt_cell(rc.tail + t, ev.b, 16, 8);
  if (deformed) for (int k = 0; k < 4; k++) t += fmt_cell(rc.tail + t, 0.0, 16, 8);   // mc->lastCx1.. are never assigned upstream (quirk Q9)
  rc.tail[t++] = '\n'; rc.tlen = t;
}
}  // namespace
std::string MakeDensity::formatEccRow(const smc_event_out& ev, int order, bool deformed) {
  RowCells rc; format_cells(ev, deformed, rc);
  std::string s;
  for (int k = 0; k < 5; k++) s.append(rc.mom[(order - 1) * 5 + k], (size_t)rc.mlen[(order - 1) * 5 + k]);
  s.append(rc.tail, (size_t)rc.tlen); return s;
}
std::string MakeDensity::formatEccRowAll(const smc_event_out& ev, bool deformed) {
  RowCells rc; format_cells(ev, deformed, rc);
  std::string s; s.reserve(49 * 16 + 8);
  for (int q = 0; q < 45; q++) s.append(rc.mom[q], (size_t)rc.mlen[q]);
  s.append(rc.tail, (size_t)rc.tlen); return s;
}
void MakeDensity::formatDensityBlock(const double* g, int Maxx, int Maxy, std::string& out) {
  // "%22.12g" cells written straight into the final buffer; most of a lattice is exact zeros (outside the event's
  // rectangle), which print as "0" without a conversion
  out.resize((size_t)Maxx * ((size_t)Maxy * 26 + 1));
  char* p = &out[0];
  for (int i = 0; i < Maxx; i++) {
    for (int j = 0; j < Maxy; j++) {
      const double v = g[(size_t)i * Maxy + j];
      if (v == 0.0 && !std::signbit(v)) { std::memset(p, ' ', 21); p[21] = '0'; p += 22; }
      else p += fmt_cell(p, v, 22, 12);
    }
    *p++ = '\n';
  }
  out.resize((size_t)(p - &out[0]));
}
void MakeDensity::formatDensity4Col(const double* g, int Maxx, int Maxy, double Xmin, double Ymin, double dx, double dy,
                                    double rap, double npart, std::string& out) {
  out.clear(); out.reserve((size_t)Maxx * Maxy * 53 + 64);
  char b[160];
  int n = std::snprintf(b, sizeof b, "# <npart>= %g xmax= %d ymax= %d\n", npart, Maxx, Maxy); out.append(b, n);
  std::string head; put_g(head, rap, 10, 3);
  for (int i = 0; i < Maxx; i++) {
    std::string xs; put_g(xs, Xmin + i * dx, 10, 3);
    for (int j = 0; j < Maxy; j++) {
      out += head; out += xs; put_g(out, Ymin + j * dy, 10, 3); put_g(out, g[(size_t)i * Maxy + j], 22, 12); out += "\n";
    }
  }
}

// ---- construction: parameters.dat keys -> smc_params (consumers listed in SURVEY.md appendix A) ------
MakeDensity::MakeDensity(ParameterReader* p, int device, smc_shard sh, const std::string& dd)
    : paraRdr(p), ctx(nullptr), ctx_ok(false), data_dir(dd), root_data_dir(dd), shard(sh) {
  if (shard.world > 1 && shard.rank > 0) {               // private output directory per rank, merged by rank 0 at the end
    data_dir = dd + "_rank" + std::to_string(shard.rank);
    mkdir(data_dir.c_str(), 0777);
  }
  bool seed_from_clock = false;
  try {
    smc_params_default(&params);
    params.which_mc_model = ival(p, "which_mc_model"); params.sub_model = ival(p, "sub_model");
    params.lambda = p->getVal("lambda"); params.tmax = ival(p, "tmax"); params.tmax_subdivision = ival(p, "tmax_subdivision");
    params.alpha = p->getVal("alpha"); params.aproj = ival(p, "Aproj"); params.atarg = ival(p, "Atarg");
    params.proj_deformed = ival(p, "proj_deformed"); params.targ_deformed = ival(p, "targ_deformed");
    params.include_nn_correlation = ival(p, "include_NN_correlation");
    params.shape_of_nucleons = ival(p, "shape_of_nucleons"); params.collision_criterion = ival(p, "collision_criterion");
    params.shape_of_entropy = ival(p, "shape_of_entropy"); params.quark_width = p->getVal("quark_width");
    params.gauss_nucl_width = p->getVal("gauss_nucl_width"); params.gaussian_lambda = p->getVal("gaussian_lambda"); params.ecm = p->getVal("ecm");
    params.bmin = p->getVal("bmin"); params.bmax = p->getVal("bmax"); params.npmin = ival(p, "Npmin"); params.npmax = ival(p, "Npmax");
    params.cutdsdy = ival(p, "cutdSdy"); params.cutdsdy_lowerbound = p->getVal("cutdSdy_lowerBound"); params.cutdsdy_upperbound = p->getVal("cutdSdy_upperBound");
    long long seed = (long long)p->getVal("randomSeed");
    if (seed < 0) { seed_from_clock = true; seed = (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count() % 1000000; }  // main.cpp:28-31
    params.randomseed = seed;
    params.finalfactor = p->getVal("finalFactor"); params.ecc_from_order = ival(p, "ecc_from_order"); params.ecc_to_order = ival(p, "ecc_to_order");
    params.maxx = p->getVal("maxx"); params.maxy = p->getVal("maxy"); params.dx = p->getVal("dx"); params.dy = p->getVal("dy");
    params.cc_fluctuation_model = ival(p, "cc_fluctuation_model"); params.cc_fluctuation_gamma_theta = p->getVal("cc_fluctuation_Gamma_theta");
    params.cc_fluctuation_k = p->getVal("cc_fluctuation_k");
    const double ptflag = p->getVal("PT_Flag");
    params.pt_order = ptflag < 0 ? ival(p, "PT_order") : 1;
    params.max_batch = (int)p->getVal("gpu_batch", 0);
    params.ncoll_cap = (int)p->getVal("ncoll_cap", 0);                  // extension: collision-list capacity per event (0: min(A*B, 6144))
    binRapidity = ival(p, "ny");
    rapMin = -p->getVal("ymax"); rapMax = -rapMin;                       // MakeDensity.cpp:54-58
    params.ny = binRapidity; params.ymax = p->getVal("ymax");
    p->setVal("rapMin", rapMin); p->setVal("rapMax", rapMax);
    finalFactor = params.finalfactor;
    deformed = (params.proj_deformed == 1 || params.targ_deformed == 1);
    if (binRapidity < 1) { err = "ny must be >= 1"; return; }
    if (ptflag < 0) { err = "PT_Flag < 0 (pT-differential tables) is unreachable in the reference (MCnucl.cpp:973 reads a misspelt key) and not built"; return; }
  } catch (std::exception& e) { err = e.what(); return; }
  const int rc = smc_create(&params, device, &ctx);
  if (rc != SMC_OK) { err = std::string("smc_create: ") + (ctx ? smc_last_error(ctx) : "failed"); return; }
  smc_get_constants(ctx, &k);
  p->setVal("siginNN", k.siginnn);                                        // MCnucl.cpp:93
  if (shard.world > 1) {                                                   // one run over several GPUs: rendezvous, one seed for all
    const char* addr = std::getenv("MASTER_ADDR"); const char* mp = std::getenv("MASTER_PORT"); const char* cp = std::getenv("SMC_COMM_PORT");
    const int port = cp ? std::atoi(cp) : (mp ? std::atoi(mp) + 1 : 29517);
    if (smc_comm_init(ctx, shard.rank, shard.world, addr ? addr : "127.0.0.1", port) != SMC_OK) { err = std::string("smc_comm_init: ") + smc_last_error(ctx); return; }
    if (seed_from_clock) {
      int64_t sd = params.randomseed;
      if (smc_comm_bcast_i64(ctx, &sd) != SMC_OK || smc_set_seed(ctx, sd) != SMC_OK) { err = smc_last_error(ctx); return; }
      params.randomseed = sd;
    }
  }
  Maxx = k.maxx_cells; Maxy = k.maxy_cells; Xmin = -params.maxx; Ymin = -params.maxy; dx = params.dx; dy = params.dy;
  if (load_tables() != 0) return;
  ctx_ok = true;
}

MakeDensity::~MakeDensity() { if (ctx) smc_destroy(ctx); }

// tables/: QuarkPos.txt (r1 r2 cos12 rows), light-ion and NN-correlated configurations (Nucleus.cpp:37-48,383-522)
static bool read_doubles(const std::string& file, std::vector<double>& v) {
  FILE* f = std::fopen(file.c_str(), "rb");
  if (!f) return false;
  std::string buf; char tmp[1 << 16]; size_t n;
  while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
  std::fclose(f);
  v.clear();
  const char* p = buf.c_str(); char* end = nullptr;
  for (;;) { const double x = std::strtod(p, &end); if (end == p) break; v.push_back(x); p = end; }
  return true;
}
int MakeDensity::load_tables() {
  std::vector<double> v;
  if (read_doubles("tables/QuarkPos.txt", v) && v.size() >= 3) {
    if (smc_load_quark_table(ctx, v.data(), (int)(v.size() / 3)) != SMC_OK) { err = smc_last_error(ctx); return 1; }
  } else std::cerr << "# tables/QuarkPos.txt not found: nucleon AABBs are the +-4w base boxes" << std::endl;
  const int A[2] = {params.aproj, params.atarg};
  for (int s = 0; s < 2; s++) {
    std::string file; int skip_head = 0, skip_tail = 0, per_nucleon = 3;
    if (A[s] == 3) { file = "tables/he3_plaintext.dat"; skip_tail = 4; }
    else if (A[s] == 4) file = "tables/he4_plaintext.dat";
    else if (A[s] == 12) { file = "tables/carbon_plaintext.dat"; skip_head = 2; }
    else if (A[s] == 16) file = "tables/oxygen_plaintext.dat";
    else if (params.include_nn_correlation == 1 && A[s] == 197) { file = "tables/au197-sw-full_3Bchains-conf1820.dat"; per_nucleon = 5; }
    else if (params.include_nn_correlation == 1 && A[s] == 208) { file = "tables/pb208-1.dat"; per_nucleon = 4; }
    else continue;
    if (!read_doubles(file, v)) { err = "Error: " + file + " does not find!"; return 1; }
    const size_t per_cfg = (size_t)skip_head + (size_t)A[s] * per_nucleon + skip_tail;
    const size_t ncfg = v.size() / per_cfg;
    std::vector<double> xyz(ncfg * A[s] * 3);
    for (size_t c = 0; c < ncfg; c++) for (int i = 0; i < A[s]; i++) for (int d = 0; d < 3; d++)
      xyz[(c * A[s] + i) * 3 + d] = v[c * per_cfg + skip_head + (size_t)i * per_nucleon + d];
    if (ncfg == 0 || smc_load_config_table(ctx, s, xyz.data(), (int)ncfg, A[s]) != SMC_OK) { err = std::string("configuration table ") + file + ": " + smc_last_error(ctx); return 1; }
  }
  if (params.which_mc_model == 1 && (params.sub_model == 100 || params.sub_model == 101)) {
    // rcBKfunc::rcBKfunc (rcBKfunc.cpp:15-178): javier/ft_rcbk_mv_qs02_*.dat, 121 Y-bins x 101 rows "Y kt N_F N_A"
    const int nq = params.sub_model == 100 ? 59 : 30, maxy = 121, maxkt = 101;
    std::vector<double> kt((size_t)nq * maxy * maxkt), na(kt.size());
    for (int iq = 0; iq < nq; iq++) {
      char name[96];
      if (params.sub_model == 100) {
        const int q = iq + 2;
        if (q % 10 == 0) std::snprintf(name, sizeof name, "javier/ft_rcbk_mv_qs02_%d_ad.dat", q / 10);
        else std::snprintf(name, sizeof name, "javier/ft_rcbk_mv_qs02_%02d_ad.dat", q);
      } else std::snprintf(name, sizeof name, "javier/ft_rcbk_mv_qs02_0168_g1_119_%d.dat", iq + 1);
      if (!read_doubles(name, v)) { err = std::string("Error unable to open file ") + name; return 1; }
      const size_t need = (size_t)maxy * maxkt * 4;
      if (v.size() < need) { err = "ERROR reading phi(x,kt) tables, too few entries !"; return 1; }
      if (v.size() > need) { err = "ERROR reading phi(x,kt) tables, too many entries !"; return 1; }
      for (size_t r = 0; r < (size_t)maxy * maxkt; r++) { kt[(size_t)iq * maxy * maxkt + r] = v[4 * r + 1]; na[(size_t)iq * maxy * maxkt + r] = v[4 * r + 3]; }
    }
    if (smc_load_rcbk_tables(ctx, kt.data(), na.data(), nq, maxy, maxkt) != SMC_OK) { err = smc_last_error(ctx); return 1; }
  }
  if (params.which_mc_model == 1) {                                       // MakeDensity.cpp:108-132
    std::cout << "MCnucl::makeTable(): precalculating dNdy for all combinations of Ta and Tb." << std::endl;
    std::vector<double> tab((size_t)k.kln_tmax * k.kln_tmax);
    if (smc_build_kln_table(ctx, tab.data()) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    if (shard.rank == 0) {                                                // dumpdNdyTable4Col, MCnucl.cpp:1027-1048 (append)
      std::string s; char b[128];
      for (int i = 1; i < k.kln_tmax; i++) for (int j = 1; j < k.kln_tmax; j++) {
        int n = std::snprintf(b, sizeof b, "%10.3f%10.3f%10.3f%22.12f\n", rapMin, k.kln_dt * i, k.kln_dt * j, tab[(size_t)i * k.kln_tmax + j]); s.append(b, n);
      }
      write_file(path("dNdyTable.dat"), s, true);
    }
    std::cout << "MCnucl::makeTable(): done" << std::endl << std::endl;
  }
  return 0;
}

void MakeDensity::shard_range(int nevent, uint64_t* first, int* count) const {
  const long long lo = (long long)nevent * shard.rank / shard.world, hi = (long long)nevent * (shard.rank + 1) / shard.world;
  *first = (uint64_t)lo; *count = (int)(hi - lo);
}

int MakeDensity::run(int operation, int nevent) {
  switch (operation) {
    case 1: return generate_profile_ebe(nevent);
    case 2: return generate_profile_ebe_Jet(nevent);
    case 3: return generate_profile_average(nevent);
    case 9: return generateEccTable(nevent);
    default: std::cout << "Error: operation choice " << operation << " not recognized." << std::endl; return 1;
  }
}

// ---- operation 9: minimum-bias eccentricity table ---------------------------------------------------
// The GPU produces ~0.75 M events/s and every event is 2 x 1970 bytes of text (ten tables per branch), so the loop is a
// three-stage pipeline over chunks of events: the main thread keeps the GPU busy (smc_run_events on chunk k+1), a formatter
// stage turns chunk k into text on all host cores (contiguous slices, so the files keep event order), and a writer stage
// appends chunk k-1 to the ten / twenty tables, one appender per file.  Every number of an event is formatted once: the
// nine per-order rows and the all-orders row are assembled from the same 49 right-aligned cells.
int MakeDensity::generateEccTable(int nevent) {
  const int from_order = ival(paraRdr, "ecc_from_order"), to_order = ival(paraRdr, "ecc_to_order");
  const bool use_sd = paraRdr->getVal("use_sd") != 0, use_ed = paraRdr->getVal("use_ed") != 0;
  const bool binary = paraRdr->getVal("output_binary", 0) != 0;          // extension: raw smc_event_out rows next to the text tables
  uint64_t first; int count; shard_range(nevent, &first, &count);
  const int chunk = std::max(1, (int)paraRdr->getVal("host_chunk", 65536));
  const int lo = std::max(from_order, 1), hi = std::min(to_order, 9);
  const unsigned nthr = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  const int ny = binRapidity;            // one row per event and rapidity slice (MakeDensity.cpp:2170-2193)
  struct Chunk { std::vector<smc_event_out> out; int n = 0; long done = 0; };
  struct Text { std::vector<std::vector<std::string>> part; const Chunk* src = nullptr; long done = 0; };   // part[thread][table 1..10]
  Chunk buf[3]; for (auto& b : buf) b.out.resize((size_t)std::max(1, std::min(count, chunk)) * ny);
  Text txt[2]; for (auto& t : txt) t.part.assign(nthr, std::vector<std::string>(11));
  std::mutex m; std::condition_variable cv;
  std::deque<Chunk*> full, empty; std::deque<Text*> tfull, tempty; bool finished = false, formatted = false; long failed = 0;
  for (auto& b : buf) empty.push_back(&b);
  for (auto& t : txt) tempty.push_back(&t);
  const char* base[2] = {"sn_ecc_eccp_%d.dat", "en_ecc_eccp_%d.dat"};     // en == sn numerically (quirk Q2)
  // stage 3: append a formatted chunk to the tables, one appender per file, all at once (the copy into the page cache is
  // what bounds a single writer at ~2 GB/s, and a million events are 4 GB of text)
  std::thread writer([&] {
    for (;;) {
      Text* t;
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return formatted || !tfull.empty(); }); if (tfull.empty()) return; t = tfull.front(); tfull.pop_front(); }
      std::vector<std::thread> wr;
      for (int f = 0; f < 2; f++) {
        if ((f == 0 && !use_sd) || (f == 1 && !use_ed)) continue;
        for (int o = lo; o <= 10; o++) {
          if (o > hi && o < 10) continue;
          wr.emplace_back([&, f, o] {
            char name[128]; std::snprintf(name, sizeof name, base[f], o);
            FILE* fp = std::fopen(path(name).c_str(), "ab");
            if (!fp) { std::fprintf(stderr, "cannot open %s\n", path(name).c_str()); return; }
            for (unsigned q = 0; q < nthr; q++) std::fwrite(t->part[q][o].data(), 1, t->part[q][o].size(), fp);
            std::fclose(fp);
          });
        }
      }
      for (auto& w : wr) w.join();
      std::cout << "processed events: " << t->done << " / " << count << "\r" << std::flush;
      { std::lock_guard<std::mutex> l(m); tempty.push_back(t); } cv.notify_all();
    }
  });
  // stage 2: rows -> text
  std::thread formatter([&] {
    for (;;) {
      Chunk* c; Text* t;
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return finished || !full.empty(); }); if (full.empty()) break; c = full.front(); full.pop_front(); }
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !tempty.empty(); }); t = tempty.front(); tempty.pop_front(); }
      std::vector<std::thread> th;
      std::vector<long> bad(nthr, 0);
      for (unsigned q = 0; q < nthr; q++) th.emplace_back([&, q] {
        const int a = (int)((long)c->n * ny * q / nthr), b = (int)((long)c->n * ny * (q + 1) / nthr);
        auto& P = t->part[q];
        for (auto& r : P) r.clear();
        for (int o = lo; o <= hi; o++) P[o].reserve((size_t)(b - a) * 200);
        P[10].reserve((size_t)(b - a) * 860);
        RowCells rc;
        for (int e = a; e < b; e++) {
          if (c->out[e].status != SMC_OK) { bad[q]++; continue; }
          format_cells(c->out[e], deformed, rc);
          for (int n = 1; n <= 9; n++) {
            const bool own = n >= lo && n <= hi;
            for (int k = 0; k < 5; k++) {
              const char* s = rc.mom[(n - 1) * 5 + k]; const size_t L = (size_t)rc.mlen[(n - 1) * 5 + k];
              if (own) P[n].append(s, L);
              P[10].append(s, L);
            }
            if (own) P[n].append(rc.tail, (size_t)rc.tlen);
          }
          P[10].append(rc.tail, (size_t)rc.tlen);
        }
      });
      for (auto& x : th) x.join();
      for (long v : bad) failed += v;
      if (binary) { FILE* fp = std::fopen(path("ecc_rows.bin").c_str(), "ab"); if (fp) { std::fwrite(c->out.data(), sizeof(smc_event_out), (size_t)c->n * ny, fp); std::fclose(fp); } }
      t->done = c->done;
      { std::lock_guard<std::mutex> l(m); empty.push_back(c); tfull.push_back(t); } cv.notify_all();
    }
    { std::lock_guard<std::mutex> l(m); formatted = true; } cv.notify_all();
  });
  int rc = 0;
  for (int done = 0; done < count && !rc; done += chunk) {
    const int n = std::min(chunk, count - done);
    Chunk* c;
    { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !empty.empty(); }); c = empty.front(); empty.pop_front(); }
    if (smc_run_events(ctx, first + done, n, SMC_RUN_MOMENTS, c->out.data()) != SMC_OK) { err = smc_last_error(ctx); rc = 1; n_failed_events = -1; { std::lock_guard<std::mutex> l(m); empty.push_back(c); } break; }
    c->n = n; c->done = done + n;
    { std::lock_guard<std::mutex> l(m); full.push_back(c); } cv.notify_all();
  }
  { std::lock_guard<std::mutex> l(m); finished = true; } cv.notify_all();
  formatter.join(); writer.join();
  std::cout << std::endl;
  if (failed) {      // an event whose collision list overflowed ncoll_cap has no row: say so loudly, the table is short
    std::cerr << "superMC_b200: " << failed << " event(s) exceeded the per-event capacities (status != 0) and have no row; raise ncoll_cap" << std::endl;
    n_failed_events = failed;
  }
  if (!rc && shard.world > 1) rc = merge_rank_outputs();
  return rc;
}

// several ranks: every rank wrote its own directory (rank 0: data/, rank r: data_rank<r>/).  After a barrier rank 0 appends
// the tables in rank order -- the result is byte-identical to a 1-GPU run because event k's row does not depend on which
// GPU computed it -- and moves per-event files over.
int MakeDensity::merge_rank_outputs() {
  if (smc_comm_barrier(ctx) != SMC_OK) { err = smc_last_error(ctx); return 1; }
  if (shard.rank == 0) {
    const std::string root = root_data_dir;
    for (int r = 1; r < shard.world; r++) {
      const std::string rd = root + "_rank" + std::to_string(r);
      DIR* d = opendir(rd.c_str());
      if (!d) continue;
      std::vector<std::string> names;
      while (dirent* de = readdir(d)) { const std::string n = de->d_name; if (n != "." && n != "..") names.push_back(n); }
      closedir(d);
      std::sort(names.begin(), names.end());
      for (const std::string& n : names) {
        const std::string src = rd + "/" + n, dst = root + "/" + n;
        const bool per_event = n.find("_event_") != std::string::npos;
        const bool last_wins = (n == "wounded.data" || n == "nucl1.data" || n == "nucl2.data");      // rewritten per event: the last event is on the last rank
        if (per_event || last_wins) { std::rename(src.c_str(), dst.c_str()); continue; }
        if (n == "dNdyTable.dat") { std::remove(src.c_str()); continue; }                             // every rank builds the same table
        FILE* in = std::fopen(src.c_str(), "rb"); FILE* out = std::fopen(dst.c_str(), "ab");
        if (in && out) { char b[1 << 16]; size_t k; while ((k = std::fread(b, 1, sizeof b, in)) > 0) std::fwrite(b, 1, k, out); }
        if (in) std::fclose(in);
        if (out) std::fclose(out);
        std::remove(src.c_str());
      }
      rmdir(rd.c_str());
    }
  }
  if (smc_comm_barrier(ctx) != SMC_OK) { err = smc_last_error(ctx); return 1; }
  return 0;
}

// ---- list dumps (MCnucl::dumpBinaryTable / dumpparticipantTable / dumpSpectatorsTable, MCnucl.cpp:1177-1269) ----
std::string smc_fmt_xy(const double* rows, int n, int stride) {        // setprecision(3) setw(10) x2
  std::string s; s.reserve((size_t)n * 21);
  for (int i = 0; i < n; i++) { put_g(s, rows[(size_t)i * stride], 10, 3); put_g(s, rows[(size_t)i * stride + 1], 10, 3); s += "\n"; } return s;
}
std::string smc_fmt_participants(const double* rows, int n) {          // Nucleus::dumpParticipants, Nucleus.cpp:753-764
  std::string s; s.reserve((size_t)n * 28);
  for (int i = 0; i < n; i++) { put_g(s, rows[(size_t)i * 8], 10, 3); s += "   "; put_g(s, rows[(size_t)i * 8 + 1], 10, 3); s += "   "; s += ((int)rows[(size_t)i * 8 + 2] == 1 ? "1\n" : "2\n"); }
  return s;
}
std::string smc_fmt_quarks(const double* rows, int n) {                // Nucleus::dumpQuarks, Nucleus.cpp:780-797: x y at precision 3, the box at the default 6
  std::string s; s.reserve((size_t)n * 64);
  for (int i = 0; i < n; i++) {
    const double* r = rows + (size_t)i * 6;
    put_g(s, r[0], 10, 3); put_g(s, r[1], 10, 3);
    for (int k = 2; k < 6; k++) { s += ' '; put_g(s, r[k], 0, 6); }
    s += "\n";
  }
  return s;
}
std::string smc_fmt_spectators(const double* rows, int n) {            // scientific, setprecision(4), setw(10) on x only
  std::string s; s.reserve((size_t)n * 36);
  auto sci = [&](double v, int width) { char b[40]; const auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::scientific, 4); const int k = (int)(r.ptr - b); if (k < width) s.append((size_t)(width - k), ' '); s.append(b, (size_t)k); };
  for (int i = 0; i < n; i++) { sci(rows[(size_t)i * 3], 10); s += "  "; sci(rows[(size_t)i * 3 + 1], 0); s += "  "; sci(rows[(size_t)i * 3 + 2], 0); s += "\n"; }
  return s;
}

// shared body of operations 1 and 2
static int ebe_common(MakeDensity* self, smc_ctx* ctx, ParameterReader* paraRdr, bool jet, int nevent, uint64_t first, int count,
                      const std::string& data_dir, int Maxx, int Maxy, double Xmin, double Ymin, double dx, double dy, double rapMin,
                      bool deformed, int batch, std::string& err) {
  const bool use_sd = paraRdr->getVal("use_sd") != 0, use_ed = paraRdr->getVal("use_ed") != 0;
  const bool use_block = paraRdr->getVal("use_block") != 0, use_4col = paraRdr->getVal("use_4col") != 0;
  const bool o_rb = jet && paraRdr->getVal("output_rho_binary") != 0, o_ta = jet && paraRdr->getVal("output_TA") != 0;
  const bool o_rhob = jet && paraRdr->getVal("output_rhob") != 0, o_sp = jet && paraRdr->getVal("output_spectator_density") != 0;
  const int from_order = (int)paraRdr->getVal("ecc_from_order"), to_order = (int)paraRdr->getVal("ecc_to_order");
  unsigned flags = SMC_RUN_MOMENTS | SMC_RUN_KEEP_RHO | SMC_RUN_LISTS;
  if (o_ta || o_rhob) flags |= SMC_RUN_THICKNESS;
  if (o_rb) flags |= SMC_RUN_RHO_BINARY;
  if (o_sp) flags |= SMC_RUN_SPECTATORS;
  const size_t G = (size_t)Maxx * Maxy;
  const bool binary = paraRdr->getVal("output_binary", 0) != 0;          // extension: raw float64 lattices (<stem>.bin) instead of text
  WriterPool pool(std::max(2u, std::min(64u, std::thread::hardware_concurrency())));
  const int ny = std::max(1, self->params.ny);
  std::vector<smc_event_out> out_all((size_t)batch * ny);
  // the reference writes every rapidity slice to the same per-event file names (MakeDensity.cpp:311-445): the files hold
  // the last slice, the eccentricity files one row per slice
  rapMin = rapMin + (-rapMin - rapMin) / ny * (ny - 1);
  auto P = [&](const char* fmt, long ev) { char b[160]; std::snprintf(b, sizeof b, fmt, ev); return data_dir + "/" + b; };
  // one page-locked block per (batch, grid kind), filled by ONE strided device->host copy and shared by the writer jobs
  // (page-locking 140 MB takes tens of milliseconds: the blocks go back to a free list when the last writer job lets go)
  struct PinBuf { double* p; size_t n; explicit PinBuf(size_t n_) : p((double*)smc_pinned_alloc(n_ * sizeof(double))), n(n_) {} ~PinBuf() { smc_pinned_free(p); } };
  typedef std::shared_ptr<PinBuf> Buf;
  struct PinPool { std::mutex m; std::vector<PinBuf*> free_list; ~PinPool() { for (PinBuf* b : free_list) delete b; } };
  auto pins = std::make_shared<PinPool>();
  auto pin_get = [pins](size_t n) -> Buf {
    PinBuf* b = nullptr;
    { std::lock_guard<std::mutex> l(pins->m);
      for (size_t i = 0; i < pins->free_list.size(); i++) if (pins->free_list[i]->n >= n) { b = pins->free_list[i]; pins->free_list.erase(pins->free_list.begin() + (long)i); break; } }
    if (!b) b = new PinBuf(n);
    return Buf(b, [pins](PinBuf* q) { std::lock_guard<std::mutex> l(pins->m); pins->free_list.push_back(q); });
  };
  // stem2 (optional): a second file set with the same numbers (sd and ed are identical, quirk Q2): formatted once, written twice
  auto grid_job = [&](Buf keep, const double* g, const std::string& stem, double npart, const std::string& stem2 = std::string()) {
    auto both = [=](const char* ext, const std::string& text) { write_file(stem + ext, text, false); if (!stem2.empty()) write_file(stem2 + ext, text, false); };
    if (binary) { pool.submit([=] { (void)keep; both(".bin", std::string((const char*)g, G * sizeof(double))); }); return; }
    if (use_4col) pool.submit([=] { (void)keep; std::string s; MakeDensity::formatDensity4Col(g, Maxx, Maxy, Xmin, Ymin, dx, dy, rapMin, npart, s); both("_4col.dat", s); });
    if (use_block) pool.submit([=] { (void)keep; std::string s; MakeDensity::formatDensityBlock(g, Maxx, Maxy, s); both("_block.dat", s); });
  };
  auto fetch_all = [&](int n, int which, double scale) -> Buf {
    Buf b = pin_get((size_t)n * G);
    if (!b->p) { err = "smc_pinned_alloc failed"; return Buf(); }
    if (smc_get_grids(ctx, 0, n, which, b->p) != SMC_OK) { err = smc_last_error(ctx); return Buf(); }
    if (scale != 1.0) for (size_t q = 0; q < (size_t)n * G; q++) b->p[q] *= scale;
    return b;
  };
  const double ff = self->params.finalfactor;
  batch = std::max(1, std::min(batch, smc_max_batch(ctx)));               // the getters address one device batch
  auto ck = [&](int rc) { if (rc != SMC_OK && err.empty()) err = smc_last_error(ctx); return rc != SMC_OK; };
  std::vector<double> last_part, last_nu[2]; bool have_last = false;      // the event wounded.data / nucl1.data / nucl2.data end up holding
  const bool timing = std::getenv("SMC_TIMING") != nullptr;
  double t_run = 0, t_fetch = 0, t_lists = 0; auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  for (int done = 0; done < count; done += batch) {
    const int n = std::min(batch, count - done);
    const auto t0 = now();
    if (smc_run_events(ctx, first + done, n, flags, out_all.data()) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    const auto t1 = now(); t_run += secs(t0, t1);
    Buf b_rho, b_rb, b_ta, b_tb, b_sum, b_sa, b_sb;
    if ((use_sd || use_ed) && !(b_rho = fetch_all(n, SMC_GRID_RHO, ff))) return 1;
    if (o_rb && !(b_rb = fetch_all(n, SMC_GRID_RHO_BINARY, 1.0))) return 1;
    if (o_ta || o_rhob) {
      if (!(b_ta = fetch_all(n, SMC_GRID_TA1, 1.0)) || !(b_tb = fetch_all(n, SMC_GRID_TA2, 1.0))) return 1;
      if (o_rhob) { b_sum = pin_get((size_t)n * G); if (!b_sum->p) { err = "smc_pinned_alloc failed"; return 1; } for (size_t q = 0; q < (size_t)n * G; q++) b_sum->p[q] = b_ta->p[q] + b_tb->p[q]; }
    }
    if (o_sp && (!(b_sa = fetch_all(n, SMC_GRID_SPEC_A, 1.0)) || !(b_sb = fetch_all(n, SMC_GRID_SPEC_B, 1.0)))) return 1;
    const auto t2 = now(); t_fetch += secs(t1, t2);
    // The per-event lists (participants, collisions, quarks, nucleons -> text) are independent of each other: a few threads
    // walk the batch (the getters only read the host mirror of the batch's records), the main thread then appends the
    // blocks of binary.dat / quarks.data in event order, once per batch.
    std::vector<std::string> bin_s(n), quark_s(n);
    int e_last = -1; for (int e = 0; e < n; e++) if (out_all[(size_t)e * ny].status == SMC_OK) e_last = e;
    { int np0 = 0; if (n > 0 && e_last >= 0 && ck(smc_get_participants(ctx, e_last, nullptr, &np0))) return 1; }      // fills the mirror before the threads start
    const unsigned nlt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    std::vector<int> lerr(nlt, 0);
    std::vector<std::thread> lth;
    for (unsigned q = 0; q < nlt; q++) lth.emplace_back([&, q] {
      auto bad = [&](int rc) { if (rc != SMC_OK) lerr[q] = 1; return rc != SMC_OK; };
      for (int e = (int)q; e < n; e += (int)nlt) {
        const long event = (long)(first + done + e) + 1;                   // the reference counts events from 1
        const smc_event_out* out = out_all.data() + (size_t)e * ny - e;     // out[e] = slice 0 of event e
        if (out[e].status != SMC_OK) continue;
        int np = 0, nc = 0, ns = 0;
        if (bad(smc_get_participants(ctx, e, nullptr, &np)) || bad(smc_get_collisions(ctx, e, nullptr, &nc))) return;
        std::vector<double> part((size_t)std::max(np, 1) * 8), coll((size_t)std::max(nc, 1) * 6);
        if (bad(smc_get_participants(ctx, e, part.data(), &np)) || bad(smc_get_collisions(ctx, e, coll.data(), &nc))) return;
        if (jet) {
          if (bad(smc_get_spectators(ctx, e, nullptr, &ns))) return;
          std::vector<double> spec((size_t)std::max(ns, 1) * 3); if (bad(smc_get_spectators(ctx, e, spec.data(), &ns))) return;
          write_file(P("ParticipantTable_event_%ld.dat", event), smc_fmt_participants(part.data(), np), true);
          write_file(P("Spectators_event_%ld.dat", event), smc_fmt_spectators(spec.data(), ns), false);
          write_file(P("BinaryCollisionTable_event_%ld.dat", event), smc_fmt_xy(coll.data(), nc, 6), true);
        } else bin_s[e] = smc_fmt_xy(coll.data(), nc, 6);
        { int nq = 0; if (bad(smc_get_quarks(ctx, e, nullptr, &nq))) return;
          std::vector<double> qk((size_t)std::max(nq, 1) * 6); if (bad(smc_get_quarks(ctx, e, qk.data(), &nq))) return;
          quark_s[e] = smc_fmt_quarks(qk.data(), nq); }
        // dumpBinaryTable side files (MCnucl.cpp:1193-1209).  wounded.data, nucl1.data and nucl2.data are rewritten by every
        // event: what remains is the last event's, the only one kept here (written after the loop)
        if (e == e_last) {
          last_part.assign(part.begin(), part.begin() + (size_t)np * 8);
          for (int sd = 0; sd < 2; sd++) {
            int na = 0; if (bad(smc_get_nucleons(ctx, e, sd, nullptr, &na))) return;
            last_nu[sd].resize((size_t)std::max(na, 1) * 8); if (bad(smc_get_nucleons(ctx, e, sd, last_nu[sd].data(), &na))) return;
            last_nu[sd].resize((size_t)na * 8);
          }
          have_last = true;
        }
        if (jet) for (int f = 0; f < 2; f++) {                              // per-event eccentricity files (:327-345)
          if ((f == 0 && !use_sd) || (f == 1 && !use_ed)) continue;
          char nm[160];
          for (int o = std::max(from_order, 1); o <= std::min(to_order, 9); o++) {
            std::snprintf(nm, sizeof nm, "%s_ecc_eccp_%d_event_%ld.dat", f == 0 ? "sn" : "en", o, event);
            std::string rows; for (int iy = 0; iy < ny; iy++) rows += MakeDensity::formatEccRow(out[e + iy], o, deformed);
            write_file(data_dir + "/" + nm, rows, true);
          }
          std::snprintf(nm, sizeof nm, "%s_ecc_eccp_%d_event_%ld.dat", f == 0 ? "sn" : "en", 10, event);
          { std::string rows; for (int iy = 0; iy < ny; iy++) rows += MakeDensity::formatEccRowAll(out[e + iy], deformed); write_file(data_dir + "/" + nm, rows, true); }
        }
      }
    });
    for (auto& t : lth) t.join();
    for (int v : lerr) if (v) { if (err.empty()) err = smc_last_error(ctx); return 1; }
    std::string app_binary, app_quarks;
    for (int e = 0; e < n; e++) { app_binary += bin_s[e]; app_quarks += quark_s[e]; }
    for (int e = 0; e < n; e++) {
      const long event = (long)(first + done + e) + 1;
      const smc_event_out* out = out_all.data() + (size_t)e * ny - e;
      if (out[e].status != SMC_OK) { std::cerr << "event " << event << ": status " << out[e].status << std::endl; continue; }
      const double npart = out[e].npart1 + out[e].npart2;
      if (use_sd || use_ed) {
        const double* g = b_rho->p + (size_t)e * G;
        if (use_sd && use_ed) grid_job(b_rho, g, P("sd_event_%ld", event), npart, P("ed_event_%ld", event));   // identical numbers (quirk Q2)
        else if (use_sd) grid_job(b_rho, g, P("sd_event_%ld", event), npart);
        else grid_job(b_rho, g, P("ed_event_%ld", event), npart);
      }
      if (o_rb) grid_job(b_rb, b_rb->p + (size_t)e * G, P("rho_binary_event_%ld", event), npart);
      if (o_ta) { grid_job(b_ta, b_ta->p + (size_t)e * G, P("nuclear_thickness_TA_event_%ld", event), npart); grid_job(b_tb, b_tb->p + (size_t)e * G, P("nuclear_thickness_TB_event_%ld", event), npart); }
      if (o_rhob) grid_job(b_sum, b_sum->p + (size_t)e * G, P("rhob_event_%ld", event), npart);
      if (o_sp) { grid_job(b_sa, b_sa->p + (size_t)e * G, P("spectator_density_A_event_%ld", event), npart); grid_job(b_sb, b_sb->p + (size_t)e * G, P("spectator_density_B_event_%ld", event), npart); }
    }
    if (!app_binary.empty()) write_file(data_dir + "/binary.dat", app_binary, true);
    if (!app_quarks.empty()) write_file(data_dir + "/quarks.data", app_quarks, true);
    t_lists += secs(t2, now());
  }
  if (timing) std::cerr << "# operation 1/2 main thread: smc_run_events " << t_run << " s, lattices device->host " << t_fetch
                        << " s, lists + job hand-over (blocks when the writers lag) " << t_lists << " s" << std::endl;
  if (have_last) {
    write_file(data_dir + "/wounded.data", smc_fmt_participants(last_part.data(), (int)(last_part.size() / 8)), false);
    for (int s = 0; s < 2; s++) write_file(data_dir + (s == 0 ? "/nucl1.data" : "/nucl2.data"), smc_fmt_xy(last_nu[s].data(), (int)(last_nu[s].size() / 8), 8), false);
  }
  pool.wait();
  return 0;
}

int MakeDensity::generate_profile_ebe(int nevent) {
  uint64_t first; int count; shard_range(nevent, &first, &count);
  int rc = ebe_common(this, ctx, paraRdr, false, nevent, first, count, data_dir, Maxx, Maxy, Xmin, Ymin, dx, dy, rapMin, deformed, 256, err);
  if (!rc && shard.world > 1) rc = merge_rank_outputs();
  return rc;
}
int MakeDensity::generate_profile_ebe_Jet(int nevent) {
  uint64_t first; int count; shard_range(nevent, &first, &count);
  int rc = ebe_common(this, ctx, paraRdr, true, nevent, first, count, data_dir, Maxx, Maxy, Xmin, Ymin, dx, dy, rapMin, deformed, 128, err);
  if (!rc && shard.world > 1) rc = merge_rank_outputs();
  return rc;
}

// ---- operation 3: averaged profiles (src/MakeDensity.cpp:736-2103) ----------------------------------
int MakeDensity::average_accumulate(int nevent) {
  const int from = ival(paraRdr, "average_from_order"), to = ival(paraRdr, "average_to_order");
  const int branches = (paraRdr->getVal("use_sd") != 0 ? 1 : 0) | (paraRdr->getVal("use_ed") != 0 ? 2 : 0);
  if (!branches) { err = "operation 3 needs use_sd or use_ed"; return 1; }
  if (smc_avg_begin(ctx, from, to, ival(paraRdr, "generate_reaction_plane_avg_profile") == 1, branches) != SMC_OK) { err = smc_last_error(ctx); return 1; }
  uint64_t first; int count; shard_range(nevent, &first, &count);
  const int chunk = 1024;
  std::vector<smc_event_out> out(chunk);
  for (int done = 0; done < count; done += chunk) {
    const int n = std::min(chunk, count - done);
    if (smc_avg_run(ctx, first + done, n, out.data()) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    last_npart = out[n - 1].npart1 + out[n - 1].npart2;
    std::cout << "processing event: " << done + n << std::endl;          // :1574
  }
  return 0;
}

int MakeDensity::average_write() {
  const int from = ival(paraRdr, "average_from_order"), to = ival(paraRdr, "average_to_order");
  const bool use_sd = paraRdr->getVal("use_sd") != 0, use_ed = paraRdr->getVal("use_ed") != 0;
  const bool use_block = paraRdr->getVal("use_block") != 0, use_4col = paraRdr->getVal("use_4col") != 0;
  const bool rp = ival(paraRdr, "generate_reaction_plane_avg_profile") == 1;
  const bool o_tatb = ival(paraRdr, "output_TATB") == 1, o_rb = ival(paraRdr, "output_rho_binary") == 1;
  const bool o_ta = ival(paraRdr, "output_TA") == 1, o_sp = ival(paraRdr, "output_spectator_density") == 1;
  const size_t G = (size_t)Maxx * Maxy;
  WriterPool pool(std::max(2u, std::min(16u, std::thread::hardware_concurrency())));
  auto emit = [&](int order, int variant, int quantity, int branch, const char* fmt) -> int {
    auto g = std::make_shared<std::vector<double>>(G);
    if (smc_avg_get(ctx, order, variant, quantity, branch, g->data()) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    char stem[200]; std::snprintf(stem, sizeof stem, fmt, order);
    const std::string st = path(stem); const double npart = last_npart;
    const int mx = Maxx, my = Maxy; const double x0 = Xmin, y0 = Ymin, ddx = dx, ddy = dy;
    const double rap = rapMin + (rapMax - rapMin) / binRapidity * (binRapidity - 1);      // the files hold the last slice (:1579-1900)
    if (use_4col) pool.submit([=] { std::string s; MakeDensity::formatDensity4Col(g->data(), mx, my, x0, y0, ddx, ddy, rap, npart, s); write_file(st + "_4col.dat", s, false); });
    if (use_block) pool.submit([=] { std::string s; MakeDensity::formatDensityBlock(g->data(), mx, my, s); write_file(st + "_block.dat", s, false); });
    return 0;
  };
  for (int order = from; order <= to; order++) {
    for (int branch = 0; branch < 2; branch++) {
      if ((branch == 0 && !use_sd) || (branch == 1 && !use_ed)) continue;
      const bool sd = branch == 0;
      if (emit(order, 0, SMC_AVG_SD, branch, sd ? "sdAvg_order_%d" : "edAvg_order_%d")) return 1;
      if (rp && emit(order, 1, SMC_AVG_SD, branch, sd ? "sdAvg_RP_order_%d" : "edAvg_RP_order_%d")) return 1;
      if (o_tatb) {
        if (emit(order, 0, SMC_AVG_TATB, branch, sd ? "TATB_fromSd_order_%d" : "TATB_fromEd_order_%d")) return 1;
        if (rp && emit(order, 1, SMC_AVG_TATB, branch, sd ? "TATB_fromSd_RP_order_%d" : "TATB_fromEd_RP_order_%d")) return 1;
      }
      if (o_rb) {
        if (emit(order, 0, SMC_AVG_RHO_BINARY, branch, sd ? "rho_binary_fromSd_order_%d" : "rho_binary_fromEd_order_%d")) return 1;
        if (rp && emit(order, 1, SMC_AVG_RHO_BINARY, branch, sd ? "rho_binary_fromSd_RP_order_%d" : "rho_binary_fromEd_RP_order_%d")) return 1;
      }
      if (o_ta) {
        if (emit(order, 0, SMC_AVG_TA, branch, sd ? "nuclear_thickness_TA_fromSd_order_%d" : "nuclear_thickness_TA_fromEd_order_%d")) return 1;
        if (emit(order, 0, SMC_AVG_TB, branch, sd ? "nuclear_thickness_TB_fromSd_order_%d" : "nuclear_thickness_TB_fromEd_order_%d")) return 1;
        if (rp) {
          if (emit(order, 1, SMC_AVG_TA, branch, sd ? "nuclear_thickness_TA_fromSd_RP_order_%d" : "nuclear_thickness_TA_fromEd_RP_order_%d")) return 1;
          if (emit(order, 1, SMC_AVG_TB, branch, sd ? "nuclear_thickness_TB_fromSd_RP_order_%d" : "nuclear_thickness_TB_fromEd_RP_order_%d")) return 1;
        }
      }
      if (o_sp) {
        if (emit(order, 0, SMC_AVG_SPEC_A, branch, sd ? "spectator_density_A_fromSd_order_%d" : "spectator_density_A_fromEd_order_%d")) return 1;
        if (emit(order, 0, SMC_AVG_SPEC_B, branch, sd ? "spectator_density_B_fromSd_order_%d" : "spectator_density_B_fromEd_order_%d")) return 1;
      }
    }
  }
  pool.wait();
  return 0;
}

int MakeDensity::generate_profile_average(int nevent) {
  if (average_accumulate(nevent)) return 1;
  if (shard.world > 1) {
    // the only exchange of the run: accumulator sums + accepted-event counts, one all-reduce (NCCL over NVLink, or the
    // peer-memory kernel when ranks share a GPU); the header of the 4-column files carries the LAST event's Npart
    if (smc_avg_allreduce(ctx) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    std::vector<double> all(shard.world, 0.0);
    if (smc_comm_gather_doubles(ctx, &last_npart, 1, all.data(), shard.world, nullptr) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    int rc = 0;
    if (shard.rank == 0) { for (int r = 0; r < shard.world; r++) if (all[r] > 0) last_npart = all[r]; rc = average_write(); }
    if (smc_comm_barrier(ctx) != SMC_OK) { err = smc_last_error(ctx); return 1; }
    return rc;
  }
  return average_write();
}
