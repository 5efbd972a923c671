// MakeDensity -- host-side driver with the reference's operation modes (reference src/MakeDensity.h:43-62,
// src/main.cpp:44-64): event-by-event profiles (operation 1 and 2), averaged profiles (operation 3) and the
// minimum-bias eccentricity table (operation 9), writing the reference's data/ file layouts
// (SURVEY.md appendix B).  All per-event physics runs on the GPU behind include/supermc_b200.h; this
// class only batches events, formats text and (operation 3) reads back the averaged grids.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/supermc_b200.h"
#include "ParameterReader.h"

// list writers (MCnucl::dumpBinaryTable / dumpparticipantTable / dumpSpectatorsTable, Nucleus::dumpNucleons)
std::string smc_fmt_xy(const double* rows, int n, int stride);
std::string smc_fmt_participants(const double* rows8, int n);
std::string smc_fmt_spectators(const double* rows3, int n);
std::string smc_fmt_quarks(const double* rows6, int n);

struct smc_shard { int rank, world; };   // events [rank*nev/world, (rank+1)*nev/world) of the global id range

class MakeDensity {
 public:
  MakeDensity(ParameterReader* paraRdr, int device = 0, smc_shard shard = smc_shard{0, 1}, const std::string& data_dir = "data");
  ~MakeDensity();
  bool ok() const { return ctx_ok; }
  const std::string& error() const { return err; }

  int generate_profile_ebe(int nevent);        // operation 1, src/MakeDensity.cpp:502-730
  int generate_profile_ebe_Jet(int nevent);    // operation 2, src/MakeDensity.cpp:148-498
  int generate_profile_average(int nevent);    // operation 3, src/MakeDensity.cpp:736-2103
  int generateEccTable(int nevent);            // operation 9, src/MakeDensity.cpp:2108-2240
  int average_accumulate(int nevent);          // operation 3, event loop only (sums stay on the GPU)
  int average_write();                         // operation 3, the files of src/MakeDensity.cpp:1579-1900
  int run(int operation, int nevent);

  // text writers, byte-compatible with the reference's iostream formatting
  static std::string formatEccRow(const smc_event_out& ev, int order, bool deformed);                 // :2437-2468
  static std::string formatEccRowAll(const smc_event_out& ev, bool deformed);                          // :2473-2499
  static void formatDensityBlock(const double* g, int Maxx, int Maxy, std::string& out);              // :2597-2609
  static void formatDensity4Col(const double* g, int Maxx, int Maxy, double Xmin, double Ymin, double dx, double dy,
                                double rap, double npart, std::string& out);                           // :2573-2595
  smc_ctx* context() { return ctx; }
  smc_params params;

 private:
  int load_tables();
  int merge_rank_outputs();                    // several ranks: rank 0 appends / moves the other ranks' files in rank order
  void shard_range(int nevent, uint64_t* first, int* count) const;
  std::string path(const std::string& name) const { return data_dir + "/" + name; }
  ParameterReader* paraRdr;
  smc_ctx* ctx;
  bool ctx_ok;
  std::string err, data_dir, root_data_dir;
  smc_shard shard;
  smc_constants k;
  int Maxx, Maxy;
  double Xmin, Ymin, dx, dy, rapMin, rapMax, finalFactor;
  int binRapidity;
  bool deformed;
  double last_npart = 0;
 public:
  long n_failed_events = 0;                    // events without a row in the last operation-9 run (capacity overflow), -1 after an error
};
