// superMC_b200.e -- drop-in for the reference executable (reference src/main.cpp:19-73): reads
// parameters.dat, applies name=value overrides, runs the selected operation on the GPU.
// Multi-GPU: one process per GPU; RANK / WORLD_SIZE / LOCAL_RANK (torchrun convention, e.g.
// `torchrun --no-python --nproc-per-node 8 ./superMC_b200.e ...` or a shell loop) select the shard of global event ids and
// the device; MASTER_ADDR / MASTER_PORT (+1) or SMC_COMM_PORT is where rank 0 listens for the rendezvous.  Operation 3
// all-reduces its accumulators (NCCL over NVLink); the tables and per-event files of the ranks are merged by rank 0.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <unistd.h>
#include "MakeDensity.h"

int main(int argc, char* argv[]) {
  const auto t_start = std::chrono::steady_clock::now();
  const bool timing = std::getenv("SMC_TIMING") != nullptr;      // phase times on stderr (start-up is CUDA initialisation + memory pools)
  ParameterReader paraRdr;
  try {
    paraRdr.readFromFile("parameters.dat");
    paraRdr.readFromArguments(argc, argv);
  } catch (std::exception& e) { std::cout << e.what() << std::endl; return 255; }
  paraRdr.echo();
  const char* er = std::getenv("RANK"); const char* ew = std::getenv("WORLD_SIZE"); const char* el = std::getenv("LOCAL_RANK");
  smc_shard sh{er ? std::atoi(er) : 0, ew ? std::atoi(ew) : 1};
  const int device = el ? std::atoi(el) : 0;
  MakeDensity dens(&paraRdr, device, sh, "data");      // rank r > 0 writes data_rank<r>/, merged into data/ by rank 0
  if (!dens.ok()) { std::cerr << "superMC_b200: " << dens.error() << std::endl; return 255; }
  if (timing) std::cerr << "# start-up (parameters, CUDA context, tables): " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() << " s" << std::endl;
  int nevent = 0, operation = 0;
  try { nevent = (int)paraRdr.getVal("nev"); operation = (int)paraRdr.getVal("operation"); }
  catch (std::exception& e) { std::cout << e.what() << std::endl; return 255; }
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = dens.run(operation, nevent);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc) std::cerr << "superMC_b200: " << dens.error() << std::endl;
  std::cout << "Time elapsed (in seconds): " << dt << std::endl;
  if (timing) std::cerr << "# event loop: " << dt << " s; since process start: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() << " s (context teardown follows)" << std::endl;
  // every file is closed and every rank has passed its last barrier: returning through the destructors would only hand
  // several GB of device memory back piece by piece (0.3-0.6 s); the driver reclaims it with the process
  const char* fe = std::getenv("SMC_FAST_EXIT");
  if (!(fe && fe[0] == '0')) { std::cout.flush(); std::cerr.flush(); std::fflush(nullptr); _exit(rc); }
  return rc;
}
