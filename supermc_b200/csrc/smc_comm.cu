// smc_comm.cu -- one process per GPU: rendezvous, barrier and the two exchanges the path has.
//
// The reference's only parallel mode is "start 8 copies with different seeds" (CollectDataAccordingToSettings.py:110-115);
// nothing is ever combined.  Here the ranks of one node own contiguous event-id ranges of ONE run, and two things cross
// GPUs:
//   * operation 3: the averaged-profile accumulator sums (<= ~30 MB of FP64 per GPU) + the accepted-event counter, once
//     per run -> smc_avg_allreduce;
//   * per-event scalar rows for the centrality sort -> smc_comm_gather_doubles (rank 0 receives).
// Plumbing: a TCP star on MASTER_ADDR (rank 0 listens) carries the small control blobs (ncclUniqueId, CUDA IPC handles,
// counters, barriers).  Data path of the all-reduce:
//   * "nccl": ncclAllReduce on the library's stream over NVLink / NVSwitch.  libnccl.so.2 is opened at run time, so the
//     library has no link-time dependency and shares the copy a host process (e.g. torch) has already loaded;
//   * "ipc": every rank maps the peers' accumulator blocks through CUDA IPC and ONE kernel per rank sums its slice of all
//     blocks straight out of peer memory (P2P loads over NVLink, fixed rank order => the result does not depend on
//     timing), then the reduced slices are copied back peer to peer.  Used when ranks share a device (NCCL refuses
//     duplicate GPUs -- that is how the 2-rank path is tested on a 1-GPU box) or when libnccl is absent.
// Every wait has a timeout and returns SMC_ERR_STATE instead of hanging.
#include <arpa/inet.h>
#include <dlfcn.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>
#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "smc_ctx.h"

namespace {

// ---- the few NCCL declarations needed (nccl.h is not included: the library is resolved with dlopen) ----
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
    CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};

struct Comm {
  int rank = 0, world = 1;
  int listen_fd = -1;
  std::vector<int> fds;          // rank 0: socket of rank r at [r]; others: [0] = socket to rank 0
  bool nccl = false; NcclApi api; NcclComm comm = nullptr;
  double* scratch = nullptr; size_t scratch_doubles = 0;
  double last_allreduce_ms = 0.0;
  std::string backend;
};

const int kTimeoutS = 300;

bool send_all(int fd, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) { ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL); if (k <= 0) { if (errno == EINTR) continue; return false; } c += k; n -= (size_t)k; }
  return true;
}
bool recv_all(int fd, void* p, size_t n) {
  char* c = (char*)p;
  while (n) { ssize_t k = ::recv(fd, c, n, 0); if (k <= 0) { if (k < 0 && errno == EINTR) continue; return false; } c += k; n -= (size_t)k; }
  return true;
}
void set_timeouts(int fd) {
  timeval tv; tv.tv_sec = kTimeoutS; tv.tv_usec = 0;
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv); setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
  int one = 1; setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
}

// every rank contributes n bytes; every rank receives world * n bytes in rank order (star through rank 0)
bool allgather(Comm* c, const void* mine, size_t n, std::vector<char>& all) {
  all.assign((size_t)c->world * n, 0);
  if (c->world == 1) { std::memcpy(all.data(), mine, n); return true; }
  if (c->rank == 0) {
    std::memcpy(all.data(), mine, n);
    for (int r = 1; r < c->world; r++) if (!recv_all(c->fds[r], all.data() + (size_t)r * n, n)) return false;
    for (int r = 1; r < c->world; r++) if (!send_all(c->fds[r], all.data(), all.size())) return false;
    return true;
  }
  return send_all(c->fds[0], mine, n) && recv_all(c->fds[0], all.data(), all.size());
}
bool barrier(Comm* c) { char b = 1; std::vector<char> all; return allgather(c, &b, 1, all); }

struct PeerPtrs { const double* p[16]; int n; };
// rank r's share of the all-reduce: out[i - lo] = sum over ranks (fixed order) of block_q[i], read from peer memory
__global__ void reduce_slice_kernel(PeerPtrs pp, size_t lo, size_t hi, double* out) {
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < pp.n; q++) s += pp.p[q][i];
    out[i - lo] = s;
  }
}

Comm* comm_of(smc_ctx* ctx) { return (Comm*)ctx->comm; }

}  // namespace

extern "C" int smc_comm_init(smc_ctx* ctx, int rank, int world, const char* addr, int port) {
  if (!ctx || world < 1 || rank < 0 || rank >= world || world > 16) return SMC_ERR_PARAM;
  if (ctx->comm) FAIL(SMC_ERR_STATE, "smc_comm_init: already initialised");
  Comm* c = new Comm(); c->rank = rank; c->world = world; ctx->comm = c; c->backend = "single";
  if (world == 1) return SMC_OK;
  CK(cudaSetDevice(ctx->device));
  const std::string host = (addr && *addr) ? addr : "127.0.0.1";
  sockaddr_in sa; std::memset(&sa, 0, sizeof sa); sa.sin_family = AF_INET; sa.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, host.c_str(), &sa.sin_addr) != 1) FAIL(SMC_ERR_PARAM, "smc_comm_init: address must be a dotted IPv4 address (use 127.0.0.1 on one node)");
  if (rank == 0) {
    c->listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
    int one = 1; setsockopt(c->listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    if (::bind(c->listen_fd, (sockaddr*)&sa, sizeof sa) != 0 || ::listen(c->listen_fd, world) != 0)
      FAIL(SMC_ERR_STATE, std::string("smc_comm_init: cannot listen on ") + host + ":" + std::to_string(port) + ": " + std::strerror(errno));
    timeval tv; tv.tv_sec = kTimeoutS; tv.tv_usec = 0; setsockopt(c->listen_fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    c->fds.assign(world, -1);
    for (int k = 1; k < world; k++) {
      int fd = ::accept(c->listen_fd, nullptr, nullptr);
      if (fd < 0) FAIL(SMC_ERR_STATE, "smc_comm_init: timed out waiting for the other ranks");
      set_timeouts(fd);
      int32_t r = -1;
      if (!recv_all(fd, &r, sizeof r) || r < 1 || r >= world || c->fds[r] >= 0) { ::close(fd); FAIL(SMC_ERR_STATE, "smc_comm_init: bad hello from a rank"); }
      c->fds[r] = fd;
    }
  } else {
    int fd = -1;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      if (::connect(fd, (sockaddr*)&sa, sizeof sa) == 0) break;
      ::close(fd); fd = -1;
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > kTimeoutS) FAIL(SMC_ERR_STATE, "smc_comm_init: rank 0 is not listening");
      std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    set_timeouts(fd);
    int32_t r = rank;
    if (!send_all(fd, &r, sizeof r)) FAIL(SMC_ERR_STATE, "smc_comm_init: hello failed");
    c->fds.assign(1, fd);
  }
  // which data path: NCCL needs one distinct GPU per rank
  char bus[64]; std::memset(bus, 0, sizeof bus); cudaDeviceGetPCIBusId(bus, sizeof bus, ctx->device);
  std::vector<char> all;
  if (!allgather(c, bus, sizeof bus, all)) FAIL(SMC_ERR_STATE, "smc_comm_init: exchange failed");
  bool distinct = true;
  for (int a = 0; a < world; a++) for (int b = a + 1; b < world; b++) if (!std::strncmp(all.data() + (size_t)a * 64, all.data() + (size_t)b * 64, 64)) distinct = false;
  const char* force = getenv("SMC_COMM_BACKEND");
  int want_nccl = (distinct && !(force && !std::strcmp(force, "ipc")) && c->api.load()) ? 1 : 0;
  if (force && !std::strcmp(force, "nccl") && !want_nccl) FAIL(SMC_ERR_STATE, "SMC_COMM_BACKEND=nccl: needs libnccl.so.2 and one distinct GPU per rank");
  int32_t w = want_nccl;                       // all ranks must agree (libnccl might be missing on one)
  if (!allgather(c, &w, sizeof w, all)) FAIL(SMC_ERR_STATE, "smc_comm_init: exchange failed");
  for (int r = 0; r < world; r++) { int32_t v; std::memcpy(&v, all.data() + (size_t)r * sizeof v, sizeof v); if (!v) want_nccl = 0; }
  if (want_nccl) {
    NcclUniqueId id; std::memset(&id, 0, sizeof id);
    if (rank == 0 && c->api.GetUniqueId(&id) != kNcclSuccess) FAIL(SMC_ERR_STATE, "ncclGetUniqueId failed");
    if (!allgather(c, &id, sizeof id, all)) FAIL(SMC_ERR_STATE, "smc_comm_init: exchange failed");
    std::memcpy(&id, all.data(), sizeof id);
    const int rc = c->api.CommInitRank(&c->comm, world, id, rank);
    if (rc != kNcclSuccess) FAIL(SMC_ERR_STATE, std::string("ncclCommInitRank: ") + (c->api.GetErrorString ? c->api.GetErrorString(rc) : "error"));
    c->nccl = true; c->backend = "nccl";
  } else c->backend = "ipc";
  return SMC_OK;
}

extern "C" void smc_comm_finalize(smc_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  Comm* c = comm_of(ctx);
  if (c->comm) c->api.CommDestroy(c->comm);
  if (c->scratch) cudaFree(c->scratch);
  for (int fd : c->fds) if (fd >= 0) ::close(fd);
  if (c->listen_fd >= 0) ::close(c->listen_fd);
  delete c; ctx->comm = nullptr;
}

extern "C" const char* smc_comm_backend(const smc_ctx* ctx) { return (ctx && ctx->comm) ? ((Comm*)ctx->comm)->backend.c_str() : "none"; }
extern "C" double smc_comm_last_allreduce_ms(const smc_ctx* ctx) { return (ctx && ctx->comm) ? ((Comm*)ctx->comm)->last_allreduce_ms : 0.0; }

extern "C" int smc_comm_barrier(smc_ctx* ctx) {
  if (!ctx || !ctx->comm) return SMC_ERR_STATE;
  CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream));
  if (!barrier(comm_of(ctx))) FAIL(SMC_ERR_STATE, "smc_comm_barrier: a rank went away");
  return SMC_OK;
}

// rank 0's value reaches every rank (e.g. the clock-derived seed of randomSeed < 0, src/main.cpp:28-31)
extern "C" int smc_comm_bcast_i64(smc_ctx* ctx, int64_t* v) {
  if (!ctx || !ctx->comm || !v) return SMC_ERR_STATE;
  std::vector<char> all;
  if (!allgather(comm_of(ctx), v, sizeof *v, all)) FAIL(SMC_ERR_STATE, "smc_comm_bcast_i64: a rank went away");
  std::memcpy(v, all.data(), sizeof *v);
  return SMC_OK;
}

extern "C" int smc_comm_gather_doubles(smc_ctx* ctx, const double* mine, int64_t n, double* all, int64_t cap, int64_t* n_per_rank) {
  if (!ctx || !ctx->comm || n < 0 || (n > 0 && !mine)) return SMC_ERR_PARAM;
  Comm* c = comm_of(ctx);
  std::vector<char> cnt;
  if (!allgather(c, &n, sizeof n, cnt)) FAIL(SMC_ERR_STATE, "smc_comm_gather_doubles: a rank went away");
  std::vector<int64_t> counts(c->world);
  std::memcpy(counts.data(), cnt.data(), cnt.size());
  if (n_per_rank) for (int r = 0; r < c->world; r++) n_per_rank[r] = counts[r];
  if (c->rank == 0) {
    int64_t tot = 0; for (int64_t v : counts) tot += v;
    if (!all || cap < tot) FAIL(SMC_ERR_PARAM, "smc_comm_gather_doubles: receive buffer too small");
    std::memcpy(all, mine, (size_t)n * sizeof(double));
    int64_t off = n;
    for (int r = 1; r < c->world; r++) { if (!recv_all(c->fds[r], all + off, (size_t)counts[r] * sizeof(double))) FAIL(SMC_ERR_STATE, "smc_comm_gather_doubles: a rank went away"); off += counts[r]; }
  } else if (n > 0 && !send_all(c->fds[0], mine, (size_t)n * sizeof(double))) FAIL(SMC_ERR_STATE, "smc_comm_gather_doubles: rank 0 went away");
  return SMC_OK;
}

// sum of the accumulator blocks (smc_avg_device_buffer) and of the accepted-event counters over all ranks, in place
extern "C" int smc_avg_allreduce(smc_ctx* ctx) {
  if (!ctx) return SMC_ERR_PARAM;
  if (!ctx->d_avg) FAIL(SMC_ERR_STATE, "smc_avg_begin first");
  if (!ctx->comm) FAIL(SMC_ERR_STATE, "smc_comm_init first");
  Comm* c = comm_of(ctx);
  if (c->world == 1) return SMC_OK;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->avg_doubles;
  std::vector<char> all;
  int64_t cnt = ctx->avg_count;
  if (!allgather(c, &cnt, sizeof cnt, all)) FAIL(SMC_ERR_STATE, "smc_avg_allreduce: a rank went away");
  int64_t total = 0;
  for (int r = 0; r < c->world; r++) { int64_t v; std::memcpy(&v, all.data() + (size_t)r * sizeof v, sizeof v); total += v; }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  if (c->nccl) {
    const int rc = c->api.AllReduce(ctx->d_avg, ctx->d_avg, n, kNcclFloat64, kNcclSum, c->comm, ctx->stream);
    if (rc != kNcclSuccess) FAIL(SMC_ERR_STATE, std::string("ncclAllReduce: ") + (c->api.GetErrorString ? c->api.GetErrorString(rc) : "error"));
    CK(cudaEventRecord(ctx->ev1, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
  } else {
    const int W = c->world;
    const size_t per = (n + W - 1) / W, lo = std::min(n, per * c->rank), hi = std::min(n, lo + per);
    if (c->scratch_doubles < per) { if (c->scratch) cudaFree(c->scratch); c->scratch = nullptr; CK(cudaMalloc(&c->scratch, per * sizeof(double))); c->scratch_doubles = per; }
    cudaIpcMemHandle_t mine[2];
    CK(cudaIpcGetMemHandle(&mine[0], ctx->d_avg)); CK(cudaIpcGetMemHandle(&mine[1], c->scratch));
    CK(cudaStreamSynchronize(ctx->stream));                          // the sums of this rank are complete
    if (!allgather(c, mine, sizeof mine, all)) FAIL(SMC_ERR_STATE, "smc_avg_allreduce: a rank went away");   // doubles as the barrier
    std::vector<void*> blk(W, nullptr), scr(W, nullptr);
    PeerPtrs pp; pp.n = W;
    for (int q = 0; q < W; q++) {
      if (q == c->rank) { blk[q] = ctx->d_avg; scr[q] = c->scratch; }
      else {
        cudaIpcMemHandle_t h[2]; std::memcpy(h, all.data() + (size_t)q * sizeof h, sizeof h);
        CK(cudaIpcOpenMemHandle(&blk[q], h[0], cudaIpcMemLazyEnablePeerAccess)); CK(cudaIpcOpenMemHandle(&scr[q], h[1], cudaIpcMemLazyEnablePeerAccess));
      }
      pp.p[q] = (const double*)blk[q];
    }
    if (hi > lo) { reduce_slice_kernel<<<296, 256, 0, ctx->stream>>>(pp, lo, hi, c->scratch); ctx->launches++; CK(cudaGetLastError()); }
    CK(cudaStreamSynchronize(ctx->stream));
    if (!barrier(c)) FAIL(SMC_ERR_STATE, "smc_avg_allreduce: a rank went away");      // every slice is reduced, nobody reads the blocks any more
    for (int q = 0; q < W; q++) {
      const size_t ql = std::min(n, per * q), qh = std::min(n, ql + per);
      if (qh > ql) CK(cudaMemcpyAsync(ctx->d_avg + ql, scr[q], (qh - ql) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    if (!barrier(c)) FAIL(SMC_ERR_STATE, "smc_avg_allreduce: a rank went away");      // the scratch slices have been read
    for (int q = 0; q < W; q++) if (q != c->rank) { cudaIpcCloseMemHandle(blk[q]); cudaIpcCloseMemHandle(scr[q]); }
  }
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); c->last_allreduce_ms = ms; }
  ctx->avg_count = total;
  return SMC_OK;
}
