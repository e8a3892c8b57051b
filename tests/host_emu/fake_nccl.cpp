// TEST-ONLY: an in-process stand-in for the eight NCCL entry points the engine binds with dlopen (msm.cu: NcclApi), for
// the emulated host build of tests/test_host_emu_pipeline.py.  "Ranks" are host threads of one process and "device"
// memory is host memory (cuda_rt_emu.h), so the all-gather is a rendezvous of the ranks' threads and a memcpy.  The test
// points MGB_NCCL_LIB at the library built from this file; nothing in the product links or loads it.
// Fault injection: MGB_FAKE_NCCL_ASYNC_ERROR=1 makes ncclCommGetAsyncError report a failure.
// Ranks in DIFFERENT processes (tests/bench_dry_run.py --ranks 2): with MGB_FAKE_NCCL_DIR set, ncclCommInitRank builds a
// communicator whose all-gather goes through files in that directory (<id>.<sequence>.<rank>, written under a temporary
// name and renamed, so a reader never sees a partial record).
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <unistd.h>

typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclInternalError = 3, ncclInvalidArgument = 4, ncclRemoteError = 6 };
struct ncclUniqueId { char internal[128]; };

namespace {
struct Group {
  int world = 0, joined = 0, arrived = 0, generation = 0, alive = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<const void*> send;
  void barrier(std::unique_lock<std::mutex>& lk) {
    const int g = generation;
    if (++arrived == world) { arrived = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != g; });
  }
};
std::mutex g_mu;
std::map<std::string, Group*> g_groups;
int g_next_id = 1;
}  // namespace

struct ncclComm { Group* grp; int rank; std::string dir, id; int world = 0; long seq = 0; };
typedef ncclComm* ncclComm_t;

extern "C" {

ncclResult_t ncclGetVersion(int* v) { *v = 99999; return ncclSuccess; }   // recognisably not a real NCCL

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(id, 0, sizeof(*id));
  snprintf(id->internal, sizeof(id->internal), "fake-nccl-%d", g_next_id++);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* comm, int world, ncclUniqueId id, int rank) {
  if (!comm || world < 1 || rank < 0 || rank >= world) return ncclInvalidArgument;
  if (const char* dir = getenv("MGB_FAKE_NCCL_DIR")) {       // one process per rank: file rendezvous
    ncclComm* c = new ncclComm{nullptr, rank};
    c->dir = dir; c->id = id.internal; c->world = world;
    *comm = c;
    return ncclSuccess;
  }
  Group* grp;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    Group*& slot = g_groups[std::string(id.internal, sizeof(id.internal))];
    if (!slot) { slot = new Group(); slot->world = world; slot->alive = world; slot->send.assign(world, nullptr); }
    grp = slot;
  }
  if (grp->world != world) return ncclInvalidArgument;
  *comm = new ncclComm{grp, rank};
  std::unique_lock<std::mutex> lk(grp->mu);      // collective: returns when every rank has joined
  grp->joined++;
  grp->cv.notify_all();
  grp->cv.wait(lk, [&] { return grp->joined >= grp->world; });
  return ncclSuccess;
}

ncclResult_t ncclCommInitAll(ncclComm_t* comms, int n, const int*) {
  if (!comms || n < 1) return ncclInvalidArgument;
  Group* grp = new Group();
  grp->world = grp->joined = grp->alive = n;
  grp->send.assign(n, nullptr);
  for (int r = 0; r < n; r++) comms[r] = new ncclComm{grp, r};
  return ncclSuccess;
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, int /*dtype: bytes*/, ncclComm_t comm, void* /*stream*/) {
  if (!comm->grp) {                                          // file rendezvous between processes
    const std::string base = comm->dir + "/" + comm->id + "." + std::to_string(comm->seq++) + ".";
    const std::string mine = base + std::to_string(comm->rank), tmp = mine + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(send, 1, count, f) != count) return ncclInternalError;
    fclose(f);
    if (rename(tmp.c_str(), mine.c_str()) != 0) return ncclInternalError;
    for (int r = 0; r < comm->world; r++) {
      const std::string peer = base + std::to_string(r);
      FILE* g = nullptr;
      for (long spins = 0; !(g = fopen(peer.c_str(), "rb")); spins++) {
        if (spins > 3600L * 100) return ncclRemoteError;     // an hour: the peer is gone
        usleep(10000);
      }
      const size_t got = fread((char*)recv + (size_t)r * count, 1, count, g);
      fclose(g);
      if (got != count) return ncclInternalError;
    }
    return ncclSuccess;
  }
  Group* grp = comm->grp;
  std::unique_lock<std::mutex> lk(grp->mu);
  grp->send[comm->rank] = send;
  grp->barrier(lk);                              // every rank has published its buffer
  for (int r = 0; r < grp->world; r++) memcpy((char*)recv + (size_t)r * count, grp->send[r], count);
  grp->barrier(lk);                              // nobody overwrites its send buffer before all have copied
  return ncclSuccess;
}

ncclResult_t ncclCommGetAsyncError(ncclComm_t, ncclResult_t* async) {
  const char* ev = getenv("MGB_FAKE_NCCL_ASYNC_ERROR");
  *async = (ev && atoi(ev)) ? ncclRemoteError : ncclSuccess;
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  if (!comm) return ncclSuccess;
  if (!comm->grp) { delete comm; return ncclSuccess; }
  Group* grp = comm->grp;
  bool last;
  { std::lock_guard<std::mutex> lk(grp->mu); last = --grp->alive == 0; }
  if (last) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto it = g_groups.begin(); it != g_groups.end(); ++it)
      if (it->second == grp) { g_groups.erase(it); break; }
    delete grp;
  }
  delete comm;
  return ncclSuccess;
}

const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : (r == ncclRemoteError ? "remote process exited or there was a network error (fake)" : "fake NCCL error"); }

}  // extern "C"
