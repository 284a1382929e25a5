// Data-parallel gradient all-reduce over NCCL (NVLink 5 / NVSwitch).  The reference has no
// distributed code at all; this is the one exchange step of batch data
// parallelism.  libnccl is resolved at run time with dlopen so that the copy
// torch already loaded is the one used (no second NCCL in the process).
#include <dlfcn.h>
#include <mutex>

#include "common.cuh"

namespace gwn {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat = 7, ncclSum = 0, ncclAvg = 4 };

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_world = 1;
static std::mutex g_mu;

static int load_nccl() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  GWN_REQUIRE(g_nccl.h != nullptr, "cannot dlopen libnccl.so.2: %s", dlerror());
#define GWN_SYM(field, name)                                                     \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.h, name)); \
  GWN_REQUIRE(g_nccl.field != nullptr, "libnccl lacks %s", name)
  GWN_SYM(GetUniqueId, "ncclGetUniqueId");
  GWN_SYM(CommInitRank, "ncclCommInitRank");
  GWN_SYM(AllReduce, "ncclAllReduce");
  GWN_SYM(CommDestroy, "ncclCommDestroy");
  GWN_SYM(GetErrorString, "ncclGetErrorString");
#undef GWN_SYM
  return 0;
}
#define GWN_NCCL(expr)                                                                   \
  do {                                                                                   \
    ncclResult_t r__ = (expr);                                                           \
    if (r__ != 0) {                                                                      \
      set_error("%s -> %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
      return -3;                                                                         \
    }                                                                                    \
  } while (0)
}  // namespace gwn

using namespace gwn;

extern "C" int gwn_comm_unique_id(void* out128) {
  std::lock_guard<std::mutex> lk(g_mu);
  GWN_REQUIRE(out128 != nullptr, "comm_unique_id: NULL");
  if (int rc = load_nccl()) return rc;
  GWN_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(out128)));
  return 0;
}

extern "C" int gwn_comm_init(const void* id128, int rank, int world) {
  std::lock_guard<std::mutex> lk(g_mu);
  GWN_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, "comm_init: bad argument");
  GWN_REQUIRE(g_comm == nullptr, "comm_init: communicator already initialised");
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id = *reinterpret_cast<const ncclUniqueId*>(id128);
  GWN_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
  g_world = world;
  return 0;
}

extern "C" int gwn_comm_allreduce_avg(float* buf, long long count, void* stream) {
  GWN_REQUIRE(g_comm != nullptr, "comm_allreduce_avg: communicator not initialised");
  GWN_REQUIRE(buf && count >= 0, "comm_allreduce_avg: bad argument");
  if (count == 0) return 0;
  GWN_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat, ncclAvg, g_comm,
                            reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int gwn_comm_destroy(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) {
    GWN_NCCL(g_nccl.CommDestroy(g_comm));
    g_comm = nullptr;
  }
  return 0;
}
