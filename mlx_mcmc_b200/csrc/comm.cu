// Observation sharding: the per-gradient exchange.  Every rank holds all chains and a row shard of
// (X, y); after K6 the contiguous [Cp, Dp] gradient partial plus the [Cp] sum of squared residuals is summed over
// ranks with one ncclAllReduce on the compute stream, so every rank finishes with bit-identical log p / gradient
// and takes identical accept / U-turn decisions without any further synchronisation.
//
// NCCL is bound at run time (dlopen): the library loads and every other entry point works on a box without NCCL.
// The process usually already has libnccl.so.2 mapped (torch links it), in which case that copy is reused.
#include <dlfcn.h>
#include <string.h>

#include <string>

#include "glm.cuh"

namespace b2m {

struct Id128 {  // ncclUniqueId: 128 opaque bytes, passed by value to ncclCommInitRank
  char bytes[128];
};

namespace {

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  const char *env = getenv("B2M_NCCL_LIB");
  if (!h && env) h = dlopen(env, RTLD_NOW);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) {
    set_error(std::string("NCCL is not available: ") + dlerror());
    return 3;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.ReduceScatter = reinterpret_cast<decltype(a.ReduceScatter)>(dlsym(h, "ncclReduceScatter"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.AllGather || !a.ReduceScatter || !a.CommDestroy ||
      !a.GetErrorString) {
    set_error("NCCL library lacks an expected symbol");
    return 3;
  }
  g_nccl = a;
  return 0;
}

constexpr int kNcclFloat32 = 7, kNcclInt64 = 4, kNcclSum = 0;

#define B2M_CHECK_NCCL(expr)                                                             \
  do {                                                                                   \
    int _r = (expr);                                                                     \
    if (_r != 0) {                                                                       \
      set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r));                  \
      return 4;                                                                          \
    }                                                                                    \
  } while (0)

}  // namespace

struct Comm {
  void *nccl = nullptr;
  int nranks = 1, rank = 0;
};

int comm_unique_id(uint8_t *out128) {
  if (int rc = load_nccl()) return rc;
  Id128 id;
  memset(&id, 0, sizeof(id));
  B2M_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, id.bytes, 128);
  return 0;
}

int comm_init(const uint8_t *id128, int nranks, int rank, Comm **out) {
  if (int rc = load_nccl()) return rc;
  Id128 id;
  memcpy(id.bytes, id128, 128);
  Comm *c = new Comm();
  c->nranks = nranks;
  c->rank = rank;
  int r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
  if (r != 0) {
    set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    delete c;
    return 4;
  }
  *out = c;
  return 0;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
  delete c;
}

int comm_nranks(const Comm *c) { return c ? c->nranks : 1; }
int comm_rank(const Comm *c) { return c ? c->rank : 0; }

// in place: rank r's `count` elements live at buf + r * count; afterwards every rank holds all nranks * count
int comm_allgather_inplace(Comm *c, void *buf, int64_t count, int elt_bytes, cudaStream_t st) {
  B2M_REQUIRE(c && c->nccl, "allgather: communicator is NULL");
  const int dt = elt_bytes == 8 ? kNcclInt64 : kNcclFloat32;   // 8-byte elements travel as int64, 4-byte as float32
  char *base = static_cast<char *>(buf);
  B2M_CHECK_NCCL(g_nccl.AllGather(base + (size_t)c->rank * count * elt_bytes, buf, (size_t)count, dt, c->nccl, st));
  return 0;
}

// in place: sums buf[nranks * count] over ranks; rank r ends with its block, at buf + r * count
int comm_reducescatter_f32_inplace(Comm *c, float *buf, int64_t count, cudaStream_t st) {
  B2M_REQUIRE(c && c->nccl, "reducescatter: communicator is NULL");
  B2M_CHECK_NCCL(g_nccl.ReduceScatter(buf, buf + (size_t)c->rank * count, (size_t)count, kNcclFloat32, kNcclSum, c->nccl, st));
  return 0;
}

int comm_allreduce_f32(Comm *c, float *buf, int64_t n, cudaStream_t st) {
  B2M_REQUIRE(c && c->nccl, "allreduce: communicator is NULL");
  B2M_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, c->nccl, st));
  return 0;
}

int comm_allreduce_i64(Comm *c, int64_t *buf, int64_t n, cudaStream_t st) {
  B2M_REQUIRE(c && c->nccl, "allreduce: communicator is NULL");
  B2M_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclInt64, kNcclSum, c->nccl, st));
  return 0;
}

}  // namespace b2m
