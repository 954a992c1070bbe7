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

constexpr int kNcclFloat32 = 7, kNcclInt64 = 4, kNcclSum = 0, kNcclMax = 2;

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

int comm_allreduce_f32_max(Comm *c, float *buf, int64_t n, cudaStream_t st) {
  B2M_REQUIRE(c && c->nccl, "allreduce: communicator is NULL");
  B2M_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclFloat32, kNcclMax, c->nccl, st));
  return 0;
}

int comm_allreduce_i64(Comm *c, int64_t *buf, int64_t n, cudaStream_t st) {
  B2M_REQUIRE(c && c->nccl, "allreduce: communicator is NULL");
  B2M_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclInt64, kNcclSum, c->nccl, st));
  return 0;
}

// ---------------------------------------------------------------- peer window (B2M_SLICE_PEER)
// One cudaMalloc'd window per rank, shared with the other ranks of the box through CUDA IPC handles (the handles
// travel over torch.distributed; include/b200mcmc.h).  All traffic through it is plain NVLink stores from the state
// kernel (packed rows) and from K6's epilogue (gradient tiles), ordered by release / acquire flags at system scope.
size_t peer_layout(PeerWindow &w, int64_t C, int nranks, int Dp) {
  auto up = [](size_t x) { return (x + 1023) & ~size_t(1023); };
  w.C = C; w.nranks = nranks; w.own = C / nranks; w.Dp = Dp;
  size_t o = 0;
  w.off_g = o;     o = up(o + sizeof(float) * (size_t)C * Dp);
  w.off_ss = o;    o = up(o + sizeof(float) * (size_t)C);
  w.off_bh = o;    o = up(o + sizeof(__half) * (size_t)C * Dp);
  w.off_bl = o;    o = up(o + sizeof(__half) * (size_t)C * Dp);
  w.off_meta = o;  o = up(o + sizeof(float4) * (size_t)C);
  w.off_bflag = o; o = up(o + sizeof(uint64_t) * kMaxPeers);
  w.off_gflag = o; o = up(o + sizeof(uint64_t) * kMaxPeers);
  w.off_done = o;  o = up(o + sizeof(int64_t) * kMaxPeers);
  w.off_err = o;   o = up(o + sizeof(int) * 4);
  w.bytes = o;
  return o;
}

int peer_alloc(int64_t bytes, void **ptr, uint8_t *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  void *p = nullptr;
  B2M_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  B2M_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    return 2;
  }
  memcpy(handle64, &h, 64);
  B2M_CHECK_CUDA(cudaDeviceSynchronize());
  *ptr = p;
  return 0;
}

int peer_open(const uint8_t *handle64, void **ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B2M_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int peer_close(void *ptr) {
  B2M_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int peer_free(void *ptr) {
  B2M_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

int glm_peer_attach(GlmModel &g, void *const *windows, int n_ranks, int rank, int64_t n_chains, int64_t bytes) {
  B2M_REQUIRE(g.comm && comm_nranks(g.comm) == n_ranks && comm_rank(g.comm) == rank,
              "peer window: attach the NCCL communicator (b2m_model_set_comm) first; ranks must agree");
  PeerWindow w;
  const size_t need = peer_layout(w, n_chains, n_ranks, g.Dp);
  B2M_REQUIRE((size_t)bytes >= need, "peer window: the window is smaller than b2m_model_peer_bytes()");
  w.rank = rank;
  for (int s = 0; s < n_ranks; ++s) w.base[s] = static_cast<char *>(windows[s]);
  w.seq = 0;
  g.pw = w;
  if (!g.blk_counter) {
    B2M_CHECK_CUDA(cudaMalloc(reinterpret_cast<void **>(&g.blk_counter), sizeof(unsigned) * 8));
    B2M_CHECK_CUDA(cudaMemset(g.blk_counter, 0, sizeof(unsigned) * 8));
  }
  return 0;
}

}  // namespace b2m
