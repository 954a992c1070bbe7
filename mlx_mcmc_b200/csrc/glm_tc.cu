// K5 / K6 on the 5th-generation tensor cores: TMA-staged tcgen05 GEMMs in 3xTF32.
//
//   D[m, n] = sum_k (Ah + Al)[m, k] * (Bh + Bl)[n, k]   ~=   Ah.Bh + Ah.Bl + Al.Bh      (fp32 accumulate in TMEM)
//
// Every operand is K-major fp32 holding tf32-representable values (round-to-nearest hi/lo split, glm.cuh), so
// the dropped Al.Bl term is ~2^-22 relative: the two contractions hold the 1e-5 log-prob / gradient tolerance.
//
// Two launch shapes share the kernel body (template NCTA):
//   NCTA = 1   one CTA = one 128 x BLOCK_N tile                                   (small / compacted batches)
//   NCTA = 2   a CTA pair (cluster of 2, tcgen05 cta_group::2) = one 256 x BLOCK_N tile: each CTA stages its own
//              128 rows of A and HALF of the B tile, the leader CTA issues the MMAs for both, each CTA's TMEM
//              receives its own 128 rows.  Per k-block a CTA pulls 64 KB instead of 96 KB through L2 -- the v1
//              kernel was pinned at the ~6300 B/clk L2->SM return path (profiles/r01_tc_gemm_c4_v1_ncu_summary.md).
//
// Kernel anatomy (one 128 x BLOCK_N output tile per CTA, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of the four operand tiles of a k-block
//               (BLOCK_K = 32 floats = one 128-byte swizzle row) into a ring of shared-memory stages
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (3 MMAs x 4 k-steps per stage),
//               tcgen05.commit releases the stage / hands a finished K-chunk to the promotion warps
//   warps 2..9  promotion + epilogue.  The tensor core adds into its fp32 accumulator with truncation, so a
//               long K loop drifts (~T * 2^-24 relative after T MMAs: measured 4e-6 at K = 1000, would be
//               ~2e-3 at K = 100,000).  Therefore TMEM holds only a CHUNK of 4 k-blocks (K = 128) at a time,
//               double buffered; these warps tcgen05.ld each finished chunk and add it round-to-nearest into
//               fp32 master accumulators in registers (lane = output row, 8 warps = 4 lane quarters x 2 column
//               halves).  At the end:
//                 RESID: z = y - c - acc, per-row sum z^2, R = w z / sigma^2 split into tf32 hi/lo   (K5)
//                 PLAIN: store the split-K partial of G                                              (K6)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "glm.cuh"

namespace b2m {

namespace {

constexpr int BLOCK_M = 128;
constexpr int ROW_BYTES = 128;                        // one SWIZZLE_128B row of K: 32 tf32 values or 64 fp16 values
constexpr int MMA_K_BYTES = 32;                       // one MMA k-step: 8 tf32 or 16 fp16 values
constexpr int A_TILE_BYTES = BLOCK_M * ROW_BYTES;     // 16 KB
// elements of K per k-block (one swizzle row) for the two operand encodings
template <bool F16> struct Enc { static constexpr int BLOCK_K = F16 ? 64 : 32; };
constexpr int BLOCK_K = 32;                           // tf32 k-block (host-side helpers of the tf32 path)
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;   // TMA warp, MMA warp, 8 promotion/epilogue warps
constexpr int DEFAULT_CHUNK_KB = 4;                     // k-blocks accumulated inside the tensor core per chunk

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- cluster / CTA-pair helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): the arrival only says "this
  // warp has drained its TMEM reads / issued its loads"; a cluster-scope release compiled to MEMBAR.ALL.GPU + ERRBAR
  // and was 7 % of the K5 stall samples
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// L2 eviction policies for the TMA loads (opt-in, B2M_TC_L2_HINTS=1).  K6 streams the 1.6 GB residual operand past the 0.4 GB
// transposed design matrix that sixteen row tiles re-read, and reads 3.1-4.0 GB from DRAM for 2.05 GB of operands.  Hints
// (K6: A -> evict_first, B -> evict_last; K5: its small, re-read A operand -> evict_last) were measured and did not help
// (K5 + K6 3.63 ms with, 3.42-3.51 ms without); what did was the tile order (column tile fastest, see the tile loop).
// code: 1 = evict_first, 2 = evict_last (0 = no hint: the plain load is issued instead)
__device__ __forceinline__ uint64_t l2_policy(int code) {
  uint64_t p;
  if (code == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// ---- dependency counters of the concurrent K5 || K6 launch (global memory, gpu scope)
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until *p >= need.  A wait that lasts ~1.5 s raises err[0] (device memory: every later wait returns at once) and the
// pinned host flag err_host: a broken dependency makes the call fail, it cannot hang the GPU.
__device__ __forceinline__ void spin_ge(const int *p, int need, volatile int *err, volatile int *err_host) {
  if (ld_acquire_gpu(p) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < need) {
    __nanosleep(64);
    if (*err) return;
    if (clock64() - t0 > (1ll << 31)) { *err = 1; *err_host = 1; __threadfence_system(); return; }
  }
}
// TMA load issued by one CTA of a pair; completion bytes are signalled on the LEADER's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void *dst, const CUtensorMap *map, uint32_t leader_bar, int x, int y,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(x), "r"(y), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives on the mbarrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool F16, int NCTA>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (F16) {
    if (NCTA == 2) umma_f16_pair(tmem_d, adesc, bdesc, idesc, accumulate); else umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    if (NCTA == 2) umma_tf32_pair(tmem_d, adesc, bdesc, idesc, accumulate); else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 | LBO (unused for swizzled K-major, = 1) << 16 | SBO = 1024 B (8 rows x 128 B) >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
// a_format / b_format: 0 = F16, 2 = TF32
__host__ __device__ constexpr uint32_t make_idesc(int n, int m = BLOCK_M, bool f16 = false) {
  return (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct EpiParams {
  // RESID
  const float *y;
  const float *inv_var;
  void *Rh, *Rl;          // [Cp, Np] residual operand of K6: float (tf32 values) or __half
  float *ss_part;
  const float *a_unscale; // F16: [Cp] 1 / (row scale of the A operand), applied to the accumulators
  const float *r_scale;   // F16: [Cp] row scale of R before the fp16 split
  int64_t Cp;
  int Np, N_valid;
  float loc_const, weight;
  // PLAIN
  float *Gpart;  // [splits, Cp, Dp]
  int Dp;
  // both: the launch is a no-op once *skip_flag >= skip_target (every chain of the call has finished; the host learns
  // it a few launches late because it reads the counter without draining the stream)
  const int *skip_flag;
  int skip_target;
  int nt_fast;    // K6 tile order: column tile fastest (see the tile loop); 0 = pair row fastest (B2M_TC_K6_ORDER=0)
  int l2_hints;   // CTA-pair kernels: L2 eviction-policy codes of the TMA loads, A | B << 2 (see l2_policy)
  // FUSED (concurrent K5 || K6, tc_gemm_fused_kernel): the residual operand is a ring of `fz_ring` slabs of `fz_slab`
  // 256-observation tiles, [Cp, fz_pitch] halves; fz_ready[pair row][slab] counts the epilogue warps of K5 tiles that
  // have stored their part of the slab, fz_turn[pair row][K6 column tile] the epilogue warps of K6 tiles that have added
  // their slab partial into G (slab order: the sum is deterministic)
  int fz_slab, fz_ring, fz_pitch, fz_nslabs, fz_tn5, fz_tn6;
  int *fz_ready, *fz_turn;
  volatile int *fz_err, *fz_err_host;
  // PUSH (K6 of a peer-sliced observation shard): the finished tile goes, unscaled, straight into the window of the
  // rank that owns its chains -- slot `push_rank` of that rank's [nranks][own][Dp] gradient block -- over NVLink
  float *push_dst[kMaxPeers];
  int push_own, push_rank;
  const float *r_unscale;       // [Cp]
  const float *inv_col_scale;   // [Dp]
};

template <int BLOCK_N, int NCTA>
struct Cfg {
  static constexpr int B_ROWS = BLOCK_N / NCTA;         // rows of the B tile staged by one CTA
  static constexpr int B_TILE_BYTES = B_ROWS * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 6 ? 6 : (192 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;        // two chunk accumulators (power of two >= 32)
  static constexpr int COLS_PER_WARP = BLOCK_N / 2;    // each promotion warp owns 32 rows x half of the columns
};

// MODE: 0 = PLAIN (split-K partial of G), 1 = RESID (K5 residual epilogue), 2 = PUSH (K6 tile -> owner's window)
// FUSED: the body runs as one role of tc_gemm_fused_kernel (CTA pairs `group` of `n_groups` of that role); the residual
// operand is the slab ring described at EpiParams and the two roles synchronise through its counters.
template <int BLOCK_N, int MODE, int NCTA, bool F16, bool FUSED>
__device__ __forceinline__ void
tc_gemm_body(const CUtensorMap &tmAh, const CUtensorMap &tmAl, const CUtensorMap &tmBh, const CUtensorMap &tmBl,
             const int k_blocks_total, const int k_blocks_per_split, const int CHUNK_KB, const int mma_mask, const int Tm,
             const int Tn, const int n_tiles, const EpiParams &E, const int group, const int n_groups) {
  using C = Cfg<BLOCK_N, NCTA>;
  constexpr bool RESID = MODE == 1, PUSH = MODE == 2;
  static_assert(!FUSED || (NCTA == 2 && F16 && !PUSH), "the concurrent launch exists for fp16 CTA-pair tiles");
  constexpr int BLOCK_K = Enc<F16>::BLOCK_K;   // K elements per k-block (shadows the tf32 constant)
  constexpr int SCRATCH_BYTES = (RESID || PUSH || FUSED) ? NUM_EPI_WARPS * 32 * 33 * 4 : 0;   // per-warp transpose scratch of the epilogue
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float *scratch_base = reinterpret_cast<float *>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::STAGES * C::STAGE_BYTES + SCRATCH_BYTES);
  uint64_t *empty = full + C::STAGES;
  uint64_t *tmem_full = empty + C::STAGES;   // [2]
  uint64_t *tmem_empty = tmem_full + 2;      // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = NCTA == 2 ? cluster_ctarank() : 0u;   // rank inside the CTA pair; 0 = leader (issues the MMAs)
  // Persistent tile loop: CTA group g (a CTA, or a pair) takes tiles g, g + G, g + 2G, ...  K5: tile t = (pair-row
  // t % Tm, observation tile t / Tm) -- the sixteen pairs on one observation tile share the B tile (design matrix rows),
  // the A operand (packed positions, 16 MB) lives in L2.  K6: tile t = (column tile t % Tn, pair-row (t / Tn) % Tm, K split
  // t / (Tm Tn)) -- the Tn neighbours share the A tile, which is the big stream there (the residual operand: 1.6 GB per
  // evaluation at C4 against 0.4 GB of transposed design matrix); with the pair row fastest the Tn readers of one residual
  // tile were 16 tiles apart, often in different rounds of the 74 pairs, and the tile came from DRAM more than once.  Barriers and TMEM
  // are set up once; the producer and the MMA issuer run ahead into the next tile while the promotion warps are
  // still in the epilogue of the previous one (both TMEM chunk buffers are free by then).
#define B2M_DECODE_TILE(t)                                                         \
  const bool ntf_ = !RESID && E.nt_fast;                                           \
  const int mp_ = ntf_ ? ((t) / Tn) % Tm : (t) % Tm;                               \
  const int nt = ntf_ ? (t) % Tn : ((t) / Tm) % Tn;                                \
  const int zs = (t) / (Tm * Tn);                                                  \
  const int m0 = (mp_ * NCTA + (int)cta) * BLOCK_M, n0 = nt * BLOCK_N;             \
  const int kb0 = zs * k_blocks_per_split;                                         \
  const int kb1 = min(kb0 + k_blocks_per_split, k_blocks_total);                   \
  const int nkb = kb1 - kb0;                                                       \
  const int n_chunks = (nkb + CHUNK_KB - 1) / CHUNK_KB;                            \
  (void)m0; (void)n0; (void)nt; (void)zs; (void)nkb; (void)n_chunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    // pair: the leader's `full` collects one arrival per producer (2) and the bytes of both CTAs' loads; the leader's
    // `tmem_empty` collects the promotion threads of both CTAs; `empty` / `tmem_full` are signalled in both CTAs by
    // the multicast tcgen05.commit
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], NCTA); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], NCTA * NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();   // barriers initialised and TMEM allocated in both CTAs
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // eviction-policy codes of the two operands (0 = plain load), set by the launcher
      const int code_a = E.l2_hints & 3, code_b = (E.l2_hints >> 2) & 3;
      const uint64_t pol_a = l2_policy(code_a), pol_b = l2_policy(code_b);
      for (int t = group; t < n_tiles; t += n_groups) {
      B2M_DECODE_TILE(t)
      if (FUSED && !RESID) {
        // K6 role: the slab's rows of this pair are complete once every epilogue warp (8 per CTA) of every K5 tile that
        // covers them has arrived.  The residuals were written through the generic proxy and are read by TMA: acquire,
        // then order the async proxy behind it.
        const int tiles = min(E.fz_slab, E.fz_tn5 - zs * E.fz_slab);
        spin_ge(E.fz_ready + mp_ * E.fz_nslabs + zs, tiles * NCTA * NUM_EPI_WARPS, E.fz_err, E.fz_err_host);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      // column of the A operand: K6 of the concurrent launch reads slab zs from its slot of the ring
      const int ka0 = (FUSED && !RESID) ? ((zs % E.fz_ring) * k_blocks_per_split - kb0) * BLOCK_K : 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        unsigned char *st = smem + stage * C::STAGE_BYTES;
        if (NCTA == 2) {
          const uint32_t lbar = mapa(smem_u32(&full[stage]), 0);   // the leader's barrier
          if (cta == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
          else mbar_arrive_cluster(lbar);
          const int nb = n0 + (int)cta * C::B_ROWS;               // this CTA's half of the B tile
          if (code_a) {
            tma_load_2d_pair_hint(st, &tmAh, lbar, kb * BLOCK_K + ka0, m0, pol_a);
            tma_load_2d_pair_hint(st + A_TILE_BYTES, &tmAl, lbar, kb * BLOCK_K + ka0, m0, pol_a);
          } else {
            tma_load_2d_pair(st, &tmAh, lbar, kb * BLOCK_K + ka0, m0);
            tma_load_2d_pair(st + A_TILE_BYTES, &tmAl, lbar, kb * BLOCK_K + ka0, m0);
          }
          if (code_b) {
            tma_load_2d_pair_hint(st + 2 * A_TILE_BYTES, &tmBh, lbar, kb * BLOCK_K, nb, pol_b);
            tma_load_2d_pair_hint(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, &tmBl, lbar, kb * BLOCK_K, nb, pol_b);
          } else {
            tma_load_2d_pair(st + 2 * A_TILE_BYTES, &tmBh, lbar, kb * BLOCK_K, nb);
            tma_load_2d_pair(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, &tmBl, lbar, kb * BLOCK_K, nb);
          }
        } else {
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          tma_load_2d(st, &tmAh, &full[stage], kb * BLOCK_K, m0);
          tma_load_2d(st + A_TILE_BYTES, &tmAl, &full[stage], kb * BLOCK_K, m0);
          tma_load_2d(st + 2 * A_TILE_BYTES, &tmBh, &full[stage], kb * BLOCK_K, n0);
          tma_load_2d(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, &tmBl, &full[stage], kb * BLOCK_K, n0);
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0 && cta == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N, BLOCK_M * NCTA, F16);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t gch = 0;   // chunk counter across tiles: buffer = gch & 1, barrier parity = (gch >> 1) & 1
      for (int t = group; t < n_tiles; t += n_groups) {
      B2M_DECODE_TILE(t)
      int kb = 0;
      for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
        const int buf = gch & 1;
        mbar_wait(&tmem_empty[buf], ((gch >> 1) & 1) ^ 1);   // the promotion warps have drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
        const int kb_end = min(kb + CHUNK_KB, nkb);
        for (int kc = 0; kb < kb_end; ++kb, ++kc) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t dAh = make_desc(st), dAl = make_desc(st + A_TILE_BYTES);
          const uint64_t dBh = make_desc(st + 2 * A_TILE_BYTES), dBl = make_desc(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k) {
            const uint64_t adv = (uint64_t)((k * MMA_K_BYTES) >> 4);  // start-address advance inside the swizzle row
            umma<F16, NCTA>(tmem_d, dAh + adv, dBh + adv, idesc, (kc | k) != 0 ? 1u : 0u);
            if (mma_mask & 2) umma<F16, NCTA>(tmem_d, dAh + adv, dBl + adv, idesc, 1u);
            if (mma_mask & 4) umma<F16, NCTA>(tmem_d, dAl + adv, dBh + adv, idesc, 1u);
          }
          // frees the stage (in both CTAs of a pair) once the MMAs above have read it
          if (NCTA == 2) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        // chunk accumulator complete (in both CTAs' TMEM)
        if (NCTA == 2) umma_commit_pair(&tmem_full[buf]); else umma_commit(&tmem_full[buf]);
      }
      }
    }
  } else {
    // ===== promotion + epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int CW = C::COLS_PER_WARP;
    uint32_t gch = 0;
    for (int t = group; t < n_tiles; t += n_groups) {
    B2M_DECODE_TILE(t)
    const int m = m0 + q * 32 + lane;
    B2M_ASSERT(m >= 0 && (int64_t)m < E.Cp && nkb > 0);
    float acc[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) acc[j] = 0.f;
    bool turn_ok = true;
    if (FUSED && !RESID && zs > 0) {
      // K6 role: has the previous slab's partial of this output tile been added?  Sampled now, off the critical path
      // (the tile that passes the turn on started a whole round of tiles earlier); the epilogue only waits if not.
      int ok = 0;
      if (lane == 0) ok = ld_acquire_gpu(E.fz_turn + mp_ * E.fz_tn6 + nt) >= zs * NCTA * NUM_EPI_WARPS;
      turn_ok = __shfl_sync(0xffffffffu, ok, 0) != 0;
    }
    for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
      const int buf = gch & 1;
      mbar_wait(&tmem_full[buf], (gch >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BLOCK_N + half * CW;
#pragma unroll
      for (int cb = 0; cb < CW / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(trow + cb * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cb * 32 + j] += __uint_as_float(v[j]);   // round-to-nearest promotion
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {   // one arrival per warp: a cluster-scope release per thread was 13 % of the v2 stall samples
        if (NCTA == 2) mbar_arrive_cluster(mapa(smem_u32(&tmem_empty[buf]), 0));   // the leader's MMA warp waits on it
        else mbar_arrive(&tmem_empty[buf]);
      }
    }
    const int nb = n0 + half * CW;
    if (RESID) {
      // Residual epilogue.  The thread owns row m (a chain) and CW consecutive columns (observations); written
      // straight from registers each store instruction would touch 32 different rows (16 B per row: uncoalesced,
      // partial-sector writes -- the first version spent more time here than in the MMAs).  Each warp therefore
      // transposes 32x32 blocks through a private shared-memory scratch (33-float pitch: conflict-free both ways)
      // and stores whole 128-byte row segments.
      float ss = 0.f;
      const float ivw = E.inv_var[m] * E.weight;
      // F16: the accumulators carry the row scale of the A operand (delta); R gets its own row scale before the split
      const float a_un = F16 ? E.a_unscale[m] : 1.0f;
      const float rs_lane = F16 ? E.r_scale[m] : 1.0f;   // lane r holds the R scale of row r of this warp's 32 rows
      float *scratch = scratch_base + (warp - 2) * (32 * 33);
      const int64_t row0 = (int64_t)(m0 + q * 32);
      // where the residuals go: [Cp, Np], or (concurrent launch) the slab's slot of the [Cp, fz_pitch] ring
      const int r_pitch = FUSED ? E.fz_pitch : E.Np;
      const int slab = FUSED ? nt / E.fz_slab : 0;
      const int rcol = FUSED ? ((slab % E.fz_ring) * E.fz_slab + nt % E.fz_slab) * BLOCK_N + half * CW : nb;
      if (FUSED && slab >= E.fz_ring) {
        // the slot still holds slab - ring until every K6 tile of this pair row has consumed it (its turn counters are
        // raised after the tile's last MMA has read the operand)
        if (lane < E.fz_tn6) spin_ge(E.fz_turn + mp_ * E.fz_tn6 + lane, (slab - E.fz_ring + 1) * NCTA * NUM_EPI_WARPS, E.fz_err, E.fz_err_host);
        __syncwarp();
      }
      // the observations of this warp's columns: one coalesced load per 32-column block (lane = column), handed to
      // the row-owning threads by shuffle (a per-element __ldg serialised 128 L2 round trips per thread)
      float yv[CW / 32];
#pragma unroll
      for (int b = 0; b < CW / 32; ++b) {
        const int n = nb + b * 32 + lane;
        yv[b] = (n < E.N_valid) ? __ldg(E.y + n) - E.loc_const : 0.f;
      }
      // (columns n >= N_valid need no test below: the padded rows of X are zero, so acc = 0 there, and yv = 0)
#pragma unroll
      for (int b = 0; b < CW / 32; ++b) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float yj = __shfl_sync(0xffffffffu, yv[b], j);
          const float z = yj - (F16 ? acc[b * 32 + j] * a_un : acc[b * 32 + j]);
          ss = fmaf(z, z, ss);
          // F16: the row scale of R is applied here, by the thread that owns the row (same two multiplications, in the
          // same order, as scaling after the transposition -- without a shuffle and two multiplies per stored pair)
          scratch[lane * 33 + j] = F16 ? (z * ivw) * rs_lane : z * ivw;
        }
        __syncwarp();
        if (F16) {
          // two rows per instruction: lanes 0-15 take row r, lanes 16-31 row r + 1, each lane two adjacent columns
          // (bank = 33 row + 2 col': even banks for one half-warp, odd for the other -- conflict free) packed into
          // one half2 store per array: half the conversions and half the store instructions of a per-element loop
          const int hrow = lane >> 4, cl = (lane & 15) * 2;
          const int64_t off = (row0 + hrow) * r_pitch + rcol + b * 32 + cl;
          __half2 *rh = reinterpret_cast<__half2 *>(static_cast<__half *>(E.Rh) + off);
          __half2 *rl = reinterpret_cast<__half2 *>(static_cast<__half *>(E.Rl) + off);
#pragma unroll 8
          for (int r = 0; r < 32; r += 2) {
            const float v0 = scratch[(r + hrow) * 33 + cl], v1 = scratch[(r + hrow) * 33 + cl + 1];
            const __half2 hi = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(hi);
            rh[(int64_t)r * (r_pitch / 2)] = hi;
            rl[(int64_t)r * (r_pitch / 2)] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          }
        } else {
          const int64_t off = row0 * E.Np + nb + b * 32 + lane;
          float *rh = static_cast<float *>(E.Rh) + off, *rl = static_cast<float *>(E.Rl) + off;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            float hi, lo;
            split_tf32(scratch[r * 33 + lane], hi, lo);
            rh[(int64_t)r * E.Np] = hi;
            rl[(int64_t)r * E.Np] = lo;
          }
        }
        __syncwarp();
      }
      E.ss_part[((int64_t)nt * 2 + half) * E.Cp + m] = ss;
      if (FUSED) {
        // publish this warp's part of the slab to the K6 role: every lane orders its own stores (gpu scope, and ahead of
        // the async proxy that will read them), then one arrival per warp
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          __threadfence();   // cumulative: the warp's stores, observed through the barrier, precede the arrival
          atomicAdd(E.fz_ready + mp_ * E.fz_nslabs + slab, 1);
        }
      }
    } else if (PUSH) {
      // The tile's 128 rows belong to one owner rank (own % 128 == 0).  Each warp transposes its 32 x 32 blocks through
      // the shared-memory scratch so that every store instruction writes one 128-byte row segment into the owner's
      // window: plain NVLink stores, issued while the MMA warp is already accumulating the next tile.  The operand
      // scales are undone here (the ranks' scales differ); the owner adds the per-source slots in rank order.
      const int row0 = m0 + q * 32;
      const int owner = row0 / E.push_own;
      B2M_ASSERT(owner >= 0 && owner < kMaxPeers && E.push_dst[owner] != nullptr && (row0 + 31) / E.push_own == owner);
      float *dst = E.push_dst[owner] + ((int64_t)E.push_rank * E.push_own + (row0 - owner * E.push_own)) * E.Dp + nb;
      const float ru = E.r_unscale ? E.r_unscale[m] : 1.0f;   // lane r: scale of row r of this warp's 32 rows
      float *scratch = scratch_base + (warp - 2) * (32 * 33);
#pragma unroll
      for (int b = 0; b < CW / 32; ++b) {
#pragma unroll
        for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = acc[b * 32 + j] * ru;
        __syncwarp();
        const float ics = E.inv_col_scale ? E.inv_col_scale[nb + b * 32 + lane] : 1.0f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) dst[(int64_t)r * E.Dp + b * 32 + lane] = scratch[r * 33 + lane] * ics;
        __syncwarp();
      }
    } else if (FUSED) {
      // K6 role: the slab partial is added into the one [Cp, Dp] gradient buffer in slab order (turn counter of the output
      // tile: 16 epilogue-warp arrivals per slab), so the sum is the same on every run.  The operand of this tile has
      // been read completely by now (its last chunk was committed), which is what the K5 role waits for before it
      // overwrites the ring slot.  Row segments of 128 bytes through the transposition scratch.  Every address receives
      // its additions one slab after the other (the turn is passed on only after a gpu-scope fence), so the rounding is
      // that of a sequential sum.
      if (zs > 0 && !turn_ok) {   // (the counter was already sampled when the tile began: normally nothing to wait for)
        if (lane == 0) spin_ge(E.fz_turn + mp_ * E.fz_tn6 + nt, zs * NCTA * NUM_EPI_WARPS, E.fz_err, E.fz_err_host);
        __syncwarp();
      }
      float *g = E.Gpart + (int64_t)(m0 + q * 32) * E.Dp + nb;
      float *scratch = scratch_base + (warp - 2) * (32 * 33);
      // after the transposition a lane takes four adjacent columns of rows rr, rr + 4, ...: one 16-byte reduction (or store)
      // per lane, four whole 128-byte row segments per warp instruction; scratch reads are conflict free (bank = rr + cc + k)
      const int rr = lane >> 3, cc = (lane & 7) * 4;
#pragma unroll
      for (int b = 0; b < CW / 32; ++b) {
#pragma unroll
        for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = acc[b * 32 + j];
        __syncwarp();
        float *gp = g + b * 32 + cc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rr + 4 * i;
          const float v0 = scratch[r * 33 + cc], v1 = scratch[r * 33 + cc + 1], v2 = scratch[r * 33 + cc + 2], v3 = scratch[r * 33 + cc + 3];
          float *dst = gp + (int64_t)r * E.Dp;
          if (zs > 0) {
            // reductions performed at L2, no value returned: a load + add + store per row exposed one L2 / DRAM round trip
            // per row to the promotion warps (measured: the K6 role ran 3.5x slower than its MMAs)
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
          } else {
            *reinterpret_cast<float4 *>(dst) = make_float4(v0, v1, v2, v3);
          }
        }
        __syncwarp();
      }
      // one arrival per warp; lane 0's fence is cumulative over the warp's reductions (ordered before it by the barrier)
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(E.fz_turn + mp_ * E.fz_tn6 + nt, 1);
      }
    } else {
      float *g = E.Gpart + ((int64_t)zs * E.Cp + m) * E.Dp + nb;
#pragma unroll
      for (int j = 0; j < CW; j += 4) *reinterpret_cast<float4 *>(g + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    }
    // the writers themselves order their NVLink stores before anything that follows the kernel (the "partials stored"
    // flag is raised by the next kernel on this stream): a system-scope fence per pushing thread, once per launch
    if (PUSH) __threadfence_system();
  }
#undef B2M_DECODE_TILE

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();   // no CTA of a pair may leave while its peer still signals it
  if (warp == 1) {
    if (NCTA == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

template <int BLOCK_N, int MODE, int NCTA, bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, int k_blocks_total,
               int k_blocks_per_split, int CHUNK_KB, int mma_mask, int Tm, int Tn, int n_tiles, const __grid_constant__ EpiParams E) {
  if (E.skip_flag && *reinterpret_cast<const volatile int *>(E.skip_flag) >= E.skip_target) return;   // uniform over the grid
  tc_gemm_body<BLOCK_N, MODE, NCTA, F16, false>(tmAh, tmAl, tmBh, tmBl, k_blocks_total, k_blocks_per_split, CHUNK_KB, mma_mask, Tm,
                                                Tn, n_tiles, E, (int)blockIdx.x / NCTA, (int)gridDim.x / NCTA);
}

// K5 || K6 of one evaluation in ONE launch (fp16 encoding, CTA pairs, 256-column tiles).  The first `groups5` CTA pairs
// run K5 over the observation tiles in order, the rest run K6 one slab of observations behind them; the residual operand
// lives in a ring of a few slabs that stays in L2 (it never reaches HBM: two launches write and re-read 3.2 GB of it per
// evaluation at C4), the transposed design matrix is read once per slab while the sixteen pair rows are on it, and the
// slab partials of the gradient are added into one buffer in slab order.  Both roles are the bodies of the separate
// kernels; only the dependency counters (EpiParams::fz_*) are new.  All CTAs of the launch are resident at once (one per
// SM), so a role waiting for the other cannot starve it.
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_fused_kernel(const __grid_constant__ CUtensorMap a5h, const __grid_constant__ CUtensorMap a5l,
                     const __grid_constant__ CUtensorMap b5h, const __grid_constant__ CUtensorMap b5l,
                     const __grid_constant__ CUtensorMap a6h, const __grid_constant__ CUtensorMap a6l,
                     const __grid_constant__ CUtensorMap b6h, const __grid_constant__ CUtensorMap b6l, int kb5, int chunk5,
                     int Tm, int kb6_total, int kb6_per, int chunk6, int groups5, const __grid_constant__ EpiParams E5,
                     const __grid_constant__ EpiParams E6) {
  if (E5.skip_flag && *reinterpret_cast<const volatile int *>(E5.skip_flag) >= E5.skip_target) return;   // uniform over the grid
  const int group = (int)blockIdx.x / 2, n_groups = (int)gridDim.x / 2;
  if (group < groups5)
    tc_gemm_body<256, 1, 2, true, true>(a5h, a5l, b5h, b5l, kb5, kb5, chunk5, 7, Tm, E5.fz_tn5, Tm * E5.fz_tn5, E5, group, groups5);
  else
    tc_gemm_body<256, 0, 2, true, true>(a6h, a6l, b6h, b6l, kb6_total, kb6_per, chunk6, 7, Tm, E6.fz_tn6,
                                        Tm * E6.fz_tn6 * E6.fz_nslabs, E6, group - groups5, n_groups - groups5);
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, cols] row-major fp32, box = 32 columns (128 bytes) x box_rows rows, 128-byte swizzle.
// Encoded maps are cached by (base, rows, cols, box): a NUTS leaf would otherwise pay eight driver calls per
// value+gradient, which is visible at the small configurations where a leaf is ~100 us of device time.
int make_map_uncached(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int box_rows, bool f16);

struct MapKey {
  const void *base;
  int64_t rows, cols;
  int box;
  bool operator==(const MapKey &o) const { return base == o.base && rows == o.rows && cols == o.cols && box == o.box; }
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
};
static thread_local std::vector<MapSlot> tl_maps;

// `cols` in elements of the operand encoding (float for tf32, __half for f16: box is negative-coded as -box_rows)
int make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int box_rows, bool f16 = false) {
  const MapKey key{base, rows, cols, f16 ? -box_rows : box_rows};
  for (const MapSlot &s : tl_maps)
    if (s.key == key) {
      *map = s.map;
      return 0;
    }
  if (int rc = make_map_uncached(map, base, rows, cols, box_rows, f16)) return rc;
  if (tl_maps.size() >= 256) tl_maps.clear();   // workspaces were reallocated many times: start over
  tl_maps.push_back(MapSlot{key, *map});
  return 0;
}

int make_map_uncached(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int box_rows, bool f16) {
  EncodeTiledFn fn = encode_fn();
  B2M_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)box_rows};   // one 128-byte swizzle row of K
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 2;
  }
  return 0;
}

// ---------------------------------------------------------------- per-launch timing (bench.py's roofline)
struct Prof {
  bool on = false;
  // [0] = K5 (residual epilogue), [1] = K6 (gradient), [2] = NUTS state kernel, [3] = peer wait kernel, [4] = peer signal
  // kernel: start, stop, start, ...
  std::vector<cudaEvent_t> ev[5];
} g_prof;

template <int BLOCK_N, int MODE, int NCTA, bool F16>
int launch_tc_n(const CUtensorMap &Ah, const CUtensorMap &Al, const CUtensorMap &Bh, const CUtensorMap &Bl, dim3 grid,
                int kb_total, int kb_per_split, const EpiParams &E, cudaStream_t st, int chunk_kb, int mma_mask) {
  using C = Cfg<BLOCK_N, NCTA>;
  constexpr bool RESID = MODE == 1;
  auto kernel = tc_gemm_kernel<BLOCK_N, MODE, NCTA, F16>;
  constexpr int SMEM = C::SMEM_BYTES + (MODE != 0 ? NUM_EPI_WARPS * 32 * 33 * 4 : 0);
  // per device: the attribute belongs to the function on the current device.  Setting it again is harmless, so a plain
  // atomic flag is enough for concurrent first calls from several threads.
  static std::atomic<bool> configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    B2M_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  // `grid` arrives as (128-row tiles, column tiles, K splits); the launch is one persistent CTA group per SM (pair)
  const int Tm = (int)grid.x / NCTA, Tn = (int)grid.y, n_tiles = Tm * Tn * (int)grid.z;
  int n_groups = n_tiles < 148 / NCTA ? n_tiles : 148 / NCTA;
  {   // experiments: fewer persistent CTA groups (read once per process, glm.cuh)
    const int want = RESID ? tuning().groups_resid : tuning().groups_grad;
    if (want > 0 && want < n_groups) n_groups = want;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    B2M_CHECK_CUDA(cudaEventCreate(&e0));
    B2M_CHECK_CUDA(cudaEventCreate(&e1));
    B2M_CHECK_CUDA(cudaEventRecord(e0, st));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_groups * NCTA), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = NCTA == 2 ? 1 : 0;
  B2M_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, Ah, Al, Bh, Bl, kb_total, kb_per_split, chunk_kb, mma_mask, Tm, Tn, n_tiles, E));
  if (g_prof.on) {
    B2M_CHECK_CUDA(cudaEventRecord(e1, st));
    g_prof.ev[RESID ? 0 : 1].push_back(e0);
    g_prof.ev[RESID ? 0 : 1].push_back(e1);
  }
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BLOCK_N, int MODE>
int launch_tc(int ncta, const CUtensorMap &Ah, const CUtensorMap &Al, const CUtensorMap &Bh, const CUtensorMap &Bl,
              dim3 grid, int kb_total, int kb_per_split, const EpiParams &E, cudaStream_t st,
              int chunk_kb = DEFAULT_CHUNK_KB, int mma_mask = 7, bool f16 = false) {
  if constexpr (MODE == 2) {   // the push epilogue exists for the shape it is used with: fp16 encoding, CTA pairs
    B2M_REQUIRE(f16 && ncta == 2, "K6 push epilogue: needs the fp16 encoding and whole 256-row tiles");
    return launch_tc_n<BLOCK_N, 2, 2, true>(Ah, Al, Bh, Bl, grid, kb_total, kb_per_split, E, st, chunk_kb, mma_mask);
  } else {
    if (f16) {
      if (ncta == 2) return launch_tc_n<BLOCK_N, MODE, 2, true>(Ah, Al, Bh, Bl, grid, kb_total, kb_per_split, E, st, chunk_kb, mma_mask);
      return launch_tc_n<BLOCK_N, MODE, 1, true>(Ah, Al, Bh, Bl, grid, kb_total, kb_per_split, E, st, chunk_kb, mma_mask);
    }
    if (ncta == 2) return launch_tc_n<BLOCK_N, MODE, 2, false>(Ah, Al, Bh, Bl, grid, kb_total, kb_per_split, E, st, chunk_kb, mma_mask);
    return launch_tc_n<BLOCK_N, MODE, 1, false>(Ah, Al, Bh, Bl, grid, kb_total, kb_per_split, E, st, chunk_kb, mma_mask);
  }
}

}  // namespace

bool tc_available() { return encode_fn() != nullptr; }

void tc_profile(bool enable) {
  for (auto &v : g_prof.ev) {
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
  }
  g_prof.on = enable;
}

// bracket one launch of kind 2..4 (the per-chain kernels of the fused NUTS loop) with events while profiling is on
void prof_mark(int kind, cudaStream_t st) {
  if (!g_prof.on || kind < 0 || kind >= 5) return;
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_prof.ev[kind].push_back(e);
}

// out = {total ms, launches} per kind, n_kinds of them, since tc_profile(true); synchronises the device
int tc_profile_read_n(double *out, int n_kinds) {
  B2M_CHECK_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < n_kinds && k < 5; ++k) {
    double ms = 0.0;
    for (size_t i = 0; i + 1 < g_prof.ev[k].size(); i += 2) {
      float t = 0.f;
      B2M_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof.ev[k][i], g_prof.ev[k][i + 1]));
      ms += t;
    }
    out[2 * k] = ms;
    out[2 * k + 1] = (double)(g_prof.ev[k].size() / 2);
  }
  return 0;
}

// out4 = {K5 total ms, K5 launches, K6 total ms, K6 launches} since tc_profile(true); synchronises the device
int tc_profile_read(double *out4) {
  B2M_CHECK_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < 2; ++k) {
    double ms = 0.0;
    for (size_t i = 0; i + 1 < g_prof.ev[k].size(); i += 2) {
      float t = 0.f;
      B2M_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof.ev[k][i], g_prof.ev[k][i + 1]));
      ms += t;
    }
    out4[2 * k] = ms;
    out4[2 * k + 1] = (double)(g_prof.ev[k].size() / 2);
  }
  return 0;
}

static Tuning &tuning_storage() {
  static Tuning t = [] {   // C++11 magic static: initialised once, thread safe
    Tuning x;
    auto geti = [](const char *name, int dflt) {
      const char *v = getenv(name);
      return v ? atoi(v) : dflt;
    };
    x.groups_resid = geti("B2M_TC_GROUPS_RESID", 0);
    x.groups_grad = geti("B2M_TC_GROUPS_GRAD", 0);
    x.chunk_resid = geti("B2M_TC_CHUNK_RESID", 0);
    x.chunk_grad = geti("B2M_TC_CHUNK_GRAD", 0);
    x.pair = geti("B2M_TC_PAIR", 1);
    x.l2_hints = geti("B2M_TC_L2_HINTS", 0);
    x.k6_order = geti("B2M_TC_K6_ORDER", 1);
    x.fuse = geti("B2M_TC_FUSE", 0);
    x.fuse_slab = geti("B2M_TC_FUSE_SLAB", 0);
    x.fuse_ring = geti("B2M_TC_FUSE_RING", 0);
    x.fuse_groups5 = geti("B2M_TC_FUSE_GROUPS5", 0);
    x.fuse_hints5 = geti("B2M_TC_FUSE_HINTS5", -1);
    x.fuse_hints6 = geti("B2M_TC_FUSE_HINTS6", -1);
    return x;
  }();
  return t;
}
const Tuning &tuning() { return tuning_storage(); }

// tests / experiments: change one knob of the process (not thread safe against running launches; b2m_tuning_set)
int tuning_set(const char *name, int value) {
  Tuning &t = tuning_storage();
  struct { const char *n; int *p; } tab[] = {
      {"groups_resid", &t.groups_resid}, {"groups_grad", &t.groups_grad}, {"chunk_resid", &t.chunk_resid},
      {"chunk_grad", &t.chunk_grad}, {"pair", &t.pair}, {"l2_hints", &t.l2_hints}, {"k6_order", &t.k6_order}, {"fuse", &t.fuse},
      {"fuse_slab", &t.fuse_slab}, {"fuse_ring", &t.fuse_ring}, {"fuse_groups5", &t.fuse_groups5},
      {"fuse_hints5", &t.fuse_hints5}, {"fuse_hints6", &t.fuse_hints6}};
  for (auto &e : tab)
    if (name && !strcmp(name, e.n)) { *e.p = value; return 0; }
  set_error(std::string("b2m_tuning_set: unknown knob '") + (name ? name : "(null)") + "'");
  return 2;
}

int grad_block_n(const GlmModel &g) { return g.Dp % 256 == 0 ? 256 : (g.Dp % 128 == 0 ? 128 : 64); }

// CTA pairs need whole 256-row tiles of chains; B2M_TC_PAIR=0 forces the single-CTA kernel (experiments)
static int pair_mode(int64_t Cp) {
  if (tuning().pair == 0) return 1;
  return (Cp % 256 == 0) ? 2 : 1;
}

// split-K factor of K6: enough CTAs to fill the 148 SMs in whole waves, each split at least 8 k-blocks deep
int grad_splits(const GlmModel &g, int64_t Cp) {
  const int64_t tiles = (Cp / BLOCK_M) * (g.Dp / grad_block_n(g));   // CTAs per split
  const int kb = g.Np / (g.use_tc == 2 ? 64 : 32);
  int max_s = kb / 8 > 0 ? kb / 8 : 1;
  if (max_s > 32) max_s = 32;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_s; ++s) {
    const int64_t ctas = tiles * s;
    const int64_t waves = (ctas + 147) / 148;
    // efficiency of the last wave, minus a small charge per split for writing / re-reading the partials
    const double eff = (double)ctas / (double)(waves * 148) - 0.004 * (s - 1);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

int tc_gemm_resid(GlmModel &g, int64_t Cp, cudaStream_t st) {
  const int ncta = pair_mode(Cp);
  const bool f16 = g.use_tc == 2;
  const int bk = f16 ? 64 : 32;
  const void *A_h = f16 ? (const void *)g.B16h : (const void *)g.Bh, *A_l = f16 ? (const void *)g.B16l : (const void *)g.Bl;
  const void *B_h = f16 ? (const void *)g.X16h : (const void *)g.Xh, *B_l = f16 ? (const void *)g.X16l : (const void *)g.Xl;
  CUtensorMap Ah, Al, Bh, Bl;
  if (make_map(&Ah, A_h, Cp, g.Dp, BLOCK_M, f16) || make_map(&Al, A_l, Cp, g.Dp, BLOCK_M, f16) ||
      make_map(&Bh, B_h, g.Np, g.Dp, 256 / ncta, f16) || make_map(&Bl, B_l, g.Np, g.Dp, 256 / ncta, f16))
    return 2;
  EpiParams E{};
  E.y = g.y0; E.inv_var = g.inv_var; E.ss_part = g.ss_part; E.Cp = Cp; E.Np = g.Np;
  E.skip_flag = g.skip_flag; E.skip_target = g.skip_target;
  E.l2_hints = tuning().l2_hints ? 2 : 0;   // A (packed positions, re-read by every column tile): evict_last
  E.Rh = f16 ? (void *)g.R16h : (void *)g.Rh;
  E.Rl = f16 ? (void *)g.R16l : (void *)g.Rl;
  E.a_unscale = g.a_unscale; E.r_scale = g.r_scale;
  E.N_valid = g.N; E.loc_const = 0.f; E.weight = g.weight;   // y0 is already centred and shifted (glm.cu)
  dim3 grid((unsigned)(Cp / BLOCK_M), g.Np / 256, 1);
  // promotion interval: 128 values of K for tf32; 256 for fp16 -- the MMAs of a tile take half as long there, and the
  // longer chunk lets the issuer run far enough into the next tile to cover the epilogue (measured at C4: 4.28 ->
  // 4.12 ms per evaluation, gradient error 1.3e-6 -> 2.1e-6; K5's accumulators only hold X (beta - beta0))
  // fp16 encoding with K >= 512: two chunks of 8 k-blocks (512 values of K) per tile instead of four of 4 -- TMEM holds two
  // chunk accumulators, so the issuer can then run a WHOLE tile ahead of the residual epilogue instead of half a tile
  // (measured at C4: K5 1.94 -> 1.89 ms; gradient error vs float64 unchanged at 2e-6, full-size test green)
  const int ck = tuning().chunk_resid > 0 ? tuning().chunk_resid : ((f16 && g.Dp / bk >= 8) ? 8 : DEFAULT_CHUNK_KB);
  return launch_tc<256, 1>(ncta, Ah, Al, Bh, Bl, grid, g.Dp / bk, g.Dp / bk, E, st, ck, 7, f16);
}

int tc_gemm_grad(GlmModel &g, int64_t Cp, cudaStream_t st) {
  const int ncta = pair_mode(Cp);
  const bool f16 = g.use_tc == 2;
  const int bk = f16 ? 64 : 32;
  const int bn = grad_block_n(g);
  const int splits = grad_splits(g, Cp);
  const int kb_total = g.Np / bk;
  const int kb_per = (kb_total + splits - 1) / splits;
  const void *A_h = f16 ? (const void *)g.R16h : (const void *)g.Rh, *A_l = f16 ? (const void *)g.R16l : (const void *)g.Rl;
  const void *B_h = f16 ? (const void *)g.XT16h : (const void *)g.XTh, *B_l = f16 ? (const void *)g.XT16l : (const void *)g.XTl;
  CUtensorMap Ah, Al, Bh, Bl;
  if (make_map(&Ah, A_h, Cp, g.Np, BLOCK_M, f16) || make_map(&Al, A_l, Cp, g.Np, BLOCK_M, f16) ||
      make_map(&Bh, B_h, g.Dp, g.Np, bn / ncta, f16) || make_map(&Bl, B_l, g.Dp, g.Np, bn / ncta, f16))
    return 2;
  EpiParams E{};
  E.Gpart = g.G; E.Cp = Cp; E.Dp = g.Dp;
  E.skip_flag = g.skip_flag; E.skip_target = g.skip_target;
  E.l2_hints = tuning().l2_hints ? (1 | 2 << 2) : 0;   // A (residual stream): evict_first, B (transposed design matrix): evict_last
  E.nt_fast = tuning().k6_order;
  dim3 grid((unsigned)(Cp / BLOCK_M), g.Dp / bn, (unsigned)((kb_total + kb_per - 1) / kb_per));
  g.g_splits = (int)grid.z;
  const int ck = tuning().chunk_grad > 0 ? tuning().chunk_grad : (f16 ? DEFAULT_CHUNK_KB / 2 : DEFAULT_CHUNK_KB);
  if (bn == 256) return launch_tc<256, 0>(ncta, Ah, Al, Bh, Bl, grid, kb_total, kb_per, E, st, ck, 7, f16);
  if (bn == 128) return launch_tc<128, 0>(ncta, Ah, Al, Bh, Bl, grid, kb_total, kb_per, E, st, ck, 7, f16);
  return launch_tc<64, 0>(ncta, Ah, Al, Bh, Bl, grid, kb_total, kb_per, E, st, ck, 7, f16);
}

// ---------------------------------------------------------------- K5 || K6 in one launch
namespace {
struct FusePlan {
  bool on = false;
  int slab = 0, ring = 0, n_slabs = 0, groups5 = 0, n_pairs = 0;
};
// The concurrent launch pays when the residual operand is far larger than L2 (it then costs two passes over HBM) and
// there are enough slabs for the one-slab lag between the roles to be small against the whole evaluation.
FusePlan fuse_plan(const GlmModel &g, int64_t Cp) {
  FusePlan p;
  const Tuning &T = tuning();
  if (!T.fuse || g.use_tc != 2 || pair_mode(Cp) != 2 || g.Dp % 256 != 0) return p;
  const int tn5 = g.Np / 256;
  // slab: about 8 MB of residuals (hi + lo halves) for the batch, between 1 and 8 observation tiles, in a ring of two.
  // Measured at C4 (profiles/r02_tc_gemm_fused_c4_ncu_summary.md): a 17 MB ring stays in L2 (DRAM traffic of the
  // evaluation 6.3 -> 1.1 GB), 25 MB begins to spill (2.4 GB), 50 MB spills completely -- next to the 16 MB of packed
  // positions and the 16 MB gradient buffer the two-die L2 keeps about half of its nominal 126 MB for this pattern.
  int slab = T.fuse_slab > 0 ? T.fuse_slab : (int)((8ll << 20) / (Cp * 256 * 4));
  if (slab < 1) slab = 1;
  if (slab > 8 && T.fuse_slab <= 0) slab = 8;
  const int ring = T.fuse_ring >= 2 ? T.fuse_ring : 2;
  const int n_slabs = (tn5 + slab - 1) / slab;
  const int64_t r_bytes = Cp * (int64_t)g.Np * 4;
  if (n_slabs < 4 * ring || (r_bytes < (256ll << 20) && T.fuse < 2)) return p;   // B2M_TC_FUSE=2: whenever the shape allows
  p.n_pairs = 148 / 2;
  // the K6 role pays an epilogue (transposition + reductions at L2) per slab instead of per split: it gets the larger half
  p.groups5 = T.fuse_groups5 > 0 && T.fuse_groups5 < p.n_pairs ? T.fuse_groups5 : p.n_pairs / 2 - 2;
  p.slab = slab; p.ring = ring; p.n_slabs = n_slabs;
  p.on = true;
  return p;
}
}  // namespace

static int tc_gemm_fused(GlmModel &g, int64_t Cp, const FusePlan &P, cudaStream_t st) {
  if (g.h_fz_err && *g.h_fz_err) {
    set_error("concurrent K5 || K6 launch: a dependency wait timed out in an earlier evaluation (results are invalid)");
    return 2;
  }
  const int bk = 64, kb5 = g.Dp / bk, tn5 = g.Np / 256, tn6 = g.Dp / 256, Tm = (int)(Cp / 256);
  const int pitch = P.ring * P.slab * 256;
  // counters [ready: Tm x n_slabs | turn: Tm x tn6 | err]; the error word is cleared once, the counters before every launch
  const size_t n_cnt = (size_t)Tm * P.n_slabs + (size_t)Tm * tn6;
  if (g.fz_sync_cap < n_cnt + 1) {
    if (g.fz_sync) { B2M_CHECK_CUDA(cudaStreamSynchronize(st)); cudaFree(g.fz_sync); g.fz_sync = nullptr; g.fz_sync_cap = 0; }
    B2M_CHECK_CUDA(cudaMalloc(reinterpret_cast<void **>(&g.fz_sync), sizeof(int) * (n_cnt + 1 + 1024)));
    g.fz_sync_cap = n_cnt + 1 + 1024;
    B2M_CHECK_CUDA(cudaMemsetAsync(g.fz_sync, 0, sizeof(int) * g.fz_sync_cap, st));
  }
  if (!g.h_fz_err) {
    int *hp = nullptr;
    B2M_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&hp), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *hp = 0;
    g.h_fz_err = hp;
  }
  int *err_dev = g.fz_sync + g.fz_sync_cap - 1;
  B2M_CHECK_CUDA(cudaMemsetAsync(g.fz_sync, 0, sizeof(int) * n_cnt, st));
  CUtensorMap a5h, a5l, b5h, b5l, a6h, a6l, b6h, b6l;
  if (make_map(&a5h, g.B16h, Cp, g.Dp, BLOCK_M, true) || make_map(&a5l, g.B16l, Cp, g.Dp, BLOCK_M, true) ||
      make_map(&b5h, g.X16h, g.Np, g.Dp, 128, true) || make_map(&b5l, g.X16l, g.Np, g.Dp, 128, true) ||
      make_map(&a6h, g.R16h, Cp, pitch, BLOCK_M, true) || make_map(&a6l, g.R16l, Cp, pitch, BLOCK_M, true) ||
      make_map(&b6h, g.XT16h, g.Dp, g.Np, 128, true) || make_map(&b6l, g.XT16l, g.Dp, g.Np, 128, true))
    return 2;
  const Tuning &T = tuning();
  EpiParams E5{};
  E5.y = g.y0; E5.inv_var = g.inv_var; E5.ss_part = g.ss_part; E5.Cp = Cp; E5.Np = g.Np;
  E5.skip_flag = g.skip_flag; E5.skip_target = g.skip_target;
  E5.Rh = g.R16h; E5.Rl = g.R16l;
  E5.a_unscale = g.a_unscale; E5.r_scale = g.r_scale;
  E5.N_valid = g.N; E5.loc_const = 0.f; E5.weight = g.weight;
  E5.fz_slab = P.slab; E5.fz_ring = P.ring; E5.fz_pitch = pitch; E5.fz_nslabs = P.n_slabs; E5.fz_tn5 = tn5; E5.fz_tn6 = tn6;
  E5.fz_ready = g.fz_sync; E5.fz_turn = g.fz_sync + (size_t)Tm * P.n_slabs;
  E5.fz_err = err_dev; E5.fz_err_host = g.h_fz_err;
  EpiParams E6 = E5;
  E6.Gpart = g.G; E6.Dp = g.Dp;
  E6.nt_fast = T.k6_order;
  // K5: A = packed positions (16 MB, re-read for every observation tile) -> evict_last, B = design matrix, read once while
  // the sixteen pair rows are on the tile -> evict_first.  K6: no hints (the ring and the slab of the transposed matrix are
  // re-read within a slab time and then dead).
  E5.l2_hints = T.fuse_hints5 >= 0 ? T.fuse_hints5 : (2 | 1 << 2);
  E6.l2_hints = T.fuse_hints6 >= 0 ? T.fuse_hints6 : 0;
  g.g_splits = 1;
  const int chunk5 = T.chunk_resid > 0 ? T.chunk_resid : (kb5 >= 8 ? 8 : DEFAULT_CHUNK_KB);
  const int chunk6 = T.chunk_grad > 0 ? T.chunk_grad : DEFAULT_CHUNK_KB / 2;
  const int kb6_total = g.Np / bk, kb6_per = P.slab * 4;
  auto kernel = tc_gemm_fused_kernel;
  using C = Cfg<256, 2>;
  constexpr int SMEM = C::SMEM_BYTES + NUM_EPI_WARPS * 32 * 33 * 4;
  static std::atomic<bool> configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    B2M_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    B2M_CHECK_CUDA(cudaEventCreate(&e0));
    B2M_CHECK_CUDA(cudaEventCreate(&e1));
    B2M_CHECK_CUDA(cudaEventRecord(e0, st));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(P.n_pairs * 2), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2M_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, a5h, a5l, b5h, b5l, a6h, a6l, b6h, b6l, kb5, chunk5, Tm, kb6_total, kb6_per,
                                    chunk6, P.groups5, E5, E6));
  if (g_prof.on) {   // one launch covers both contractions: recorded under K5, no K6 entry (bench.py reads it that way)
    B2M_CHECK_CUDA(cudaEventRecord(e1, st));
    g_prof.ev[0].push_back(e0);
    g_prof.ev[0].push_back(e1);
  }
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int tc_gemm_resid_grad(GlmModel &g, int64_t Cp, cudaStream_t st) {
  const FusePlan P = fuse_plan(g, Cp);
  if (P.on) return tc_gemm_fused(g, Cp, P, st);
  if (int rc = tc_gemm_resid(g, Cp, st)) return rc;
  return tc_gemm_grad(g, Cp, st);
}

// K6 of a peer-sliced observation shard: one K pass over the rank's rows (no split-K: every output tile is produced
// once and leaves through the epilogue), tiles pushed into the owners' windows.
int tc_gemm_grad_push(GlmModel &g, int64_t Cp, cudaStream_t st) {
  const PeerWindow &w = g.pw;
  B2M_REQUIRE(w.nranks >= 2 && g.use_tc == 2 && Cp == w.C && Cp % 256 == 0 && w.own % 128 == 0,
              "K6 push epilogue: needs an attached peer window, the fp16 encoding and a full batch");
  const int bk = 64, bn = grad_block_n(g);
  const int kb_total = g.Np / bk;
  CUtensorMap Ah, Al, Bh, Bl;
  if (make_map(&Ah, g.R16h, Cp, g.Np, BLOCK_M, true) || make_map(&Al, g.R16l, Cp, g.Np, BLOCK_M, true) ||
      make_map(&Bh, g.XT16h, g.Dp, g.Np, bn / 2, true) || make_map(&Bl, g.XT16l, g.Dp, g.Np, bn / 2, true))
    return 2;
  EpiParams E{};
  E.Cp = Cp; E.Dp = g.Dp;
  E.l2_hints = tuning().l2_hints ? (1 | 2 << 2) : 0;
  E.nt_fast = tuning().k6_order;
  E.push_own = (int)w.own; E.push_rank = w.rank;
  for (int s = 0; s < w.nranks; ++s) E.push_dst[s] = reinterpret_cast<float *>(w.base[s] + w.off_g);
  E.r_unscale = g.r_unscale; E.inv_col_scale = g.inv_col_scale;
  dim3 grid((unsigned)(Cp / BLOCK_M), g.Dp / bn, 1);
  g.g_splits = 1;
  const int ck = tuning().chunk_grad > 0 ? tuning().chunk_grad : DEFAULT_CHUNK_KB / 2;
  if (bn == 256) return launch_tc<256, 2>(2, Ah, Al, Bh, Bl, grid, kb_total, kb_total, E, st, ck, 7, true);
  if (bn == 128) return launch_tc<128, 2>(2, Ah, Al, Bh, Bl, grid, kb_total, kb_total, E, st, ck, 7, true);
  return launch_tc<64, 2>(2, Ah, Al, Bh, Bl, grid, kb_total, kb_total, E, st, ck, 7, true);
}


// ---------------------------------------------------------------- raw 3xTF32 GEMM for experiments / tests
__global__ void split_kernel(const float *__restrict__ x, int64_t n, float *hi, float *lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) split_tf32(x[i], hi[i], lo[i]);
}

int debug_tc_gemm(const float *A, const float *Bm, int M, int N, int K, float *Cout, int chunk_kb, int mma_mask,
                  cudaStream_t st) {
  B2M_REQUIRE(M % BLOCK_M == 0 && N % 64 == 0 && K % BLOCK_K == 0, "debug_tc_gemm: M%128, N%64, K%32 must be 0");
  float *Ah, *Al, *Bh, *Bl;
  B2M_CHECK_CUDA(cudaMalloc(&Ah, sizeof(float) * M * (size_t)K));
  B2M_CHECK_CUDA(cudaMalloc(&Al, sizeof(float) * M * (size_t)K));
  B2M_CHECK_CUDA(cudaMalloc(&Bh, sizeof(float) * N * (size_t)K));
  B2M_CHECK_CUDA(cudaMalloc(&Bl, sizeof(float) * N * (size_t)K));
  split_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, (int64_t)M * K, Ah, Al);
  split_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(Bm, (int64_t)N * K, Bh, Bl);
  const int bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);
  const int ncta = pair_mode(M);
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc = make_map(&mAh, Ah, M, K, BLOCK_M) || make_map(&mAl, Al, M, K, BLOCK_M) || make_map(&mBh, Bh, N, K, bn / ncta) ||
           make_map(&mBl, Bl, N, K, bn / ncta);
  if (!rc) {
    EpiParams E{};
    E.Gpart = Cout; E.Cp = M; E.Dp = N;
    dim3 grid(M / BLOCK_M, N / bn, 1);
    if (bn == 256) rc = launch_tc<256, 0>(ncta, mAh, mAl, mBh, mBl, grid, K / BLOCK_K, K / BLOCK_K, E, st, chunk_kb, mma_mask);
    else if (bn == 128) rc = launch_tc<128, 0>(ncta, mAh, mAl, mBh, mBl, grid, K / BLOCK_K, K / BLOCK_K, E, st, chunk_kb, mma_mask);
    else rc = launch_tc<64, 0>(ncta, mAh, mAl, mBh, mBl, grid, K / BLOCK_K, K / BLOCK_K, E, st, chunk_kb, mma_mask);
  }
  cudaStreamSynchronize(st);
  cudaFree(Ah); cudaFree(Al); cudaFree(Bh); cudaFree(Bl);
  return rc ? 2 : 0;
}

}  // namespace b2m
