// GLM-class models: log p(theta) = priors(theta) + w * sum_n Normal(y_n | c + (X @ beta)_n, sigma)
// with beta a slice of theta and sigma a constant or a scalar parameter.
//
// The gradient of the likelihood is two dense contractions shared by all chains of a lock-step batch:
//     M = B X^T     [C, D] x [D, N]   -> residual epilogue  R = w (y - c - M) / sigma_c^2,  ss_c = sum_n (y - c - M)^2
//     G = R X       [C, N] x [N, D]
// (K5 / K6 of SURVEY.md 2b).  Two implementations sit behind the same interface:
//   * fp32 SIMT tiles (glm_simt.cu)    -- any shape, the validation path
//   * TMA + tcgen05 3xTF32 (glm_tc.cu) -- B, X, R split into tf32 hi + lo parts, fp32 accumulation in TMEM
#pragma once
#include <cuda_fp16.h>

#include "model.cuh"

namespace b2m {

struct Comm;  // comm.cu
int comm_unique_id(uint8_t *out128);
int comm_init(const uint8_t *id128, int nranks, int rank, Comm **out);
void comm_destroy(Comm *c);
int comm_nranks(const Comm *c);
int comm_rank(const Comm *c);
int comm_allgather_inplace(Comm *c, void *buf, int64_t count, int elt_bytes, cudaStream_t st);
int comm_reducescatter_f32_inplace(Comm *c, float *buf, int64_t count, cudaStream_t st);
int comm_allreduce_f32(Comm *c, float *buf, int64_t n, cudaStream_t st);
int comm_allreduce_i64(Comm *c, int64_t *buf, int64_t n, cudaStream_t st);
int comm_allreduce_f32_max(Comm *c, float *buf, int64_t n, cudaStream_t st);

constexpr int kCenterSlices = 32;  // chain slices of the deterministic two-stage mean (glm.cu, centring)
constexpr int kMaxPeers = 8;       // ranks of one NVSwitch box

// The peer window of observation sharding with sliced state (include/b200mcmc.h, B2M_SLICE_PEER).  Every rank owns one
// window with this layout; `base[s]` is rank s's window as mapped into this process (CUDA IPC), base[rank] is local.
struct PeerWindow {
  int nranks = 0, rank = 0;
  int64_t C = 0, own = 0;          // chains of the call, chains per owner (C / nranks)
  int Dp = 0;
  char *base[kMaxPeers] = {};
  // byte offsets inside a window
  size_t off_g = 0;      // float  [nranks][own][Dp]  gradient partial of MY chains from each source rank (unscaled)
  size_t off_ss = 0;     // float  [nranks][own]      sum z^2 partial of my chains from each source rank
  size_t off_bh = 0, off_bl = 0;   // __half [C][Dp]  packed position rows of ALL chains (each owner writes its rows everywhere)
  size_t off_meta = 0;   // float4 [C]                (1/sigma^2, 1 / row scale of delta', ||delta||^2, spare) per row
  size_t off_bflag = 0;  // u64 [nranks]  "rank s has published the rows of its chains for sequence number n"
  size_t off_gflag = 0;  // u64 [nranks]  "rank s has stored its gradient / sum z^2 partials of sequence number n"
  size_t off_done = 0;   // i64 [nranks]  finished chains of rank s's slice
  size_t off_err = 0;    // int           a spin-wait timed out (the call then fails instead of hanging the GPU)
  size_t bytes = 0;
  uint64_t seq = 0;      // sequence number of the last exchange (identical on every rank: the calls are collective)
};
size_t peer_layout(PeerWindow &w, int64_t C, int nranks, int Dp);

struct GlmModel {
  // problem
  int N = 0, D = 0;        // observations, regression coefficients
  int Np = 0, Dp = 0;      // padded: Np % 256 == 0, Dp % 64 == 0 (zero padded)
  int Dtot = 0;            // full parameter count
  int beta_off = 0;        // beta = theta[beta_off : beta_off + D]
  int sigma_param = -1;    // index of sigma in theta, or -1 when constant
  float sigma_const = 1.f, loc_const = 0.f, weight = 1.f;
  // observed data in HBM (owned by the handle)
  float *X = nullptr, *XT = nullptr;                         // [Np, Dp], [Dp, Np] fp32
  float *Xh = nullptr, *Xl = nullptr, *XTh = nullptr, *XTl = nullptr;  // tf32 hi / lo splits
  float *y = nullptr;                                        // [Np]
  // fp16 hi/lo encoding (use_tc == 2): X' = X * col_scale[d] (power of two, column maximum just below 2^14) split
  // into two halves, hi = fp16(x'), lo = fp16(x' - hi); same for the transposed copy
  __half *X16h = nullptr, *X16l = nullptr, *XT16h = nullptr, *XT16l = nullptr;
  float *col_scale = nullptr, *inv_col_scale = nullptr;      // [Dp]
  float *colmax = nullptr;                                   // [Dp] column maxima of |X| the scales were derived from
  float x_rownorm_max = 0.f;                                 // max_n ||X_n||_2: Cauchy-Schwarz bound of |(X delta)_n|
  unsigned *y0max_bits = nullptr;                            // device scalar: max_n |y0_n| as float bits (per recentre)
  // centring (glm.cu): beta0 = mean current position of the batch, y0 = y - c - X beta0 (float64 accumulation)
  float *y0 = nullptr, *beta0 = nullptr;                     // [Np], [Dp]
  double *center_part = nullptr;                             // [32, Dp]
  // prior terms (everything except the matvec likelihood) as a pointwise term table
  KModel prior{};
  bool has_prior = false;
  // workspace sized for `cap` chains (padded to 128)
  int64_t cap = 0;
  float *B = nullptr, *Bh = nullptr, *Bl = nullptr;          // [cap, Dp]
  float *R = nullptr, *Rh = nullptr, *Rl = nullptr;          // [cap, Np]
  __half *B16h = nullptr, *B16l = nullptr;                   // [cap, Dp]  fp16 encoding of (beta - beta0) / col_scale * row scale
  __half *R16h = nullptr, *R16l = nullptr;                   // [cap, Np]  fp16 encoding of R * row scale
  float *a_unscale = nullptr, *r_scale = nullptr, *r_unscale = nullptr;   // [cap] per-row scales of the two A operands
  float *G = nullptr;                                        // [g_splits_cap, cap, Dp] split-K partials of the gradient
  int g_splits = 1, g_splits_cap = 1;
  float *ss_part = nullptr;                                  // [Np / 64, cap] per-column-tile partial sum of squares
  float *inv_var = nullptr;                                  // [cap]
  int use_tc = 0;                                            // 0: SIMT fp32 tiles, 1: tcgen05 3xTF32, 2: tcgen05 3xFP16
  // sampler workspace (glm_samplers.cu): one arena reused across calls, plus a pinned host flag
  char *ws = nullptr;
  size_t ws_cap = 0;
  int *h_flag = nullptr, *h_ring = nullptr;   // pinned: one flag + a ring of lagged counters
  const int *skip_flag = nullptr;             // device counter; the tensor-core contractions return at once when it has
  int skip_target = 0;                        // reached skip_target (set by the asynchronous NUTS loop only)
  // observation sharding (comm.cu): this handle holds rows [r0, r0 + N) of a model with N_total rows
  Comm *comm = nullptr;
  int64_t N_total = 0;
  float *red = nullptr;                                      // [cap * Dp + cap] gradient partial || sum z^2, all-reduced
  PeerWindow pw;                                             // attached by b2m_model_peer_attach (nranks == 0: none)
  // constraint transforms (include/b200mcmc.h): per-parameter B2M_TF_* codes, device [Dtot], or nullptr; the samplers'
  // theta is then the unconstrained coordinate and `thc` [cap, Dtot] holds T(theta) of the batch being evaluated
  int *tf = nullptr;
  float *thc = nullptr;
  // K6 push epilogue (peer window): set only while a peer-sliced NUTS call is running
  bool push_on = false;
  unsigned *blk_counter = nullptr;                           // device: "last block" counters of the signalling kernels
  volatile long long *h_prog = nullptr;                      // pinned + mapped: {ticks completed, finished chains} written by the device
  // concurrent K5 || K6 launch (glm_tc.cu, tc_gemm_resid_grad): dependency counters [ready | turn] in device memory, zeroed
  // before every launch, and a pinned + mapped flag a timed-out dependency wait raises (the next call then fails)
  int *fz_sync = nullptr;
  size_t fz_sync_cap = 0;
  volatile int *h_fz_err = nullptr;
};

// knobs for experiments, read ONCE per process from the environment (never on the launch path):
//   B2M_TC_GROUPS_RESID / B2M_TC_GROUPS_GRAD  persistent CTA groups, B2M_TC_CHUNK_RESID / B2M_TC_CHUNK_GRAD promotion
//   interval in k-blocks, B2M_TC_PAIR=0 single-CTA kernels, B2M_TC_L2_HINTS=1 L2 eviction-policy hints on the TMA loads of
//   the separate kernels (measured: no gain, off by default)
//   B2M_TC_K6_ORDER=0 K6 tiles with the pair row fastest (round-1 order) instead of the column tile
//   B2M_TC_FUSE=0 (default) separate K5 / K6 launches; 1 the concurrent launch for large fp16-encoded problems (a fifth
//   of the DRAM traffic, ~10-20 % slower: DESIGN.md 4.2), 2 whenever the shape allows; B2M_TC_FUSE_SLAB 256-observation
//   tiles per slab of the concurrent launch, B2M_TC_FUSE_RING slabs in the residual ring, B2M_TC_FUSE_GROUPS5 CTA pairs given to K5,
//   B2M_TC_FUSE_HINTS5 / B2M_TC_FUSE_HINTS6 eviction-policy codes (A | B << 2; 0 none, 1 evict_first, 2 evict_last)
struct Tuning {
  int groups_resid = 0, groups_grad = 0, chunk_resid = 0, chunk_grad = 0, pair = 1, l2_hints = 0, k6_order = 1;
  int fuse = 0, fuse_slab = 0, fuse_ring = 0, fuse_groups5 = 0, fuse_hints5 = -1, fuse_hints6 = -1;
};
const Tuning &tuning();
int tuning_set(const char *name, int value);

int glm_set_comm(GlmModel &g, Comm *c, cudaStream_t st);

int glm_reserve(GlmModel &g, int64_t n_chains);
void glm_free(GlmModel &g);

// log p and gradient for `theta` [C, Dtot] (device).  grad may be NULL.  recenter: move the reference point of
// the contraction to the mean of `theta` first (the samplers do it once per iteration, from the current states).
// idx / n_rows: evaluate only the chains idx[0..n_rows) (a compacted lock-step batch); outputs land at the chains' rows.
// own_count > 0 (observation sharding with sliced state, full batch only): the partial gradients are reduce-scattered
// and only the rows [own_base, own_base + own_count) -- this rank's chains -- are finished.
int glm_logp_grad(GlmModel &g, const float *theta, int64_t C, float *logp, float *grad, cudaStream_t st,
                  bool recenter = false, const int *idx = nullptr, int64_t n_rows = 0, int64_t own_base = 0,
                  int64_t own_count = 0);
int glm_recenter(GlmModel &g, const float *theta, int64_t C, cudaStream_t st);

// the two contractions (SIMT implementation)
int simt_gemm_resid(GlmModel &g, int64_t Cp, cudaStream_t st);   // B -> R, ss_part
int simt_gemm_grad(GlmModel &g, int64_t Cp, cudaStream_t st);    // R -> G
// tcgen05 implementation (glm_tc.cu)
int tc_gemm_resid(GlmModel &g, int64_t Cp, cudaStream_t st);
int tc_gemm_grad(GlmModel &g, int64_t Cp, cudaStream_t st);
// K5 then K6 of one evaluation: ONE launch in which half of the CTA pairs run K5 and the other half K6 a slab of
// observations behind, the residual operand living in an L2-resident ring (large fp16-encoded problems), else the two
// launches above
int tc_gemm_resid_grad(GlmModel &g, int64_t Cp, cudaStream_t st);
bool tc_available();
void tc_profile(bool enable);
int tc_profile_read(double *out4);
int tc_profile_read_n(double *out, int n_kinds);
void prof_mark(int kind, cudaStream_t st);   // kinds 2 state, 3 peer wait, 4 peer signal: call before and after the launch
int grad_splits(const GlmModel &g, int64_t Cp);

int tc_gemm_grad_push(GlmModel &g, int64_t Cp, cudaStream_t st);   // K6 whose epilogue stores into the owners' windows
int glm_finish_launch(GlmModel &g, const float *theta, int64_t C, float *logp, float *grad, cudaStream_t st,
                      const int *idx, int64_t n_rows);
int glm_hmc_run(GlmModel &g, const b2m_hmc_args &a, cudaStream_t st);
int glm_nuts_run(GlmModel &g, const b2m_nuts_args &a, cudaStream_t st);
int glm_mh_run(GlmModel &g, const b2m_mh_args &a, cudaStream_t st);

// round-to-nearest split of an fp32 value into two tf32-representable parts, x ~= hi + lo
__host__ __device__ inline void split_tf32(float x, float &hi, float &lo) {
  union { float f; uint32_t u; } a, b;
  a.f = x;
  a.u = (a.u + 0x00001000u) & 0xffffe000u;
  hi = a.f;
  b.f = x - hi;
  b.u = (b.u + 0x00001000u) & 0xffffe000u;
  lo = b.f;
}

}  // namespace b2m
