// Helpers shared by the pointwise-class kernels (pointwise.cu, nuts_pointwise.cu).
#pragma once
#ifndef __CUDACC_RTC__
#include <math.h>
#endif

#include "model.cuh"

namespace b2m {

// ---------------------------------------------------------------- small helpers
template <int DMAX>
struct Vec {
  float v[DMAX];
};

template <int DMAX>
__device__ __forceinline__ void to_mailbox(const float (&q)[DMAX], float *th, int TS, int D) {
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) th[d * TS] = q[d];
}
template <int DMAX>
__device__ __forceinline__ void from_mailbox(float (&g)[DMAX], const float *gr, int TS, int D) {
#pragma unroll
  for (int d = 0; d < DMAX; ++d) g[d] = (d < D) ? gr[d * TS] : 0.f;
}

template <int DMAX>
__device__ __forceinline__ void store_vec(float *dst, const float (&q)[DMAX], int D) {
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) dst[d] = q[d];
}

// kinetic energy exactly as the reference sums it: 0.5 * (p_0^2 + p_1^2 + ...) left to right
// (hmc.py:110, nuts.py:116)
template <int DMAX>
__device__ __forceinline__ float kinetic(const float (&p)[DMAX], int D) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) s = __fadd_rn(s, __fmul_rn(p[d], p[d]));
  return __fmul_rn(0.5f, s);
}

// momentum / proposal normals for one (chain, iteration): injected or Philox (slot map in common.cuh)
template <int DMAX>
__device__ __forceinline__ void draw_normals(float (&z)[DMAX], int D, const float *inj, uint64_t seed,
                                             uint64_t gchain, uint32_t giter, uint4 first) {
  if (inj) {
#pragma unroll
    for (int d = 0; d < DMAX; ++d) z[d] = (d < D) ? inj[d] : 0.f;
    return;
  }
  float a, b;
  box_muller(first.x, first.y, a, b);
  z[0] = a;
  if (DMAX > 1) z[1] = b;
#pragma unroll
  for (int k = 0; k < (DMAX + 1) / 4; ++k) {
    if (2 + 4 * k < D) {
      uint4 w = Philox::draw(seed, gchain, giter, 1u + k);
      float n0, n1, n2, n3;
      box_muller(w.x, w.y, n0, n1);
      box_muller(w.z, w.w, n2, n3);
      if (2 + 4 * k + 0 < DMAX) z[2 + 4 * k + 0] = n0;
      if (2 + 4 * k + 1 < DMAX) z[2 + 4 * k + 1] = n1;
      if (2 + 4 * k + 2 < DMAX) z[2 + 4 * k + 2] = n2;
      if (2 + 4 * k + 3 < DMAX) z[2 + 4 * k + 3] = n3;
    }
  }
}

struct Lane;
template <int DMAX, bool COMPACT, bool GRAD>
__device__ __forceinline__ float evaluate(const KModel &km, const SModel &sm, const Lane &L, const float (&q)[DMAX],
                                          float (&g)[DMAX]);

struct Lane {
  int64_t chain;   // clamped local chain index
  bool writer;     // lane 0 of an in-range chain
  int lane, G, TS;
  unsigned gmask;  // warp lanes serving this chain
  float *th, *gr;  // mailbox columns
};

__device__ __forceinline__ Lane make_lane(int64_t n_chains, int G, unsigned char *mail, int dmax) {
  Lane L;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = t / G;
  L.lane = int(t % G);
  L.G = G;
  L.TS = blockDim.x;
  L.gmask = G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  L.writer = (c < n_chains) && (L.lane == 0);
  L.chain = c < n_chains ? c : n_chains - 1;
  float *m = reinterpret_cast<float *>(mail);
  L.th = m + threadIdx.x;
  L.gr = m + (size_t)dmax * blockDim.x + threadIdx.x;
  return L;
}

// One fused value(+gradient): compact models evaluate from registers and the constant bank, the general path goes
// through the shared-memory mailbox.  Identical arithmetic either way.
template <int DMAX, bool COMPACT, bool GRAD>
__device__ __forceinline__ float evaluate_plain(const KModel &km, const SModel &sm, const Lane &L, const float (&q)[DMAX],
                                                float (&g)[DMAX]) {
  if constexpr (COMPACT) {
    return eval_model_c<GRAD, DMAX>(km, sm, q, g, L.lane, L.G, L.gmask);
  } else {
    to_mailbox<DMAX>(q, L.th, L.TS, sm.D);
    const float lp = eval_model<GRAD>(sm, L.th, L.gr, L.TS, L.lane, L.G, L.gmask);
    if (GRAD) from_mailbox<DMAX>(g, L.gr, L.TS, sm.D);
    return lp;
  }
}

// With constraint transforms the chain state `q` is the unconstrained coordinate u: the model is evaluated at
// theta = T(u), the gradient is chained through dT/du and the log-Jacobian joins value and gradient.  km.has_tf is a
// kernel-parameter constant: models without transforms (the reference's behaviour) take the first branch unchanged.
#ifdef B2M_JIT
// specialised kernels: the transform codes are literals of the generated translation unit
__device__ constexpr int jit_tf(int d) {
  constexpr int codes[16] = B2M_JIT_TF_LIST;
  return d < 16 ? codes[d] : 0;
}
#define B2M_HAS_TF(km) (B2M_JIT_HAS_TF != 0)
#define B2M_TF_CODE(km, d) jit_tf(d)
#else
#define B2M_HAS_TF(km) ((km).has_tf != 0)
#define B2M_TF_CODE(km, d) ((d) < 16 ? (int)(km).tf[d] : 0)
#endif

template <int DMAX, bool COMPACT, bool GRAD>
__device__ __forceinline__ float evaluate(const KModel &km, const SModel &sm, const Lane &L, const float (&q)[DMAX],
                                          float (&g)[DMAX]) {
  if (!B2M_HAS_TF(km)) return evaluate_plain<DMAX, COMPACT, GRAD>(km, sm, L, q, g);
  float th[DMAX], jac[DMAX], dlj[DMAX], lj = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d) {
    const Tf t = tf_apply(B2M_TF_CODE(km, d), q[d]);
    th[d] = t.theta; jac[d] = t.jac; dlj[d] = t.dlogjac;
    if (d < sm.D) lj += t.logjac;
  }
  const float lp = evaluate_plain<DMAX, COMPACT, GRAD>(km, sm, L, th, g);
  if (GRAD) {
#pragma unroll
    for (int d = 0; d < DMAX; ++d) g[d] = fmaf(g[d], jac[d], dlj[d]);
  }
  return lp + lj;
}

// a draw as the caller wants it: the constrained value unless the call asks for the sampler's own coordinates
template <int DMAX>
__device__ __forceinline__ void store_draw(float *dst, const float (&q)[DMAX], int D, const KModel &km, bool unconstrained) {
  if (!B2M_HAS_TF(km) || unconstrained) { store_vec<DMAX>(dst, q, D); return; }
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) dst[d] = tf_constrain(B2M_TF_CODE(km, d), q[d]);
}

// diagonal mass matrix (ABI 2): im = diag(M^-1) or all ones, sm_ = sqrt of the diagonal of M
template <int DMAX>
__device__ __forceinline__ void load_mass(const float *inv_mass, int D, float (&im)[DMAX], float (&sq)[DMAX]) {
#pragma unroll
  for (int d = 0; d < DMAX; ++d) {
    im[d] = (inv_mass && d < D) ? inv_mass[d] : 1.0f;
    sq[d] = (inv_mass && d < D) ? 1.0f / sqrtf(im[d]) : 1.0f;
  }
}
// kinetic energy 0.5 * sum p_d^2 / m_d, summed left to right; identical to kinetic() when im == 1
template <int DMAX>
__device__ __forceinline__ float kinetic_m(const float (&p)[DMAX], const float (&im)[DMAX], int D) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) s = __fadd_rn(s, __fmul_rn(__fmul_rn(p[d], p[d]), im[d]));
  return __fmul_rn(0.5f, s);
}

#ifndef __CUDACC_RTC__
// ---------------------------------------------------------------- host-side launchers
inline int pick_dmax(int D) { return D <= 2 ? 2 : D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 0; }

int pick_lanes(const KModel &km, int64_t n_chains, int requested);

// per-model specialised kernels (jit.cu): entry points of a loaded module, indexed in this order
struct JitModule;
enum { JIT_LOGP_GRAD = 0, JIT_HMC = 1, JIT_MH = 2, JIT_NUTS = 3 };
int jit_load(const void *image, int dmax, JitModule **out);
void jit_unload(JitModule *m);
int jit_dmax(const JitModule *m);
int jit_launch(JitModule *m, int which, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void **params);

struct Geometry {
  dim3 grid, block;
  size_t smem;
};

inline Geometry geometry(const KModel &km, int64_t n_chains, int G, int dmax) {
  Geometry ge;
  int threads = 64;
  if (threads < G) threads = G;
  const int64_t total = n_chains * G;
  ge.block = dim3(threads);
  ge.grid = dim3((unsigned)((total + threads - 1) / threads));
  ge.smem = model_smem_bytes(km) + (km.compact ? 0 : sizeof(float) * 2 * (size_t)dmax * threads);
  return ge;
}

template <typename K>
static int prep(K kernel, size_t smem) {
  if (smem > 48 * 1024) B2M_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

#define B2M_DISPATCH_DMAX(dmax, compact, ...)                                        \
  switch ((dmax) + ((compact) ? 100 : 0)) {                                           \
    case 102: { constexpr int DM = 2; constexpr bool CP = true; __VA_ARGS__; } break;  \
    case 104: { constexpr int DM = 4; constexpr bool CP = true; __VA_ARGS__; } break;  \
    case 2: { constexpr int DM = 2; constexpr bool CP = false; __VA_ARGS__; } break;   \
    case 4: { constexpr int DM = 4; constexpr bool CP = false; __VA_ARGS__; } break;   \
    case 8: { constexpr int DM = 8; constexpr bool CP = false; __VA_ARGS__; } break;   \
    case 16: { constexpr int DM = 16; constexpr bool CP = false; __VA_ARGS__; } break; \
    default: b2m::set_error("pointwise models support at most 16 scalar parameters"); return 1; \
  }
#endif  // !__CUDACC_RTC__

}  // namespace b2m
