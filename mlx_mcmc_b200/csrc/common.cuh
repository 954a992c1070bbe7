// Shared device helpers: Philox4x32-10 counter RNG, uniform/normal transforms, error plumbing.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#endif

#ifdef B2M_JIT
#include "b200mcmc.h"      // NVRTC: headers are handed over by name (mlx_mcmc_b200/jit.py)
#else
#include "../../include/b200mcmc.h"
#endif

#ifdef __CUDACC_RTC__
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#endif

// Index / invariant checks inside the kernels, compiled in by `python -m mlx_mcmc_b200.build --debug` (-DB2M_DEBUG_ASSERTS):
// compute-sanitizer is closed on the GPU pool, so the debug build carries its own bounds asserts (a failed one traps the
// kernel and the next CUDA call returns cudaErrorAssert).  tools/debug_asserts_smoke.py runs every path on that build.
#if defined(B2M_DEBUG_ASSERTS) && !defined(__CUDACC_RTC__)
#include <assert.h>
#define B2M_ASSERT(cond) assert(cond)
#else
#define B2M_ASSERT(cond) ((void)0)
#endif

namespace b2m {

#ifndef __CUDACC_RTC__
// ---------------------------------------------------------------- error plumbing (host)
void set_error(const std::string &msg);
extern std::atomic<int64_t> g_launches;   // kernel launches issued by this library (any thread)
#endif

#define B2M_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      b2m::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)

#define B2M_REQUIRE(cond, msg)                                                                 \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      b2m::set_error(msg);                                                                     \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based: the draw for (global chain id, iteration, slot) never depends on launch
// geometry, lanes-per-chain or GPU count.  counter = (chain_lo, chain_hi, iteration, slot).
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

  __host__ __device__ static inline uint4 draw(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t slot) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)chain, c1 = (uint32_t)(chain >> 32), c2 = iter, c3 = slot;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
      uint32_t n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
      uint32_t n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// 24-bit uniform strictly inside (0,1): log(u) is always finite.
__host__ __device__ inline float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

#ifdef __CUDACC__
// Box-Muller on two words -> two independent N(0,1).
__device__ inline void box_muller(uint32_t a, uint32_t b, float &z0, float &z1) {
  float r = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincospif(2.0f * u01(b), &s, &c);
  z0 = r * c;
  z1 = r * s;
}
#endif

// Slot map (documented in DESIGN.md):
//   slot 0              : words 0,1 -> N(0,1) for d = 0,1 ; word 2 -> accept / slice uniform ; word 3 spare
//   slot 1 + k (k>=0)   : four N(0,1) for d = 2+4k .. 5+4k
//   slot 0x4000 + j     : NUTS doubling j: word 0 direction, word 1 take-uniform
//   slot 0x8000 + 1024*j + m/4, word m%4 : NUTS merge uniform m of doubling j
constexpr uint32_t SLOT_NUTS_DOUBLING = 0x4000u;
constexpr uint32_t SLOT_NUTS_MERGE = 0x8000u;

}  // namespace b2m
