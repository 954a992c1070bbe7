// GLM-class evaluation pipeline: pack -> GEMM (B X^T) with residual epilogue -> GEMM (R X) -> finish.
#include "glm.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace b2m {

// ---------------------------------------------------------------- memory
template <typename T>
static int dev_alloc(T **p, size_t n) {
  B2M_CHECK_CUDA(cudaMalloc(reinterpret_cast<void **>(p), sizeof(T) * (n ? n : 1)));
  return 0;
}
#define FREE(p) do { if (p) cudaFree(p); p = nullptr; } while (0)

static void free_workspace(GlmModel &g) {
  FREE(g.B); FREE(g.Bh); FREE(g.Bl); FREE(g.R); FREE(g.Rh); FREE(g.Rl); FREE(g.G); FREE(g.ss_part); FREE(g.inv_var);
  FREE(g.red); FREE(g.B16h); FREE(g.B16l); FREE(g.R16h); FREE(g.R16l); FREE(g.a_unscale); FREE(g.r_scale); FREE(g.r_unscale);
  g.cap = 0;
}

void glm_free(GlmModel &g) {
  free_workspace(g);
  FREE(g.X); FREE(g.XT); FREE(g.Xh); FREE(g.Xl); FREE(g.XTh); FREE(g.XTl); FREE(g.y);
  FREE(g.y0); FREE(g.beta0); FREE(g.center_part);
  FREE(g.X16h); FREE(g.X16l); FREE(g.XT16h); FREE(g.XT16l); FREE(g.col_scale); FREE(g.inv_col_scale); FREE(g.y0max_bits);
  FREE(g.ws);
  g.ws_cap = 0;
  if (g.h_flag) cudaFreeHost(g.h_flag);
  g.h_flag = nullptr;
}

int glm_reserve(GlmModel &g, int64_t n_chains) {
  const int64_t cp = (n_chains + 255) / 256 * 256;   // whole CTA-pair tiles (glm_tc.cu)
  if (cp <= g.cap) return 0;
  free_workspace(g);
  // split-K partials: a compacted batch of fewer rows may use more splits -- size G for the worst case
  size_t g_rows = (size_t)cp;
  if (g.use_tc)
    for (int64_t c = 128; c <= cp; c += 128) {
      const size_t need = (size_t)grad_splits(g, c) * (size_t)c;
      if (need > g_rows) g_rows = need;
    }
  g.g_splits_cap = (int)((g_rows + cp - 1) / cp);
  if (dev_alloc(&g.B, cp * g.Dp) || dev_alloc(&g.G, g_rows * g.Dp) || dev_alloc(&g.ss_part, (size_t)(g.Np / 64) * cp) ||
      dev_alloc(&g.inv_var, cp) || dev_alloc(&g.red, (size_t)cp * g.Dp + cp))
    return 2;
  if (g.use_tc == 0 && dev_alloc(&g.R, cp * (size_t)g.Np)) return 2;
  if (g.use_tc == 1) {
    if (dev_alloc(&g.Bh, cp * g.Dp) || dev_alloc(&g.Bl, cp * g.Dp) || dev_alloc(&g.Rh, cp * (size_t)g.Np) ||
        dev_alloc(&g.Rl, cp * (size_t)g.Np))
      return 2;
  }
  if (g.use_tc == 2) {
    if (dev_alloc(&g.B16h, cp * g.Dp) || dev_alloc(&g.B16l, cp * g.Dp) || dev_alloc(&g.R16h, cp * (size_t)g.Np) ||
        dev_alloc(&g.R16l, cp * (size_t)g.Np) || dev_alloc(&g.a_unscale, cp) || dev_alloc(&g.r_scale, cp) ||
        dev_alloc(&g.r_unscale, cp))
      return 2;
  }
  g.cap = cp;
  return 0;
}

// ---------------------------------------------------------------- data preparation (once per model)
__global__ void pad_transpose_kernel(const float *__restrict__ X, int N, int D, int Np, int Dp, float *Xp, float *XT) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Np * Dp) return;
  const int n = int(i / Dp), d = int(i % Dp);
  const float v = (n < N && d < D) ? X[(int64_t)n * D + d] : 0.f;
  Xp[i] = v;
  if (XT) XT[(int64_t)d * Np + n] = v;
}

__global__ void split_tf32_kernel(const float *__restrict__ Xp, int Np, int Dp, float *Xh, float *Xl, float *XTh, float *XTl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Np * Dp) return;
  const int n = int(i / Dp), d = int(i % Dp);
  float hi, lo;
  split_tf32(Xp[i], hi, lo);
  Xh[i] = hi; Xl[i] = lo;
  XTh[(int64_t)d * Np + n] = hi; XTl[(int64_t)d * Np + n] = lo;
}

__global__ void pad_vector_kernel(const float *__restrict__ y, int N, int Np, float *yp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) yp[i] = i < N ? y[i] : 0.f;
}

// fp16 encoding: column maxima and the largest row norm of the padded X
__global__ void __launch_bounds__(256) x_stats_kernel(const float *__restrict__ Xp, int Np, int Dp, unsigned *colmax_bits,
                                                      unsigned *rownorm_bits) {
  // block (32, 8): 32 columns x 8 row lanes, rows strided by gridDim.y * 8
  const int d = blockIdx.x * 32 + threadIdx.x;
  float cm = 0.f;
  for (int n = blockIdx.y * 8 + threadIdx.y; n < Np; n += gridDim.y * 8) cm = fmaxf(cm, fabsf(Xp[(int64_t)n * Dp + d]));
  atomicMax(colmax_bits + d, __float_as_uint(cm));   // non-negative floats order like their bit patterns
  if (blockIdx.x == 0) {                             // row norms: one warp (threadIdx.y) per row, lanes stride the columns
    for (int n = blockIdx.y * 8 + threadIdx.y; n < Np; n += gridDim.y * 8) {
      float s = 0.f;
      for (int c = threadIdx.x; c < Dp; c += 32) { const float v = Xp[(int64_t)n * Dp + c]; s = fmaf(v, v, s); }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) atomicMax(rownorm_bits, __float_as_uint(sqrtf(s)));
    }
  }
}

// Entries of a column below its maximum x 2^-17 land, after the column is scaled to just under 2^14, where the lo half
// of the fp16 split is subnormal: they keep fewer than the nominal 22 bits.  Count them (zeros are exact and excluded).
__global__ void __launch_bounds__(256) x_small_count_kernel(const float *__restrict__ Xp, int Np, int Dp,
                                                            const unsigned *__restrict__ colmax_bits, unsigned *small) {
  const int d = blockIdx.x * 32 + threadIdx.x;
  const float thr = __uint_as_float(colmax_bits[d]) * 7.62939453125e-6f;   // 2^-17
  unsigned cnt = 0;
  for (int n = blockIdx.y * 8 + threadIdx.y; n < Np; n += gridDim.y * 8) {
    const float a = fabsf(Xp[(int64_t)n * Dp + d]);
    cnt += (a > 0.f && a < thr) ? 1u : 0u;
  }
  if (cnt) atomicAdd(small + d, cnt);
}

// power-of-two scale that puts a maximum of `m` just below 2^14 (fp16 overflows at 65504 = 2^16 - 32)
__host__ __device__ inline float pow2_scale(float m) {
  if (!(m > 0.f) || !(m < 3.0e38f)) return 1.0f;
  int e;
  frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)  =>  m < 2^e
  int k = 14 - e;
  if (k > 100) k = 100;
  if (k < -100) k = -100;
  return ldexpf(1.0f, k);
}

__global__ void col_scale_kernel(const unsigned *__restrict__ colmax_bits, int Dp, float *col_scale, float *inv_col_scale) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= Dp) return;
  const float s = pow2_scale(__uint_as_float(colmax_bits[d]));
  col_scale[d] = s;
  inv_col_scale[d] = 1.0f / s;
}

__device__ __forceinline__ void split_f16(float v, __half &hi, __half &lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__global__ void split_f16_kernel(const float *__restrict__ Xp, int Np, int Dp, const float *__restrict__ col_scale,
                                 __half *Xh, __half *Xl, __half *XTh, __half *XTl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Np * Dp) return;
  const int n = int(i / Dp), d = int(i % Dp);
  __half hi, lo;
  split_f16(Xp[i] * col_scale[d], hi, lo);
  Xh[i] = hi; Xl[i] = lo;
  XTh[(int64_t)d * Np + n] = hi; XTl[(int64_t)d * Np + n] = lo;
}

int glm_build(GlmModel &g, const float *X, const float *y, int N, int D) {
  g.N = N; g.D = D;
  g.N_total = N;
  g.Np = (N + 255) / 256 * 256;
  g.Dp = (D + 63) / 64 * 64;
  const size_t nd = (size_t)g.Np * g.Dp;
  if (dev_alloc(&g.X, nd) || dev_alloc(&g.y, g.Np) || dev_alloc(&g.y0, g.Np) || dev_alloc(&g.beta0, g.Dp) ||
      dev_alloc(&g.center_part, (size_t)kCenterSlices * g.Dp) || dev_alloc(&g.y0max_bits, 1))
    return 2;
  if (g.use_tc == 0 && dev_alloc(&g.XT, nd)) return 2;
  pad_transpose_kernel<<<(unsigned)((nd + 255) / 256), 256>>>(X, N, D, g.Np, g.Dp, g.X, g.XT);
  pad_vector_kernel<<<(g.Np + 255) / 256, 256>>>(y, N, g.Np, g.y);
  g_launches += 2;
  if (g.use_tc == 2) {
    // column maxima and the largest row norm; then, per column, how many entries are so far below the column's
    // maximum (< 2^-17 of it) that the scaled fp16 split cannot hold them to full precision: data with more than
    // 0.1 % of such entries in some column (one wild outlier is enough) keeps the tf32 encoding (8-bit exponent)
    unsigned *stats = nullptr;   // [Dp] column maxima (float bits) + [1] largest row norm + [Dp] small-entry counts
    if (dev_alloc(&stats, 2 * g.Dp + 1)) return 2;
    B2M_CHECK_CUDA(cudaMemset(stats, 0, sizeof(unsigned) * (2 * g.Dp + 1)));
    x_stats_kernel<<<dim3(g.Dp / 32, 64), dim3(32, 8)>>>(g.X, g.Np, g.Dp, stats, stats + g.Dp);
    x_small_count_kernel<<<dim3(g.Dp / 32, 64), dim3(32, 8)>>>(g.X, g.Np, g.Dp, stats, stats + g.Dp + 1);
    g_launches += 2;
    std::vector<unsigned> hmax(2 * g.Dp + 1);
    B2M_CHECK_CUDA(cudaMemcpy(hmax.data(), stats, sizeof(unsigned) * (2 * g.Dp + 1), cudaMemcpyDeviceToHost));
    memcpy(&g.x_rownorm_max, &hmax[g.Dp], sizeof(float));
    bool ok = g.x_rownorm_max < 3.0e38f;
    for (int d = 0; d < g.Dp && ok; ++d) {
      float m;
      memcpy(&m, &hmax[d], sizeof(float));
      if (!(m < 3.0e38f)) ok = false;                                  // inf / NaN in the data
      if ((double)hmax[g.Dp + 1 + d] > 1e-3 * (double)(N > 0 ? N : 1)) ok = false;
    }
    const char *force = getenv("B2M_GLM_PATH");
    if (!ok && !(force && std::string(force) == "tc16")) g.use_tc = 1;   // wide-range data: tf32 encoding
    if (g.use_tc == 2) {
      if (dev_alloc(&g.X16h, nd) || dev_alloc(&g.X16l, nd) || dev_alloc(&g.XT16h, nd) || dev_alloc(&g.XT16l, nd) ||
          dev_alloc(&g.col_scale, g.Dp) || dev_alloc(&g.inv_col_scale, g.Dp))
        return 2;
      col_scale_kernel<<<(g.Dp + 127) / 128, 128>>>(stats, g.Dp, g.col_scale, g.inv_col_scale);
      split_f16_kernel<<<(unsigned)((nd + 255) / 256), 256>>>(g.X, g.Np, g.Dp, g.col_scale, g.X16h, g.X16l, g.XT16h, g.XT16l);
      g_launches += 2;
    }
    B2M_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(stats);
  }
  if (g.use_tc == 1) {
    if (dev_alloc(&g.Xh, nd) || dev_alloc(&g.Xl, nd) || dev_alloc(&g.XTh, nd) || dev_alloc(&g.XTl, nd)) return 2;
    split_tf32_kernel<<<(unsigned)((nd + 255) / 256), 256>>>(g.X, g.Np, g.Dp, g.Xh, g.Xl, g.XTh, g.XTl);
    ++g_launches;
  }
  B2M_CHECK_CUDA(cudaGetLastError());
  B2M_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

// ---------------------------------------------------------------- centering (once per sampler iteration)
// Near the posterior mode the residual z = y - X beta is a small difference of large numbers, and the gradient
// X^T z amplifies any systematic error of X beta by N (the tensor core accumulates with truncation: measured
// 1e-6 relative shrinkage of X beta => 1e-6 * N * |beta| absolute error in the gradient).  The contraction is
// therefore taken around a reference point beta0 shared by all chains of the batch:
//     z_c = (y - c - X beta0) - X (beta_c - beta0)
// with y0 = y - c - X beta0 accumulated in float64 once per sampler iteration (one GEMV, X read once) and
// beta0 = the mean of the chains' current positions.  X (beta_c - beta0) is then of the size of the
// posterior spread, and its rounding no longer dominates z.  Same function, different association.

__global__ void __launch_bounds__(256) glm_center_partial_kernel(const float *__restrict__ theta, int64_t C, int Dtot,
                                                                 int beta_off, int D, int Dp, double *__restrict__ part) {
  __shared__ double sh[8][33];
  const int d = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const int64_t per = (C + kCenterSlices - 1) / kCenterSlices;
  const int64_t c0 = (int64_t)blockIdx.y * per, c1 = (c0 + per < C) ? c0 + per : C;
  double acc = 0.0;
  if (d < D)
    for (int64_t c = c0 + ty; c < c1; c += 8) {
      const float v = theta[c * Dtot + beta_off + d];
      if (fabsf(v) <= 3.0e38f) acc += (double)v;   // a diverged / NaN chain must not move the centre
    }
  sh[ty][threadIdx.x] = acc;
  __syncthreads();
  if (ty == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x];   // fixed order => deterministic
    part[(int64_t)blockIdx.y * Dp + d] = s;
  }
}

__global__ void glm_center_final_kernel(const double *__restrict__ part, int Dp, int64_t C, float *__restrict__ beta0) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= Dp) return;
  double s = 0.0;
  for (int i = 0; i < kCenterSlices; ++i) s += part[(int64_t)i * Dp + d];
  beta0[d] = (float)(s / (double)C);
}

// y0[n] = (y[n] - c) - sum_d X[n, d] beta0[d], float64 accumulation, one warp per observation row
__global__ void __launch_bounds__(256) glm_y0_kernel(const float *__restrict__ X, const float *__restrict__ y,
                                                     const float *__restrict__ beta0, int N, int Np, int Dp,
                                                     float loc_const, float *__restrict__ y0,
                                                     unsigned *__restrict__ y0max_bits) {
  extern __shared__ float sb[];
  for (int d = threadIdx.x; d < Dp; d += blockDim.x) sb[d] = beta0[d];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= Np) return;
  double acc = 0.0;
  if (n < N) {
    const float4 *row = reinterpret_cast<const float4 *>(X + (int64_t)n * Dp);
    const float4 *b4 = reinterpret_cast<const float4 *>(sb);
    for (int i = lane; i < Dp / 4; i += 32) {
      const float4 x = __ldg(row + i), b = b4[i];
      acc += (double)x.x * b.x + (double)x.y * b.y + (double)x.z * b.z + (double)x.w * b.w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float v = n < N ? (float)(((double)y[n] - (double)loc_const) - acc) : 0.f;
    y0[n] = v;
    if (v == v) atomicMax(y0max_bits, __float_as_uint(fabsf(v)));
  }
}

int glm_recenter(GlmModel &g, const float *theta, int64_t C, cudaStream_t st) {
  dim3 gp(g.Dp / 32, kCenterSlices), bp(32, 8);
  glm_center_partial_kernel<<<gp, bp, 0, st>>>(theta, C, g.Dtot, g.beta_off, g.D, g.Dp, g.center_part);
  glm_center_final_kernel<<<(g.Dp + 127) / 128, 128, 0, st>>>(g.center_part, g.Dp, C, g.beta0);
  cudaMemsetAsync(g.y0max_bits, 0, sizeof(unsigned), st);
  glm_y0_kernel<<<(g.Np + 7) / 8, 256, sizeof(float) * g.Dp, st>>>(g.X, g.y, g.beta0, g.N, g.Np, g.Dp, g.loc_const, g.y0,
                                                                   g.y0max_bits);
  g_launches += 3;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- pack: theta -> B = beta - beta0 (+ tf32 split), 1/sigma^2
// `idx` (optional) lists the chains of a compacted batch: row c of the batch is chain idx[c] of `theta`.
__global__ void glm_pack_kernel(const float *__restrict__ theta, const float *__restrict__ beta0,
                                const int *__restrict__ idx, int64_t C, int64_t Cp,
                                int Dtot, int beta_off, int D, int Dp, int sigma_param, float sigma_const, float *B,
                                float *Bh, float *Bl, float *inv_var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cp * Dp) return;
  const int64_t row = i / Dp;
  const int d = int(i % Dp);
  const int64_t c = (row < C && idx) ? idx[row] : row;
  const float v = (row < C && d < D) ? __fsub_rn(theta[c * Dtot + beta_off + d], beta0[d]) : 0.f;
  B[i] = v;
  if (Bh) {
    float hi, lo;
    split_tf32(v, hi, lo);
    Bh[i] = hi; Bl[i] = lo;
  }
  if (d == 0) {
    const float s = (row < C && sigma_param >= 0) ? theta[c * Dtot + sigma_param] : sigma_const;
    inv_var[row] = 1.0f / (s * s);
  }
}

// fp16 encoding of the A operand of K5, one warp per chain row:
//   delta'_d = (beta_d - beta0_d) / col_scale_d          (X' = X col_scale, so delta' . X' = delta . X exactly)
//   row scale s_a = 2^k putting max_d |delta'_d| just below 2^14; hi / lo halves of delta' s_a
// and the row scale of the residual operand of K6 from a bound that is known before K5 runs:
//   |z_n| = |y0_n - (X delta)_n| <= max|y0| + ||delta||_2 max_n ||X_n||_2      (Cauchy-Schwarz)
__global__ void __launch_bounds__(128) glm_pack16_kernel(const float *__restrict__ theta, const float *__restrict__ beta0,
                                                         const int *__restrict__ idx, int64_t C, int64_t Cp, int Dtot,
                                                         int beta_off, int D, int Dp, int sigma_param, float sigma_const,
                                                         float weight, const float *__restrict__ inv_col_scale,
                                                         const unsigned *__restrict__ y0max_bits, float x_rownorm_max,
                                                         __half *Bh, __half *Bl, float *inv_var, float *a_unscale,
                                                         float *r_scale, float *r_unscale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= Cp) return;
  const bool live = row < C;
  const int64_t c = (live && idx) ? idx[row] : row;
  float amax = 0.f, n2 = 0.f;
  if (live)
    for (int d = lane; d < D; d += 32) {
      const float dl = __fsub_rn(theta[c * Dtot + beta_off + d], beta0[d]);
      amax = fmaxf(amax, fabsf(dl * inv_col_scale[d]));
      n2 = fmaf(dl, dl, n2);
    }
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  const float sa = pow2_scale(amax);
  for (int d = lane; d < Dp; d += 32) {
    float v = 0.f;
    if (live && d < D) v = __fsub_rn(theta[c * Dtot + beta_off + d], beta0[d]) * inv_col_scale[d] * sa;
    __half hi, lo;
    split_f16(v, hi, lo);
    Bh[row * Dp + d] = hi;
    Bl[row * Dp + d] = lo;
  }
  if (lane == 0) {
    const float sg = (live && sigma_param >= 0) ? theta[c * Dtot + sigma_param] : sigma_const;
    const float iv = 1.0f / (sg * sg);
    inv_var[row] = iv;
    const float bound = (__uint_as_float(*y0max_bits) + sqrtf(n2) * x_rownorm_max) * fabsf(iv * weight);
    const float sr = pow2_scale(bound);
    a_unscale[row] = 1.0f / sa;
    r_scale[row] = sr;
    r_unscale[row] = 1.0f / sr;
  }
}

// ---------------------------------------------------------------- SIMT contractions (fp32, 64x64x16 tiles)
// C[m, n] = sum_k A[m, k] * Bm[n, k]; both operands K-major; all dimensions padded to the tile.
constexpr int TM = 64, TN = 64, TK = 16;

template <bool RESID>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                        int K, int lda, int ldb, float *__restrict__ Cout, int ldc,
                                                        const float *__restrict__ y, const float *__restrict__ inv_var,
                                                        float *__restrict__ ss_part, int64_t Cp, int N_valid,
                                                        float loc_const, float weight) {
  __shared__ float sA[TK][TM + 4], sB[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  const int lr = tid / 4, lc = (tid % 4) * 4;  // each thread loads one float4 of A and one of Bm per k-step
  for (int k0 = 0; k0 < K; k0 += TK) {
    const float4 a = *reinterpret_cast<const float4 *>(A + (int64_t)(m0 + lr) * lda + k0 + lc);
    const float4 b = *reinterpret_cast<const float4 *>(Bm + (int64_t)(n0 + lr) * ldb + k0 + lc);
    sA[lc + 0][lr] = a.x; sA[lc + 1][lr] = a.y; sA[lc + 2][lr] = a.z; sA[lc + 3][lr] = a.w;
    sB[lc + 0][lr] = b.x; sB[lc + 1][lr] = b.y; sB[lc + 2][lr] = b.z; sB[lc + 3][lr] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sA[k][ty * 4 + i]; bv[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (RESID) {
    // epilogue: z = y - c - M ; R = w z / sigma^2 ; per-row partial sum of z^2 over this column tile
    __shared__ float red[TM][17];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      const float iv = inv_var[m] * weight;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        const float z = (n < N_valid) ? (y[n] - loc_const) - acc[i][j] : 0.f;
        s = fmaf(z, z, s);
        acc[i][j] = z * iv;
      }
      *reinterpret_cast<float4 *>(Cout + (int64_t)m * ldc + n0 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      red[ty * 4 + i][tx] = s;
    }
    __syncthreads();
    if (tid < TM) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) s += red[tid][j];
      ss_part[(int64_t)blockIdx.x * Cp + m0 + tid] = s;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4 *>(Cout + (int64_t)(m0 + ty * 4 + i) * ldc + n0 + tx * 4) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

int simt_gemm_resid(GlmModel &g, int64_t Cp, cudaStream_t st) {
  dim3 grid(g.Np / TN, (unsigned)(Cp / TM));
  simt_gemm_kernel<true><<<grid, 256, 0, st>>>(g.B, g.X, g.Dp, g.Dp, g.Dp, g.R, g.Np, g.y0, g.inv_var, g.ss_part, Cp,
                                                g.N, 0.f, g.weight);
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int simt_gemm_grad(GlmModel &g, int64_t Cp, cudaStream_t st) {
  dim3 grid(g.Dp / TN, (unsigned)(Cp / TM));
  simt_gemm_kernel<false><<<grid, 256, 0, st>>>(g.R, g.XT, g.Np, g.Np, g.Np, g.G, g.Dp, nullptr, nullptr, nullptr, Cp,
                                                 0, 0.f, 0.f);
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- observation sharding: reduce + all-reduce
// red[c, d] = sum over split-K partials of G ; red[Cp*Dp + c] = sum over column tiles of ss_part.  Fixed order.
__global__ void __launch_bounds__(256) glm_reduce_kernel(const float *__restrict__ G, int g_splits,
                                                         const float *__restrict__ ss_part, int n_tiles, int64_t Cp,
                                                         int Dp, const float *__restrict__ r_unscale,
                                                         const float *__restrict__ inv_col_scale, float *__restrict__ red) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nG = Cp * Dp;
  if (i < nG) {
    float v = 0.f;
    for (int s = 0; s < g_splits; ++s) v += G[(int64_t)s * nG + i];
    if (r_unscale) v *= r_unscale[i / Dp] * inv_col_scale[i % Dp];   // fp16 encoding: the ranks' scales differ
    red[i] = v;
  } else if (i < nG + Cp) {
    const int64_t c = i - nG;
    float v = 0.f;
    for (int t = 0; t < n_tiles; ++t) v += ss_part[(int64_t)t * Cp + c];
    red[i] = v;
  }
}

int glm_set_comm(GlmModel &g, Comm *c, cudaStream_t st) {
  g.comm = c;
  g.N_total = g.N;
  if (!c) return 0;
  int64_t *d = nullptr, h = g.N;
  B2M_CHECK_CUDA(cudaMalloc(&d, sizeof(int64_t)));
  B2M_CHECK_CUDA(cudaMemcpyAsync(d, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  int rc = comm_allreduce_i64(c, d, 1, st);
  if (!rc) {
    cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    g.N_total = h;
  }
  cudaFree(d);
  return rc;
}

// ---------------------------------------------------------------- finish: assemble log p and the full gradient
// One warp per chain.  Likelihood value from the partial sums (fixed summation order => deterministic),
// gradient of beta from G, gradient of sigma analytically, then the prior terms (generic densities)
// added on top.  n_tiles = number of 64-wide column tiles that wrote ss_part.
__global__ void __launch_bounds__(128) glm_finish_kernel(KModel prior, int has_prior, const float *__restrict__ theta,
                                                          int64_t row_base, int64_t C, int64_t Cp, int Dtot, int beta_off, int D, int Dp,
                                                          int sigma_param, float sigma_const, float weight, int N,
                                                          int n_tiles, const float *__restrict__ ss_part,
                                                          const float *__restrict__ G, int g_splits, float *__restrict__ logp,
                                                          float *__restrict__ grad, const int *__restrict__ idx,
                                                          const float *__restrict__ r_unscale,
                                                          const float *__restrict__ inv_col_scale) {
  extern __shared__ __align__(16) unsigned char smem[];
  SModel sm;
  sm.n_terms = 0;
  if (has_prior) model_to_smem(prior, smem, sm);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t c = row_base + (int64_t)blockIdx.x * (blockDim.x / 32) + warp;   // rows [row_base, C)
  if (c >= C) return;
  const int64_t src = idx ? idx[c] : c;       // chain served by row c of a compacted batch
  const float *th = theta + src * Dtot;
  float ss = 0.f;
  for (int t = lane; t < n_tiles; t += 32) ss += ss_part[(int64_t)t * Cp + c];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float sg = sigma_param >= 0 ? th[sigma_param] : sigma_const;
  const float iv = 1.0f / (sg * sg);
  float lp = weight * ((float)N * (-kHalfLog2Pi - logf(sg)) - 0.5f * ss * iv);
  float *gr = grad ? grad + src * Dtot : nullptr;
  // One warp per chain: with a scalar loop over d the warp walks D / 32 dependent iterations (80-120 us at D = 1000
  // whatever the number of chains: latency bound).  When the layout allows, each lane takes four consecutive
  // coefficients and all split-K partials of an iteration are independent 16-byte loads.
  const bool vec4 = gr && sigma_param < 0 && (Dtot & 3) == 0 && (beta_off & 3) == 0 && (D & 3) == 0 &&
                    ((reinterpret_cast<uintptr_t>(gr) | reinterpret_cast<uintptr_t>(G)) & 15) == 0;
  if (vec4) {
    const float ru = r_unscale ? r_unscale[c] : 1.0f;
    for (int d = 4 * lane; d < Dtot; d += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d >= beta_off && d < beta_off + D) {
        const float4 *g4 = reinterpret_cast<const float4 *>(G + (int64_t)c * Dp + (d - beta_off));
        const int64_t stride4 = (Cp * (int64_t)Dp) >> 2;
#pragma unroll 8
        for (int s = 0; s < g_splits; ++s) {   // fixed order, element by element the same sums as the scalar loop
          const float4 t = g4[(int64_t)s * stride4];
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        if (r_unscale) {
          const float4 ic = *reinterpret_cast<const float4 *>(inv_col_scale + (d - beta_off));
          v.x *= ru * ic.x; v.y *= ru * ic.y; v.z *= ru * ic.z; v.w *= ru * ic.w;
        }
      }
      *reinterpret_cast<float4 *>(gr + d) = v;
    }
    __syncwarp();
  } else if (gr) {
#pragma unroll 2
    for (int d = lane; d < Dtot; d += 32) {
      float v = 0.f;
      if (d >= beta_off && d < beta_off + D) {
#pragma unroll 8
        for (int s = 0; s < g_splits; ++s) v += G[((int64_t)s * Cp + c) * Dp + (d - beta_off)];  // fixed order
        if (r_unscale) v *= r_unscale[c] * inv_col_scale[d - beta_off];   // fp16 encoding: undo the operand scales
      }
      if (d == sigma_param) v = weight * (ss * iv - (float)N) / sg;
      gr[d] = v;
    }
    __syncwarp();
  }
  // priors: lanes stride the elements of each term; each element touches its own theta entries
  float pl = 0.f;
  for (int t = 0; t < sm.n_terms; ++t) {
    const b2m_term &T = sm.terms[t];
    float acc = 0.f, ax = 0.f, a0 = 0.f, a1 = 0.f;
    if (T.dist == B2M_NORMAL && T.x.kind == B2M_OP_PARAMVEC && T.p0.kind == B2M_OP_CONST && T.p1.kind == B2M_OP_CONST) {
      // the usual coefficient prior, sum Normal(loc, scale).log_prob(beta): constants hoisted out of the element loop
      // (same expressions as dist_eval, evaluated once)
      const float p0 = T.p0.c, p1 = T.p1.c, var = p1 * p1, base = -kHalfLog2Pi - logf(p1), w = T.weight;
      const float *__restrict__ xv = th + T.x.a;
      float *__restrict__ gv = gr ? gr + T.x.a : nullptr;
      if ((T.length & 3) == 0 && ((reinterpret_cast<uintptr_t>(xv) | reinterpret_cast<uintptr_t>(gv)) & 15) == 0) {
        for (int n = 4 * lane; n < T.length; n += 128) {   // same per-element expressions, four elements per lane
          const float4 x4 = *reinterpret_cast<const float4 *>(xv + n);
          const float z0 = x4.x - p0, z1 = x4.y - p0, z2 = x4.z - p0, z3 = x4.w - p0;
          acc += base - (0.5f * (z0 * z0)) / var;
          acc += base - (0.5f * (z1 * z1)) / var;
          acc += base - (0.5f * (z2 * z2)) / var;
          acc += base - (0.5f * (z3 * z3)) / var;
          if (gv) {
            float4 g4 = *reinterpret_cast<float4 *>(gv + n);
            g4.x += w * (-(z0 / var)); g4.y += w * (-(z1 / var)); g4.z += w * (-(z2 / var)); g4.w += w * (-(z3 / var));
            *reinterpret_cast<float4 *>(gv + n) = g4;
          }
        }
      } else {
#pragma unroll 4
        for (int n = lane; n < T.length; n += 32) {
          const float z = xv[n] - p0;
          acc += base - (0.5f * (z * z)) / var;
          if (gv) gv[n] += w * (-(z / var));
        }
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      pl += w * acc;
      __syncwarp();
      continue;
    }
    for (int n = lane; n < T.length; n += 32) {
      const float x = op_fetch(T.x, n, th, 1, sm), p0 = op_fetch(T.p0, n, th, 1, sm), p1 = op_fetch(T.p1, n, th, 1, sm);
      Elem e = dist_eval<true>(T.dist, x, p0, p1, T.k0, T.k1, T.k2);
      acc += e.lp;
      if (gr) {
        if (T.x.kind == B2M_OP_PARAM) ax += e.dx; else if (T.x.kind == B2M_OP_PARAMVEC) gr[T.x.a + n] += T.weight * e.dx;
        if (T.p0.kind == B2M_OP_PARAM) a0 += e.d0; else if (T.p0.kind == B2M_OP_PARAMVEC) gr[T.p0.a + n] += T.weight * e.d0;
        if (T.p1.kind == B2M_OP_PARAM) a1 += e.d1; else if (T.p1.kind == B2M_OP_PARAMVEC) gr[T.p1.a + n] += T.weight * e.d1;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    pl += T.weight * acc;
    if (gr && lane == 0) {
      if (T.x.kind == B2M_OP_PARAM) gr[T.x.a] += T.weight * ax;
      if (T.p0.kind == B2M_OP_PARAM) gr[T.p0.a] += T.weight * a0;
      if (T.p1.kind == B2M_OP_PARAM) gr[T.p1.a] += T.weight * a1;
    }
    __syncwarp();
  }
  if (lane == 0) logp[src] = lp + pl;
}

int glm_logp_grad(GlmModel &g, const float *theta, int64_t C, float *logp, float *grad, cudaStream_t st, bool recenter,
                  const int *idx, int64_t n_rows, int64_t own_base, int64_t own_count) {
  if (int rc = glm_reserve(g, C)) return rc;
  if (recenter)
    if (int rc = glm_recenter(g, theta, C, st)) return rc;
  if (idx) C = n_rows;   // compacted batch: rows 0..n_rows-1 are the chains idx[0..n_rows-1]
  const int64_t Cp = C <= 128 ? 128 : (C + 255) / 256 * 256;   // one 128-row tile, or whole 256-row CTA-pair tiles
  const int64_t tot = Cp * g.Dp;
  if (g.use_tc == 2)
    glm_pack16_kernel<<<(unsigned)((Cp + 3) / 4), 128, 0, st>>>(theta, g.beta0, idx, C, Cp, g.Dtot, g.beta_off, g.D, g.Dp,
                                                                g.sigma_param, g.sigma_const, g.weight, g.inv_col_scale,
                                                                g.y0max_bits, g.x_rownorm_max, g.B16h, g.B16l, g.inv_var,
                                                                g.a_unscale, g.r_scale, g.r_unscale);
  else
    glm_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(theta, g.beta0, idx, C, Cp, g.Dtot, g.beta_off, g.D, g.Dp,
                                                                    g.sigma_param, g.sigma_const, g.B, g.Bh, g.Bl, g.inv_var);
  ++g_launches;
  int n_tiles;
  if (g.use_tc) {
    if (int rc = tc_gemm_resid(g, Cp, st)) return rc;
    if (grad) if (int rc = tc_gemm_grad(g, Cp, st)) return rc;
    n_tiles = g.Np / 128;   // 256-wide tiles, two column halves each
  } else {
    if (int rc = simt_gemm_resid(g, Cp, st)) return rc;
    if (grad) if (int rc = simt_gemm_grad(g, Cp, st)) return rc;
    n_tiles = g.Np / TN;
  }
  const size_t smem = g.has_prior ? model_smem_bytes(g.prior) : 16;
  const float *ssp = g.ss_part, *Gp = g.G;
  int splits = g.use_tc ? g.g_splits : 1;
  const float *r_un = g.use_tc == 2 ? g.r_unscale : nullptr, *ics = g.use_tc == 2 ? g.inv_col_scale : nullptr;
  if (g.comm) {
    // observation shard: contiguous [Cp, Dp] gradient partial || [Cp] sum z^2, summed over ranks on this stream
    const int64_t n_red = Cp * g.Dp + Cp;
    glm_reduce_kernel<<<(unsigned)((n_red + 255) / 256), 256, 0, st>>>(g.G, grad ? splits : 0, g.ss_part, n_tiles, Cp, g.Dp, r_un, ics, g.red);
    ++g_launches;
    if (own_count > 0) {
      // sliced state: every rank only needs the sums of its own chains' rows -- two in-place reduce-scatters
      // (gradient block, sum-of-squares block) instead of the all-reduce
      const int nr = comm_nranks(g.comm);
      B2M_REQUIRE(!idx && Cp == C && C % nr == 0 && own_count == C / nr && own_base == own_count * comm_rank(g.comm),
                  "glm_logp_grad: sliced state needs a full batch whose rows divide evenly over the ranks");
      if (grad)
        if (int rc = comm_reducescatter_f32_inplace(g.comm, g.red, own_count * g.Dp, st)) return rc;
      if (int rc = comm_reducescatter_f32_inplace(g.comm, g.red + Cp * g.Dp, own_count, st)) return rc;
    } else {
      if (int rc = comm_allreduce_f32(g.comm, grad ? g.red : g.red + Cp * g.Dp, grad ? n_red : Cp, st)) return rc;
    }
    Gp = g.red; ssp = g.red + Cp * g.Dp; splits = 1; n_tiles = 1;
    r_un = nullptr; ics = nullptr;   // already unscaled before the sum over ranks
  }
  const int64_t row_base = (g.comm && own_count > 0) ? own_base : 0;
  const int64_t rows = (g.comm && own_count > 0) ? own_count : C;
  glm_finish_kernel<<<(unsigned)((rows + 3) / 4), 128, smem, st>>>(g.prior, g.has_prior ? 1 : 0, theta, row_base, row_base + rows, Cp,
                                                                    g.Dtot, g.beta_off, g.D, g.Dp, g.sigma_param, g.sigma_const,
                                                                    g.weight, (int)g.N_total, n_tiles, ssp, Gp, splits, logp, grad,
                                                                    idx, r_un, ics);
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2m
