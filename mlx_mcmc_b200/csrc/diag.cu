// posterior diagnostics on the device, so that draws [S, C, D] (16-40 GB at the large
// configurations) never have to travel to the host to be summarised.
//
// Replaces, per (chain, parameter) series:
//   compute_ess        examples/06_nuts_comparison.py:22-41   (mode 0)  and examples/02_hmc_comparison.py:111-128 (mode 1)
//   np.mean / np.std   mlx_mcmc/inference/mcmc.py:219-221     (pooled over chains in diag_params_kernel)
// and adds what the reference lists as planned (README.md:212-216): Geyer initial-positive-sequence ESS and the
// Gelman-Rubin R-hat across chains.  float32 draws, float64 accumulation.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b2m {

// One thread per series i = c * D + d; the s-th draw of the series is draws[s * CD + i] (coalesced across i).
__global__ void __launch_bounds__(128) diag_series_kernel(const float *__restrict__ draws, int64_t S, int64_t CD,
                                                          int ess_mode, float *__restrict__ mean_o,
                                                          float *__restrict__ var_o, float *__restrict__ ess_ref_o,
                                                          float *__restrict__ ess_geyer_o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CD) return;
  const float *x = draws + i;
  double sum = 0.0;
  for (int64_t s = 0; s < S; ++s) sum += (double)x[s * CD];
  const double m = sum / (double)S;
  double ss = 0.0;
  for (int64_t s = 0; s < S; ++s) {
    const double d = (double)x[s * CD] - m;
    ss += d * d;
  }
  const double var = ss / (double)S;   // np.var: ddof = 0
  mean_o[i] = (float)m;
  var_o[i] = (float)var;
  if (!ess_ref_o && !ess_geyer_o) return;

  auto autocov_sum = [&](int64_t lag) {   // sum_{s < S - lag} (x_s - m)(x_{s+lag} - m)
    double a = 0.0;
    for (int64_t s = 0; s + lag < S; ++s) a += ((double)x[s * CD] - m) * ((double)x[(s + lag) * CD] - m);
    return a;
  };

  if (ess_ref_o) {
    double ess = (double)S;
    if (var != 0.0) {
      const int64_t L = (S / 2 < 100) ? S / 2 : 100;
      double acc = 0.0;
      int64_t n_lags = 0;
      for (int64_t lag = 1; lag < L; ++lag) {
        const double r = autocov_sum(lag) / (double)(S - lag) / var;   // np.mean over the S - lag products
        acc += r;
        ++n_lags;
        if (r < 0.05 && (ess_mode == 0 || n_lags > 1)) break;
      }
      ess = (double)S / (1.0 + 2.0 * acc);
    }
    ess_ref_o[i] = (float)ess;
  }
  if (ess_geyer_o) {
    double ess = (double)S;
    if (S >= 4 && ss > 0.0) {
      double tau = -1.0;
      for (int64_t k = 0; k + 1 < S; k += 2) {
        const double pair = (k == 0 ? ss : autocov_sum(k)) / ss + autocov_sum(k + 1) / ss;   // rho_k + rho_{k+1}, 1/S norm.
        if (pair < 0.0) break;
        tau += 2.0 * pair;
      }
      if (tau < 1.0 / (double)S) tau = 1.0 / (double)S;
      ess = (double)S / tau;
    }
    ess_geyer_o[i] = (float)ess;
  }
}

// One block per parameter d: pooled mean / std over all chains and draws, R-hat, ESS summed over chains.
// out[d * 5 + {0..4}] = mean, std (ddof 0), R-hat, sum_c ESS (reference estimator), sum_c ESS (Geyer).
__global__ void __launch_bounds__(256) diag_params_kernel(const float *__restrict__ mean, const float *__restrict__ var,
                                                          const float *__restrict__ ess_ref,
                                                          const float *__restrict__ ess_geyer, int64_t S, int64_t C,
                                                          int64_t D, double *__restrict__ out) {
  __shared__ double red[5][256];
  const int64_t d = blockIdx.x;
  const int t = threadIdx.x;
  double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // sum mean, sum mean^2, sum var, sum ess_ref, sum ess_geyer
  for (int64_t c = t; c < C; c += 256) {
    const double mu = mean[c * D + d];
    a[0] += mu;
    a[1] += mu * mu;
    a[2] += var[c * D + d];
    if (ess_ref) a[3] += ess_ref[c * D + d];
    if (ess_geyer) a[4] += ess_geyer[c * D + d];
  }
  for (int k = 0; k < 5; ++k) red[k][t] = a[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {   // fixed-order tree: deterministic
    if (t < o)
      for (int k = 0; k < 5; ++k) red[k][t] += red[k][t + o];
    __syncthreads();
  }
  if (t == 0) {
    const double Cn = (double)C, Sn = (double)S;
    const double mbar = red[0][0] / Cn;
    double vm0 = red[1][0] / Cn - mbar * mbar;   // variance of the chain means, ddof 0
    if (vm0 < 0.0) vm0 = 0.0;
    const double w0 = red[2][0] / Cn;            // mean within-chain variance, ddof 0
    const double W = S > 1 ? w0 * Sn / (Sn - 1.0) : w0;
    const double b_over_s = C > 1 ? vm0 * Cn / (Cn - 1.0) : 0.0;
    out[d * 5 + 0] = mbar;
    out[d * 5 + 1] = sqrt(w0 + vm0);
    out[d * 5 + 2] = W > 0.0 ? sqrt(((Sn - 1.0) / Sn * W + b_over_s) / W) : nan("");
    out[d * 5 + 3] = red[3][0];
    out[d * 5 + 4] = red[4][0];
  }
}

int diag_series(const float *draws, int64_t S, int64_t C, int64_t D, int ess_mode, float *mean, float *var, float *ess_ref,
                float *ess_geyer, cudaStream_t st) {
  const int64_t CD = C * D;
  diag_series_kernel<<<(unsigned)((CD + 127) / 128), 128, 0, st>>>(draws, S, CD, ess_mode, mean, var, ess_ref, ess_geyer);
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int diag_params(const float *mean, const float *var, const float *ess_ref, const float *ess_geyer, int64_t S, int64_t C,
                int64_t D, double *out, cudaStream_t st) {
  diag_params_kernel<<<(unsigned)D, 256, 0, st>>>(mean, var, ess_ref, ess_geyer, S, C, D, out);
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- diagonal mass matrix from warm-up draws
// SURVEY.md 8(f) row 3 (reference: roadmap only, README.md:165,220; identity mass is hard-wired at nuts.py:113-117).
// draws [S, C, D] -> inv_mass[d] = regularised variance of coordinate d pooled over all S x C draws.
// Stage 1: block (32 coordinates x 8 row lanes) per (coordinate block, slice of the S x C rows): float64 sums of
// (x - x0) and (x - x0)^2 around the first row's value (no cancellation), reduced over the row lanes in fixed order.
// Stage 2: one thread per coordinate folds the slices in fixed order.  Deterministic; draws are read once, coalesced.
constexpr int kMassSlices = 64;

__global__ void __launch_bounds__(256) mass_partial_kernel(const float *__restrict__ draws, int64_t R, int64_t D,
                                                           double *__restrict__ part) {
  __shared__ double s1[8][33], s2[8][33];
  const int64_t d = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const int64_t per = (R + kMassSlices - 1) / kMassSlices;
  const int64_t r0 = (int64_t)blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
  double a = 0.0, b = 0.0;
  if (d < D) {
    const double x0 = (double)draws[d];
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const double v = (double)draws[r * D + d] - x0;
      if (v == v && fabs(v) < 1e30) { a += v; b += v * v; }   // a diverged chain must not poison the metric
    }
  }
  s1[ty][threadIdx.x] = a;
  s2[ty][threadIdx.x] = b;
  __syncthreads();
  if (ty == 0 && d < D) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ta += s1[i][threadIdx.x]; tb += s2[i][threadIdx.x]; }
    part[((int64_t)blockIdx.y * D + d) * 2 + 0] = ta;
    part[((int64_t)blockIdx.y * D + d) * 2 + 1] = tb;
  }
}

__global__ void mass_final_kernel(const double *__restrict__ part, int64_t R, int64_t D, float *__restrict__ inv_mass) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double a = 0.0, b = 0.0;
  for (int i = 0; i < kMassSlices; ++i) { a += part[((int64_t)i * D + d) * 2]; b += part[((int64_t)i * D + d) * 2 + 1]; }
  const double n = (double)R;
  double var = b / n - (a / n) * (a / n);
  if (!(var > 0.0)) var = 0.0;
  var = var * (n / (n + 5.0)) + 1e-3 * (5.0 / (n + 5.0));   // Stan's shrinkage towards unit scale
  inv_mass[d] = (float)var;
}

int mass_from_draws(const float *draws, int64_t S, int64_t C, int64_t D, float *inv_mass, cudaStream_t st) {
  double *part = nullptr;
  B2M_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&part), sizeof(double) * 2 * kMassSlices * D, st));
  const int64_t R = S * C;
  mass_partial_kernel<<<dim3((unsigned)((D + 31) / 32), kMassSlices), dim3(32, 8), 0, st>>>(draws, R, D, part);
  mass_final_kernel<<<(unsigned)((D + 127) / 128), 128, 0, st>>>(part, R, D, inv_mass);
  g_launches += 2;
  B2M_CHECK_CUDA(cudaGetLastError());
  B2M_CHECK_CUDA(cudaFreeAsync(part, st));
  return 0;
}

// ---------------------------------------------------------------- order statistics (MCMC.summary on device draws)
// Replaces np.median / np.percentile of mlx_mcmc/inference/mcmc.py:221-224 for draws that stay on the device.
// Radix select on the order-preserving integer image of the floats: three histogram passes (11 + 11 + 10 bits), all
// requested ranks resolved by every pass; between passes a one-block kernel walks the histogram of each rank and
// narrows its prefix.  x is read three times, coalesced; nothing is sorted or moved.
constexpr int kMaxRanks = 16;   // two neighbouring order statistics per quantile

struct SelectState {
  unsigned long long rank[kMaxRanks];   // remaining rank inside the current prefix
  unsigned prefix[kMaxRanks];           // high bits decided so far
  int n_ranks;
};

__device__ __forceinline__ unsigned float_key(float v) {
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending float order == ascending unsigned order
}
__device__ __forceinline__ float key_float(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// pass p: bits [shift, shift + nbits) of the keys whose higher bits equal a rank's prefix -> hist[rank][bin]
__global__ void __launch_bounds__(256) select_hist_kernel(const float *__restrict__ x, int64_t n, const SelectState *__restrict__ S,
                                                          int pass, unsigned long long *__restrict__ hist) {
  __shared__ unsigned sh[2048];
  __shared__ unsigned pre[kMaxRanks];
  const int nbits = pass == 2 ? 10 : 11, shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0), nbins = 1 << nbits;
  const int nr = pass == 0 ? 1 : S->n_ranks;
  if (threadIdx.x < kMaxRanks) pre[threadIdx.x] = S->prefix[threadIdx.x];
  for (int r = 0; r < nr; ++r) {
    // ranks sharing a prefix share a histogram: only the first of them counts
    bool dup = false;
    __syncthreads();
    if (pass > 0)
      for (int q = 0; q < r; ++q) dup = dup || pre[q] == pre[r];
    if (dup) continue;
    for (int i = threadIdx.x; i < nbins; i += 256) sh[i] = 0;
    __syncthreads();
    const unsigned want = pre[r];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
      const unsigned k = float_key(x[i]);
      if (pass == 0 || (k >> (shift + nbits)) == want) atomicAdd(&sh[(k >> shift) & (nbins - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += 256)
      if (sh[i]) atomicAdd(&hist[(size_t)r * 2048 + i], (unsigned long long)sh[i]);
  }
}

// one block: for every rank find the bin that holds it, extend the prefix, keep the rank inside the bin
__global__ void select_scan_kernel(SelectState *S, int pass, unsigned long long *hist) {
  __shared__ unsigned pre[kMaxRanks];
  const int nbits = pass == 2 ? 10 : 11, nbins = 1 << nbits;
  const int r = threadIdx.x, nr = S->n_ranks;
  if (r < kMaxRanks) pre[r] = S->prefix[r];
  __syncthreads();
  if (r < nr) {
    int src = r;                      // the histogram of the first rank with the same prefix (pass 0: rank 0's)
    if (pass == 0) src = 0;
    else
      for (int q = 0; q < r; ++q)
        if (pre[q] == pre[r]) { src = q; break; }
    const unsigned long long *h = hist + (size_t)src * 2048;
    unsigned long long rem = S->rank[r];
    int b = 0;
    for (; b < nbins - 1; ++b) {
      if (rem < h[b]) break;
      rem -= h[b];
    }
    S->rank[r] = rem;
    S->prefix[r] = (pass == 0 ? 0u : (pre[r] << nbits)) | (unsigned)b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxRanks * 2048; i += blockDim.x) hist[i] = 0;   // ready for the next pass
}

int quantiles(const float *x, int64_t n, const double *q, int n_q, double *out, cudaStream_t st) {
  SelectState hs{};
  double frac[8];
  hs.n_ranks = 2 * n_q;
  for (int i = 0; i < n_q; ++i) {   // numpy's default: linear interpolation at position q (n - 1)
    const double pos = q[i] * (double)(n - 1);
    const double lo = floor(pos);
    frac[i] = pos - lo;
    hs.rank[2 * i] = (unsigned long long)lo;
    hs.rank[2 * i + 1] = (unsigned long long)(lo + 1 < (double)n ? lo + 1 : lo);
  }
  SelectState *ds = nullptr;
  unsigned long long *hist = nullptr;
  B2M_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ds), sizeof(SelectState), st));
  B2M_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&hist), sizeof(unsigned long long) * kMaxRanks * 2048, st));
  B2M_CHECK_CUDA(cudaMemcpyAsync(ds, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
  B2M_CHECK_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * kMaxRanks * 2048, st));
  int blocks = (int)((n + 256 * 16 - 1) / (256 * 16));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  for (int pass = 0; pass < 3; ++pass) {
    select_hist_kernel<<<blocks, 256, 0, st>>>(x, n, ds, pass, hist);
    select_scan_kernel<<<1, 256, 0, st>>>(ds, pass, hist);
    g_launches += 2;
  }
  B2M_CHECK_CUDA(cudaGetLastError());
  B2M_CHECK_CUDA(cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, st));
  B2M_CHECK_CUDA(cudaStreamSynchronize(st));
  B2M_CHECK_CUDA(cudaFreeAsync(ds, st));
  B2M_CHECK_CUDA(cudaFreeAsync(hist, st));
  for (int i = 0; i < n_q; ++i) {
    const unsigned ka = hs.prefix[2 * i], kb = hs.prefix[2 * i + 1];
    const unsigned ua = (ka & 0x80000000u) ? (ka & 0x7fffffffu) : ~ka, ub = (kb & 0x80000000u) ? (kb & 0x7fffffffu) : ~kb;
    float a, b;
    memcpy(&a, &ua, 4);
    memcpy(&b, &ub, 4);
    out[i] = (double)a + ((double)b - (double)a) * frac[i];
  }
  return 0;
}

}  // namespace b2m
