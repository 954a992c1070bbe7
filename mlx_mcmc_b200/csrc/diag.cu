// posterior diagnostics on the device, so that draws [S, C, D] (16-40 GB at the large
// configurations) never have to travel to the host to be summarised.
//
// Replaces, per (chain, parameter) series:
//   compute_ess        examples/06_nuts_comparison.py:22-41   (mode 0)  and examples/02_hmc_comparison.py:111-128 (mode 1)
//   np.mean / np.std   mlx_mcmc/inference/mcmc.py:219-221     (pooled over chains in diag_params_kernel)
// and adds what the reference lists as planned (README.md:212-216): Geyer initial-positive-sequence ESS and the
// Gelman-Rubin R-hat across chains.  float32 draws, float64 accumulation.
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

}  // namespace b2m
