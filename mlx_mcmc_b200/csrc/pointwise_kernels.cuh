// Bodies of the pointwise-class kernels (K1 log p / gradient, K2 HMC, K3 Metropolis) as device functions, so that the
// same source serves the generic kernels of pointwise.cu (term table interpreted from the kernel parameters / shared
// memory) and the per-model kernels that mlx_mcmc_b200/jit.py generates and compiles with NVRTC (term table baked in as
// literals: B2M_JIT, model.cuh).
#pragma once
#include "pointwise.cuh"

namespace b2m {

// ---------------------------------------------------------------- K1: log p and gradient
template <int DMAX, bool COMPACT>
__device__ __forceinline__ void logp_grad_body(const KModel &km, const float *__restrict__ theta, int64_t C,
                                               float *__restrict__ logp, float *__restrict__ grad, int G) {
  extern __shared__ __align__(16) unsigned char smem[];
  SModel sm;
  unsigned char *mail = model_to_smem(km, smem, sm);
  Lane L = make_lane(C, G, mail, DMAX);
  const int D = sm.D;
  float q[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) q[d] = (d < D) ? theta[L.chain * D + d] : 0.f;
  float gq[DMAX];
  float lp;
  if (grad)
    lp = evaluate<DMAX, COMPACT, true>(km, sm, L, q, gq);
  else
    lp = evaluate<DMAX, COMPACT, false>(km, sm, L, q, gq);
  if (L.writer) {
    logp[L.chain] = lp;
    if (grad) store_vec<DMAX>(grad + L.chain * D, gq, D);
  }
}

// ---------------------------------------------------------------- K2: HMC
// hmc_step (hmc.py:113-153): momentum ~ N(0,I); H_init; L x leapfrog_step (:69-100, two half kicks,
// separate multiply and add as MLX evaluates them); H_prop; accept iff log U < -(H_prop - H_init).
// The gradient at the trajectory start is the cached gradient of the current state (the reference
// recomputes it: same number).  Warm-up rule (:164-170): for i > 10, eps *= 0.95 if the cumulative
// acceptance rate is below target else 1.05, per chain, in float64 like the python float it replaces.
template <int DMAX, bool COMPACT>
__device__ __forceinline__ void hmc_body(const KModel &km, const b2m_hmc_args &A) {
  extern __shared__ __align__(16) unsigned char smem[];
  SModel sm;
  unsigned char *mail = model_to_smem(km, smem, sm);
  const int G = A.lanes;
  Lane L = make_lane(A.n_chains, G, mail, DMAX);
  const int D = sm.D;
  const int64_t C = A.n_chains, c = L.chain;
  const uint64_t gchain = (uint64_t)(A.chain_offset + c);

  float q[DMAX], g[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) q[d] = (d < D) ? A.theta[c * D + d] : 0.f;
  double eps = A.step_size[c];
  int64_t n_acc = A.n_accept[c], n_tot = A.n_total[c];
  double h_bar = 0.0, log_eps_bar = 0.0, da_mu = 0.0;
  if (A.adapt == B2M_ADAPT_DUAL_AVERAGING) {
    h_bar = A.da_state[c * 3 + 0];
    log_eps_bar = A.da_state[c * 3 + 1];
    da_mu = A.da_state[c * 3 + 2];
  }

  float lp = evaluate<DMAX, COMPACT, true>(km, sm, L, q, g);
  float im[DMAX], sqm[DMAX];   // diagonal mass matrix: all ones unless A.inv_mass is given (x * 1.0f is exact)
  load_mass<DMAX>(A.inv_mass, D, im, sqm);

  for (int it = 0; it < A.n_iter; ++it) {
    const uint32_t giter = (uint32_t)(A.iter_offset + it);
    const size_t row = (size_t)it * C + c;
    uint4 w0 = make_uint4(0, 0, 0, 0);
    if (!A.inj_normal || !A.inj_uniform) w0 = Philox::draw(A.seed, gchain, giter, 0u);

    float p[DMAX];
    draw_normals<DMAX>(p, D, A.inj_normal ? A.inj_normal + row * D : nullptr, A.seed, gchain, giter, w0);
#pragma unroll
    for (int d = 0; d < DMAX; ++d) p[d] = __fmul_rn(p[d], sqm[d]);   // p ~ N(0, M)
    const float h_init = __fadd_rn(-lp, kinetic_m<DMAX>(p, im, D));

    float qn[DMAX], gn[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; ++d) { qn[d] = q[d]; gn[d] = g[d]; }
    float lpn = lp;
    const float half_eps = (float)(0.5 * eps), feps = (float)eps;
    for (int l = 0; l < A.n_leapfrog; ++l) {
#pragma unroll
      for (int d = 0; d < DMAX; ++d) {
        p[d] = __fadd_rn(p[d], __fmul_rn(half_eps, gn[d]));
        qn[d] = __fadd_rn(qn[d], __fmul_rn(feps, __fmul_rn(im[d], p[d])));
      }
      lpn = evaluate<DMAX, COMPACT, true>(km, sm, L, qn, gn);
#pragma unroll
      for (int d = 0; d < DMAX; ++d) p[d] = __fadd_rn(p[d], __fmul_rn(half_eps, gn[d]));
    }
    const float h_prop = __fadd_rn(-lpn, kinetic_m<DMAX>(p, im, D));
    const float u = A.inj_uniform ? A.inj_uniform[row] : u01(w0.z);
    const float log_ratio = -(__fsub_rn(h_prop, h_init));
    const bool accept = logf(u) < log_ratio;  // NaN => false => reject
    if (accept) {
#pragma unroll
      for (int d = 0; d < DMAX; ++d) { q[d] = qn[d]; g[d] = gn[d]; }
      lp = lpn;
      ++n_acc;
    }
    ++n_tot;

    if (A.adapt == B2M_ADAPT_REFERENCE) {
      if ((int64_t)giter > 10) eps *= ((double)n_acc / (double)n_tot < A.target_accept) ? 0.95 : 1.05;
    } else if (A.adapt == B2M_ADAPT_DUAL_AVERAGING) {
      // Hoffman & Gelman Alg. 5 recurrences with the constants the reference uses in nuts.py:62-68
      // a divergent trajectory (NaN / -inf energy) counts as acceptance probability 0
      float a = (log_ratio == log_ratio) ? expf(fminf(log_ratio, 0.f)) : 0.f;
      const double m = (double)((int64_t)giter - A.adapt_origin) + 1.0, eta = 1.0 / (m + 10.0);
      h_bar = (1.0 - eta) * h_bar + eta * (A.target_accept - (double)a);
      double log_eps = da_mu - sqrt(m) / 0.05 * h_bar;
      log_eps = fmin(fmax(log_eps, -10.0), 10.0);
      const double wt = pow(m, -0.75);
      log_eps_bar = wt * log_eps + (1.0 - wt) * log_eps_bar;
      eps = exp(log_eps);
    }

    if (L.writer) {
      if (A.draws)
        store_draw<DMAX>(A.draws + row * D, q, D, km, A.draws_unconstrained != 0);
      if (A.trace_energy) { A.trace_energy[row * 2] = h_init; A.trace_energy[row * 2 + 1] = h_prop; }
      if (A.trace_accept) A.trace_accept[row] = accept ? 1 : 0;
    }
  }

  if (L.writer) {
    store_vec<DMAX>(A.theta + c * D, q, D);
    A.step_size[c] = eps;
    A.n_accept[c] = n_acc;
    A.n_total[c] = n_tot;
    if (A.adapt == B2M_ADAPT_DUAL_AVERAGING) {
      A.da_state[c * 3 + 0] = h_bar;
      A.da_state[c * 3 + 1] = log_eps_bar;
    }
  }
}

// ---------------------------------------------------------------- K3: random-walk Metropolis
// metropolis.py:64-92: theta' = theta + scale * N(0,I) (multiply, then add); accept iff
// log U < lp' - lp with the current log-prob cached; NaN => reject.
template <int DMAX, bool COMPACT>
__device__ __forceinline__ void mh_body(const KModel &km, const b2m_mh_args &A) {
  extern __shared__ __align__(16) unsigned char smem[];
  SModel sm;
  unsigned char *mail = model_to_smem(km, smem, sm);
  const int G = A.lanes;
  Lane L = make_lane(A.n_chains, G, mail, DMAX);
  const int D = sm.D;
  const int64_t C = A.n_chains, c = L.chain;
  const uint64_t gchain = (uint64_t)(A.chain_offset + c);

  float q[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) q[d] = (d < D) ? A.theta[c * D + d] : 0.f;
  float lp = A.logp[c];
  float gdummy[DMAX];
  if (lp != lp) lp = evaluate<DMAX, COMPACT, false>(km, sm, L, q, gdummy);
  int64_t n_acc = A.n_accept[c];

  for (int it = 0; it < A.n_iter; ++it) {
    const uint32_t giter = (uint32_t)(A.iter_offset + it);
    const size_t row = (size_t)it * C + c;
    uint4 w0 = make_uint4(0, 0, 0, 0);
    if (!A.inj_normal || !A.inj_uniform) w0 = Philox::draw(A.seed, gchain, giter, 0u);
    float z[DMAX], qn[DMAX];
    draw_normals<DMAX>(z, D, A.inj_normal ? A.inj_normal + row * D : nullptr, A.seed, gchain, giter, w0);
#pragma unroll
    for (int d = 0; d < DMAX; ++d) qn[d] = __fadd_rn(q[d], __fmul_rn(z[d], A.proposal_scale));
    const float lpn = evaluate<DMAX, COMPACT, false>(km, sm, L, qn, gdummy);
    const float u = A.inj_uniform ? A.inj_uniform[row] : u01(w0.z);
    const bool accept = logf(u) < __fsub_rn(lpn, lp);
    if (accept) {
#pragma unroll
      for (int d = 0; d < DMAX; ++d) q[d] = qn[d];
      lp = lpn;
      ++n_acc;
    }
    if (L.writer) {
      if (A.draws)
        store_draw<DMAX>(A.draws + row * D, q, D, km, false);
      if (A.trace_accept) A.trace_accept[row] = accept ? 1 : 0;
    }
  }
  if (L.writer) {
    store_vec<DMAX>(A.theta + c * D, q, D);
    A.logp[c] = lp;
    A.n_accept[c] = n_acc;
  }
}

}  // namespace b2m
