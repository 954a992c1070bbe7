// Pointwise-class kernels: many independent chains in lock-step inside persistent kernels.
// One launch covers `n_iter` whole iterations (an HMC trajectory of L leapfrog steps + accept +
// step-size adaptation; a Metropolis step; a NUTS transition) for every local chain.
//
// Mapping: a chain is served by G lanes (G = 1 for >= 32K chains: thread per chain, observations
// broadcast from shared memory; G up to 32 for few chains x many observations: lanes stride the
// observation vector and combine with xor shuffles).  Chain state (q, p, grad) lives in registers
// (template DMAX); a shared-memory mailbox per lane carries theta / grad through the term loop.
#include "pointwise_kernels.cuh"

namespace b2m {

// ---------------------------------------------------------------- generic kernels (bodies: pointwise_kernels.cuh)
template <int DMAX, bool COMPACT>
__global__ void __launch_bounds__(128) logp_grad_kernel(const __grid_constant__ KModel km, const float *__restrict__ theta,
                                                         int64_t C, float *__restrict__ logp, float *__restrict__ grad,
                                                         int G) {
  logp_grad_body<DMAX, COMPACT>(km, theta, C, logp, grad, G);
}

template <int DMAX, bool COMPACT>
__global__ void __launch_bounds__(128) hmc_kernel(const __grid_constant__ KModel km, const __grid_constant__ b2m_hmc_args A) {
  hmc_body<DMAX, COMPACT>(km, A);
}

template <int DMAX, bool COMPACT>
__global__ void __launch_bounds__(128) mh_kernel(const __grid_constant__ KModel km, const __grid_constant__ b2m_mh_args A) {
  mh_body<DMAX, COMPACT>(km, A);
}

// ---------------------------------------------------------------- host-side launchers
int pick_lanes(const KModel &km, int64_t n_chains, int requested) {
  if (requested > 0) return requested;
  if (n_chains >= 32768) return 1;
  int g = 1;
  // enough lanes to put ~64K threads in flight, never more lanes than the longest term has elements
  while (g < 32 && (int64_t)n_chains * g < 65536 && g * 2 <= km.max_len) g <<= 1;
  return g;
}

// specialised module attached: same geometry as a compact model (theta / gradient in registers, no mailbox)
static KModel jit_view(const KModel &km) {
  KModel k = km;
  k.compact = 1;
  return k;
}

int launch_logp_grad(const KModel &km, const float *theta, int64_t C, float *logp, float *grad, int lanes,
                     cudaStream_t st, JitModule *jit) {
  if (jit) {
    KModel k = jit_view(km);
    int G = pick_lanes(k, C, lanes);
    Geometry ge = geometry(k, C, G, jit_dmax(jit));
    void *params[] = {&k, &theta, &C, &logp, &grad, &G};
    return jit_launch(jit, JIT_LOGP_GRAD, ge.grid, ge.block, ge.smem, st, params);
  }
  const int dmax = pick_dmax(km.D);
  const int G = pick_lanes(km, C, lanes);
  Geometry ge = geometry(km, C, G, dmax ? dmax : 2);
  B2M_DISPATCH_DMAX(dmax, km.compact, {
    if (int rc = prep(logp_grad_kernel<DM, CP>, ge.smem)) return rc;
    logp_grad_kernel<DM, CP><<<ge.grid, ge.block, ge.smem, st>>>(km, theta, C, logp, grad, G);
  });
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_hmc(const KModel &km, b2m_hmc_args a, cudaStream_t st, JitModule *jit) {
  if (jit) {
    KModel k = jit_view(km);
    a.lanes = pick_lanes(k, a.n_chains, a.lanes);
    Geometry ge = geometry(k, a.n_chains, a.lanes, jit_dmax(jit));
    void *params[] = {&k, &a};
    return jit_launch(jit, JIT_HMC, ge.grid, ge.block, ge.smem, st, params);
  }
  const int dmax = pick_dmax(km.D);
  a.lanes = pick_lanes(km, a.n_chains, a.lanes);
  Geometry ge = geometry(km, a.n_chains, a.lanes, dmax ? dmax : 2);
  B2M_DISPATCH_DMAX(dmax, km.compact, {
    if (int rc = prep(hmc_kernel<DM, CP>, ge.smem)) return rc;
    hmc_kernel<DM, CP><<<ge.grid, ge.block, ge.smem, st>>>(km, a);
  });
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_mh(const KModel &km, b2m_mh_args a, cudaStream_t st, JitModule *jit) {
  if (jit) {
    KModel k = jit_view(km);
    a.lanes = pick_lanes(k, a.n_chains, a.lanes);
    Geometry ge = geometry(k, a.n_chains, a.lanes, jit_dmax(jit));
    void *params[] = {&k, &a};
    return jit_launch(jit, JIT_MH, ge.grid, ge.block, ge.smem, st, params);
  }
  const int dmax = pick_dmax(km.D);
  a.lanes = pick_lanes(km, a.n_chains, a.lanes);
  Geometry ge = geometry(km, a.n_chains, a.lanes, dmax ? dmax : 2);
  B2M_DISPATCH_DMAX(dmax, km.compact, {
    if (int rc = prep(mh_kernel<DM, CP>, ge.smem)) return rc;
    mh_kernel<DM, CP><<<ge.grid, ge.block, ge.smem, st>>>(km, a);
  });
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2m
