// K4 (pointwise class): NUTS with the reference's recursive build_tree restated as an iterative,
// per-chain loop over leaves.  Chains of a warp run in lock-step; a chain whose
// doubling ended is masked until its warp-mates finish (SIMT divergence does the masking).
//
// Replaces kernels/nuts.py: nuts_step :220-285, build_tree :137-218, no_u_turn :119-135,
// leapfrog_step :89-111, hamiltonian :113-117, dual averaging :62-68,298-310.
//
// Kept from the reference (compat = B2M_COMPAT_REFERENCE):
//   * slice variable round trip u = exp(float32(-H0 + log U)) -- underflows to 0 for H0 >~ 103 so
//     log u = -inf and every leaf is "in the slice" (nuts.py:236-237,166);
//   * alpha = min(1, exp(H0 - H')) with NaN counted as 1.0 (nuts.py:173);
//   * one uniform per internal merge, drawn only when the left half is valid (nuts.py:188,204-205);
//   * n += n' even when the new subtree is invalid (nuts.py:275).
// compat = B2M_COMPAT_CORRECT keeps log u in log space and counts NaN leaves as alpha = 0.
// Every leaf costs one fused value+gradient: the gradient at each tree edge is cached.
#include "nuts_pointwise_kernel.cuh"

namespace b2m {

template <int DMAX, bool COMPACT>
__global__ void __launch_bounds__(64) nuts_kernel(const __grid_constant__ KModel km, const __grid_constant__ b2m_nuts_args A) {
  nuts_body<DMAX, COMPACT>(km, A);
}

int launch_nuts(const KModel &km, b2m_nuts_args a, cudaStream_t st, JitModule *jit) {
  if (jit) {   // specialised module attached: compact geometry (no mailbox)
    KModel k = km;
    k.compact = 1;
    a.lanes = pick_lanes(k, a.n_chains, a.lanes);
    Geometry ge = geometry(k, a.n_chains, a.lanes, jit_dmax(jit));
    void *params[] = {&k, &a};
    return jit_launch(jit, JIT_NUTS, ge.grid, ge.block, ge.smem, st, params);
  }
  const int dmax = pick_dmax(km.D);
  a.lanes = pick_lanes(km, a.n_chains, a.lanes);
  Geometry ge = geometry(km, a.n_chains, a.lanes, dmax ? dmax : 2);
  B2M_DISPATCH_DMAX(dmax, km.compact, {
    if (int rc = prep(nuts_kernel<DM, CP>, ge.smem)) return rc;
    nuts_kernel<DM, CP><<<ge.grid, ge.block, ge.smem, st>>>(km, a);
  });
  ++g_launches;
  B2M_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2m
