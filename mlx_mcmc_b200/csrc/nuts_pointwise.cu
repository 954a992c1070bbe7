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
#include "pointwise.cuh"

namespace b2m {

template <int DMAX>
__device__ __forceinline__ void copy(float (&dst)[DMAX], const float (&src)[DMAX]) {
#pragma unroll
  for (int d = 0; d < DMAX; ++d) dst[d] = src[d];
}

// reference order: delta = q_hi - q_lo; sum(delta * p) left to right (nuts.py:125-133)
template <int DMAX>
__device__ __forceinline__ bool keeps_straight(const float (&q_lo)[DMAX], const float (&q_hi)[DMAX],
                                               const float (&p_lo)[DMAX], const float (&p_hi)[DMAX], int D) {
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) {
      const float dq = __fsub_rn(q_hi[d], q_lo[d]);
      a = __fadd_rn(a, __fmul_rn(dq, p_lo[d]));
      b = __fadd_rn(b, __fmul_rn(dq, p_hi[d]));
    }
  return a >= 0.f && b >= 0.f;
}

template <int DMAX>
struct StackEntry {  // a finished LEFT subtree waiting for its sibling
  float first_q[DMAX], first_p[DMAX];
  float cand_q[DMAX], cand_g[DMAX];
  float cand_lp;
  int n;
  int n_alpha;
  double alpha;
};

template <int DMAX, bool COMPACT>
__global__ void __launch_bounds__(64) nuts_kernel(const __grid_constant__ KModel km, b2m_nuts_args A) {
  extern __shared__ __align__(16) unsigned char smem[];
  SModel sm;
  unsigned char *mail = model_to_smem(km, smem, sm);
  const int G = A.lanes;
  Lane L = make_lane(A.n_chains, G, mail, DMAX);
  const int D = sm.D, MD = A.max_tree_depth;
  const int64_t C = A.n_chains, c = L.chain;
  const uint64_t gchain = (uint64_t)(A.chain_offset + c);
  const int n_merge_slots = (1 << MD) - 1;
  const bool ref_compat = A.compat == B2M_COMPAT_REFERENCE;

  float q[DMAX], g[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) q[d] = (d < D) ? A.theta[c * D + d] : 0.f;
  double eps = A.step_size[c];
  double h_bar = A.da_state[c * 3 + 0], eps_bar = A.da_state[c * 3 + 1];
  const float mu = (float)A.da_state[c * 3 + 2];
  int64_t n_acc = A.n_accept[c], n_leaves = A.n_leaves[c], n_div = A.n_diverge[c];

  float lp = evaluate<DMAX, COMPACT, true>(km, sm, L, q, g);

  StackEntry<DMAX> stack[B2M_MAX_TREE_DEPTH];
  float im[DMAX], sqm[DMAX];   // diagonal mass matrix: all ones unless A.inv_mass is given (x * 1.0f is exact)
  load_mass<DMAX>(A.inv_mass, D, im, sqm);

  for (int it = 0; it < A.n_iter; ++it) {
    const uint32_t giter = (uint32_t)(A.iter_offset + it);
    const size_t row = (size_t)it * C + c;
    const uint4 w0 = Philox::draw(A.seed, gchain, giter, 0u);

    float p0[DMAX];
    draw_normals<DMAX>(p0, D, A.inj_normal ? A.inj_normal + row * D : nullptr, A.seed, gchain, giter, w0);
#pragma unroll
    for (int d = 0; d < DMAX; ++d) p0[d] = __fmul_rn(p0[d], sqm[d]);   // p ~ N(0, M)
    const float h0 = __fadd_rn(-lp, kinetic_m<DMAX>(p0, im, D));
    // step size of this iteration: jittered when asked for (extension), else exactly eps
    const double eps_it = A.step_size_jitter > 0.f ? eps * (1.0 + (double)A.step_size_jitter * (2.0 * (double)u01(w0.w) - 1.0)) : eps;
    const float us = A.inj_slice ? A.inj_slice[row] : u01(w0.z);
    const double log_u64 = (double)(-h0) + (double)logf(us);
    // reference: u = exp(float32(log_u)); later float(log(u))
    const float log_slice = ref_compat ? logf(expf((float)log_u64)) : (float)log_u64;
    if (L.writer && A.trace_energy) A.trace_energy[row] = h0;

    float q_lo[DMAX], p_lo[DMAX], g_lo[DMAX], q_hi[DMAX], p_hi[DMAX], g_hi[DMAX];
    copy<DMAX>(q_lo, q); copy<DMAX>(q_hi, q);
    copy<DMAX>(p_lo, p0); copy<DMAX>(p_hi, p0);
    copy<DMAX>(g_lo, g); copy<DMAX>(g_hi, g);
    float cq[DMAX], cg[DMAX], clp = lp;  // the transition's candidate (with cached lp / grad)
    copy<DMAX>(cq, q); copy<DMAX>(cg, g);

    int j = 0, n = 1;
    bool s = true;
    double alpha_sum = 0.0;
    int alpha_cnt = 0;

    while (s && j < MD) {
      uint4 wj = make_uint4(0, 0, 0, 0);
      if (!A.inj_dir || !A.inj_take) wj = Philox::draw(A.seed, gchain, giter, SLOT_NUTS_DOUBLING + j);
      const float ud = A.inj_dir ? A.inj_dir[row * MD + j] : u01(wj.x);
      const int v = ud < 0.5f ? 1 : -1;
      const float feps = (float)((double)v * eps_it), half_eps = (float)(0.5 * ((double)v * eps_it));

      // frontier = the edge we extend from
      float fq[DMAX], fp[DMAX], fg[DMAX];
      if (v == 1) { copy<DMAX>(fq, q_hi); copy<DMAX>(fp, p_hi); copy<DMAX>(fg, g_hi); }
      else        { copy<DMAX>(fq, q_lo); copy<DMAX>(fp, p_lo); copy<DMAX>(fg, g_lo); }

      // the subtree being assembled at the current leaf
      float sub_first_q[DMAX], sub_first_p[DMAX], sub_cq[DMAX], sub_cg[DMAX];
      float sub_clp = 0.f;
      int sub_n = 0, sub_na = 0;
      bool sub_s = true;
      double sub_alpha = 0.0;

      const int n_leaf = 1 << j;
      for (int i = 0; i < n_leaf; ++i) {
        // ---- leaf: one leapfrog step (two half kicks) + energy
#pragma unroll
        for (int d = 0; d < DMAX; ++d) {
          fp[d] = __fadd_rn(fp[d], __fmul_rn(half_eps, fg[d]));
          fq[d] = __fadd_rn(fq[d], __fmul_rn(feps, __fmul_rn(im[d], fp[d])));
        }
        const float flp = evaluate<DMAX, COMPACT, true>(km, sm, L, fq, fg);
#pragma unroll
        for (int d = 0; d < DMAX; ++d) fp[d] = __fadd_rn(fp[d], __fmul_rn(half_eps, fg[d]));
        ++n_leaves;
        const float h1 = __fadd_rn(-flp, kinetic_m<DMAX>(fp, im, D));
        const int n1 = (log_slice <= -h1) ? 1 : 0;
        const bool s1 = log_slice < __fsub_rn(1000.0f, h1);
        float a1 = expf(__fadd_rn(-h1, h0));
        if (a1 != a1) a1 = ref_compat ? 1.0f : 0.0f;
        a1 = fminf(a1, 1.0f);
        if (!s1) ++n_div;

        copy<DMAX>(sub_first_q, fq); copy<DMAX>(sub_first_p, fp);
        copy<DMAX>(sub_cq, fq); copy<DMAX>(sub_cg, fg);
        sub_clp = flp; sub_n = n1; sub_s = s1; sub_alpha = (double)a1; sub_na = 1;

        // ---- close every subtree that ends at this leaf
        int k = 0;
        bool finished = false;
        while (true) {
          if (k == j) { finished = true; break; }
          if ((i >> k) & 1) {
            StackEntry<DMAX> &Lf = stack[k];
            const int m = i - __popc(i) + k;  // post-order merge slot
            float um;
            if (A.inj_merge) {
              um = A.inj_merge[(row * MD + j) * n_merge_slots + m];
            } else {
              const uint4 wm = Philox::draw(A.seed, gchain, giter, SLOT_NUTS_MERGE + 1024u * j + (m >> 2));
              const uint32_t word = (m & 3) == 0 ? wm.x : (m & 3) == 1 ? wm.y : (m & 3) == 2 ? wm.z : wm.w;
              um = u01(word);
            }
            const int n_tot = Lf.n + sub_n;
            const double ratio = (double)sub_n / fmax((double)n_tot, 1.0);
            if (!((double)um < ratio)) {  // keep the left half's candidate
              copy<DMAX>(sub_cq, Lf.cand_q); copy<DMAX>(sub_cg, Lf.cand_g);
              sub_clp = Lf.cand_lp;
            }
            bool straight;
            if (v == 1) straight = keeps_straight<DMAX>(Lf.first_q, fq, Lf.first_p, fp, D);
            else        straight = keeps_straight<DMAX>(fq, Lf.first_q, fp, Lf.first_p, D);
            sub_s = sub_s && straight;
            copy<DMAX>(sub_first_q, Lf.first_q); copy<DMAX>(sub_first_p, Lf.first_p);
            sub_n = n_tot;
            sub_alpha = Lf.alpha + sub_alpha;
            sub_na = Lf.n_alpha + sub_na;
            ++k;
          } else {
            if (sub_s) {  // valid left half: park it and go build its sibling
              StackEntry<DMAX> &Lf = stack[k];
              copy<DMAX>(Lf.first_q, sub_first_q); copy<DMAX>(Lf.first_p, sub_first_p);
              copy<DMAX>(Lf.cand_q, sub_cq); copy<DMAX>(Lf.cand_g, sub_cg);
              Lf.cand_lp = sub_clp; Lf.n = sub_n; Lf.n_alpha = sub_na; Lf.alpha = sub_alpha;
              break;
            }
            ++k;  // invalid left half: the parent returns it unchanged, no sibling, no draw
          }
        }
        if (finished) break;
      }

      // ---- top level of nuts_step (nuts.py:257-280)
      if (v == 1) { copy<DMAX>(q_hi, fq); copy<DMAX>(p_hi, fp); copy<DMAX>(g_hi, fg); }
      else        { copy<DMAX>(q_lo, fq); copy<DMAX>(p_lo, fp); copy<DMAX>(g_lo, fg); }
      int took = 0;
      if (sub_s) {
        const float ut = A.inj_take ? A.inj_take[row * MD + j] : u01(wj.y);
        const double pr = fmin(1.0, (double)sub_n / fmax((double)n, 1.0));
        if ((double)ut < pr) {
          copy<DMAX>(cq, sub_cq); copy<DMAX>(cg, sub_cg);
          clp = sub_clp;
          took = 1;
        }
      }
      n += sub_n;
      s = sub_s && keeps_straight<DMAX>(q_lo, q_hi, p_lo, p_hi, D);
      alpha_sum += sub_alpha;
      alpha_cnt += sub_na;
      if (L.writer && A.trace_doubling) {
        int32_t *tr = A.trace_doubling + ((size_t)row * MD + j) * 6;
        tr[0] = v; tr[1] = sub_n; tr[2] = sub_s ? 1 : 0; tr[3] = took; tr[4] = s ? 1 : 0; tr[5] = n;
      }
      ++j;
    }

    copy<DMAX>(q, cq); copy<DMAX>(g, cg);
    lp = clp;
    const double mean_alpha = alpha_sum / fmax((double)alpha_cnt, 1.0);
    n_acc += mean_alpha > 0.5 ? 1 : 0;

    if (A.adapt == B2M_ADAPT_DUAL_AVERAGING) {
      // nuts.py:298-310 with its float32 / float64 split
      const double m = (double)((int64_t)giter - A.adapt_origin);
      const double eta = 1.0 / (m + 10.0);
      h_bar = (1.0 - eta) * h_bar + eta * (A.target_accept - mean_alpha);
      float log_eps = __fsub_rn(mu, (float)((sqrt(m + 1.0) / 0.05) * h_bar));
      log_eps = fmaxf(fminf(log_eps, 10.0f), -10.0f);
      eps = (double)expf(log_eps);
      const double wgt = pow(m + 1.0, -0.75);
      eps_bar = (double)expf((float)(wgt * log(eps) + (1.0 - wgt) * log(eps_bar)));
    }

    if (L.writer) {
      if (A.draws) store_draw<DMAX>(A.draws + row * D, q, D, km, A.draws_unconstrained != 0);
      if (A.depths) A.depths[row] = j;
      if (A.alphas) A.alphas[row] = (float)mean_alpha;
    }
  }

  if (L.writer) {
    store_vec<DMAX>(A.theta + c * D, q, D);
    A.step_size[c] = eps;
    A.da_state[c * 3 + 0] = h_bar;
    A.da_state[c * 3 + 1] = eps_bar;
    A.n_accept[c] = n_acc;
    A.n_leaves[c] = n_leaves;
    A.n_diverge[c] = n_div;
  }
}

int launch_nuts(const KModel &km, b2m_nuts_args a, cudaStream_t st) {
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
