/* b200mcmc.h -- C ABI of libb200mcmc.so, the B200 (sm_100a) sampling path behind the
 * korentomas/mlx-mcmc Python API.
 *
 * The reference has no FFI layer: its boundary is the Python call convention
 *   MCMC(log_prob).run(...)            mlx_mcmc/inference/mcmc.py:38-189
 *   hmc / nuts / metropolis_hastings   mlx_mcmc/kernels/{hmc.py:7, nuts.py:16, metropolis.py:6}
 * and everything below it (mx.grad, elementwise MLX primitives, mx.random) lives in MLX.
 * This header is what a maintainer would bind (ctypes -- see INTEGRATION.md) to replace
 * that part.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success; nonzero => b2m_last_error() (thread local text).
 *   - nothing throws across the ABI; the library never owns caller memory.
 *   - all `float*`/`double*`/`int*` in the *_args structs are DEVICE pointers allocated by the
 *     caller (torch CUDA tensors); `stream` is a cudaStream_t passed as void*.
 *   - one handle <-> one device; handles are not thread safe.
 *   - chains are independent; chain c of this process has global id chain_offset + c and its
 *     random stream depends only on (seed, global id, iteration, slot) -- never on the grid,
 *     the lanes-per-chain choice or the number of GPUs.
 */
#ifndef B200MCMC_H
#define B200MCMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2M_ABI_VERSION 1
#define B2M_MAX_TREE_DEPTH 12

/* ---- model description (produced by the Python tracer from the user's log_prob) ---- */

/* distribution tags: the reference's six classes (mlx_mcmc/distributions/<name>.py) */
enum {
  B2M_NORMAL = 0,      /* normal.py:49-56      p0=loc  p1=scale               */
  B2M_HALFNORMAL = 1,  /* halfnormal.py:55-63  p0=scale           mask x>=0   */
  B2M_EXPONENTIAL = 2, /* exponential.py:61-71 p0=rate            mask x>=0   */
  B2M_GAMMA = 3,       /* gamma.py:53-88       p0=rate(beta) k0=alpha k1=lgamma(alpha)  mask x>0 */
  B2M_BETA = 4,        /* beta.py:53-91        k0=a k1=b k2=logB(a,b)         mask 0<x<1 */
  B2M_CONSTANT = 5     /* a term with no traced operand: value k0 (Categorical with concrete
                          probs/index, categorical.py:69-93, folds to this)   */
};

/* operand kinds: value[n] of one slot of a term, n = element index inside the term */
enum {
  B2M_OP_CONST = 0,    /* c                                                    */
  B2M_OP_PARAM = 1,    /* theta[a]                                             */
  B2M_OP_DATA = 2,     /* arrays[a][n]                                         */
  B2M_OP_PARAMVEC = 3, /* theta[a + n]            (elementwise over a vector parameter) */
  B2M_OP_LIN = 4,      /* c + sum_{e in [a, a+b)} lin[e].coef * (lin[e].array>=0 ? arrays[.][n] : 1)
                            * (lin[e].param>=0 ? theta[lin[e].param] : 1)      */
  B2M_OP_MATVEC = 5    /* c + sum_d arrays[a][n, d] * theta[b + d]   (X @ beta; arrays[a] is [N, cols]) */
};

typedef struct {
  int32_t kind;
  int32_t a;
  int32_t b;
  float c;
} b2m_operand;

typedef struct {
  int32_t param; /* -1: no parameter factor */
  int32_t array; /* -1: no data factor      */
  float coef;
} b2m_lin_entry;

typedef struct {
  int32_t dist;
  int32_t length; /* elements summed by this term (1 for a scalar term) */
  float weight;   /* the term contributes weight * sum_n log p(x_n | ...) */
  float k0, k1, k2;
  b2m_operand x, p0, p1;
} b2m_term;

typedef struct {
  const float *data; /* device, row-major */
  int64_t rows;
  int64_t cols; /* 1 for a vector */
} b2m_array;

typedef struct b2m_model b2m_model;

/* ---- sampler arguments ---- */

enum {
  B2M_ADAPT_NONE = 0,
  B2M_ADAPT_REFERENCE = 1,      /* HMC: the +-5 % rule of hmc.py:164-170, per chain */
  B2M_ADAPT_DUAL_AVERAGING = 2, /* per chain: every chain is an independent replica of the reference's recurrences */
  B2M_ADAPT_POOLED = 3          /* NUTS, GLM class: ONE dual-averaging state driven by the mean acceptance statistic
                                   over the chains of the call; all chains share the step size (keeps a lock-step
                                   batch at one tree depth).  An extension: the reference has a single chain. */
};
enum { B2M_COMPAT_REFERENCE = 0, B2M_COMPAT_CORRECT = 1 };

/* Replaces hmc()'s warm-up and sampling loops, kernels/hmc.py:155-198, with hmc_step :113-153,
 * leapfrog_step :69-100, hamiltonian :102-111 and the +-5 % rule :164-170 inside one launch. */
typedef struct {
  int64_t n_chains;     /* chains in this call (local)                                       */
  int64_t chain_offset; /* global id of local chain 0                                        */
  int64_t iter_offset;  /* global iteration index of the first iteration of this call        */
  int32_t n_iter;       /* iterations to run                                                 */
  int32_t n_leapfrog;   /* L                                                                 */
  int32_t adapt;        /* B2M_ADAPT_*: applied during this call (warm-up) or not (sampling) */
  int32_t lanes;        /* lanes per chain: 0 = library picks, else 1,2,4,8,16,32            */
  double target_accept;
  uint64_t seed;
  float *theta;        /* [n_chains, D] in/out                                                */
  double *step_size;   /* [n_chains] in/out                                                   */
  int64_t *n_accept;   /* [n_chains] in/out (cumulative accepted, drives the +-5 % rule)      */
  int64_t *n_total;    /* [n_chains] in/out                                                   */
  double *da_state;    /* [n_chains, 3] (H_bar, log eps_bar, mu) for dual averaging, or NULL  */
  float *draws;        /* [n_iter, n_chains, D] or NULL (warm-up)                             */
  const float *inj_normal;  /* [n_iter, n_chains, D] momentum draws to use instead of Philox, or NULL */
  const float *inj_uniform; /* [n_iter, n_chains]    accept uniforms, or NULL                 */
  float *trace_energy;      /* [n_iter, n_chains, 2] (H_init, H_prop) or NULL                 */
  uint8_t *trace_accept;    /* [n_iter, n_chains] or NULL                                     */
} b2m_hmc_args;

/* Replaces metropolis_hastings(), kernels/metropolis.py:6-101 (proposal :66-74, accept :77-88). */
typedef struct {
  int64_t n_chains;
  int64_t chain_offset;
  int64_t iter_offset;
  int32_t n_iter;
  int32_t lanes;
  float proposal_scale;
  int32_t _pad;
  uint64_t seed;
  float *theta;       /* [n_chains, D] in/out */
  float *logp;        /* [n_chains] in/out: cached current log-prob (metropolis.py:55); NaN => recompute */
  int64_t *n_accept;  /* [n_chains] in/out */
  float *draws;       /* [n_iter, n_chains, D] or NULL */
  const float *inj_normal;  /* [n_iter, n_chains, D] or NULL */
  const float *inj_uniform; /* [n_iter, n_chains] or NULL */
  uint8_t *trace_accept;    /* [n_iter, n_chains] or NULL */
} b2m_mh_args;

/* Replaces nuts()'s loops, kernels/nuts.py:287-343, with nuts_step :220-285, build_tree :137-218
 * (iterative, one leaf per lock-step), no_u_turn :119-135 and dual averaging :62-68,298-310. */
typedef struct {
  int64_t n_chains;
  int64_t chain_offset;
  int64_t iter_offset;
  int32_t n_iter;
  int32_t max_tree_depth; /* <= B2M_MAX_TREE_DEPTH */
  int32_t adapt;          /* B2M_ADAPT_NONE, B2M_ADAPT_DUAL_AVERAGING (the reference's recurrences) or B2M_ADAPT_POOLED */
  int32_t compat;         /* B2M_COMPAT_REFERENCE: float32 slice underflow + NaN => alpha 1 (nuts.py:236-237,173)
                             B2M_COMPAT_CORRECT:   log-space slice, NaN => divergent, alpha 0 */
  int32_t lanes;
  float step_size_jitter; /* 0 = off (the reference).  > 0: every chain and iteration integrates with eps (1 + jitter (2u - 1)),
                             u from the spare word of Philox slot 0; adaptation keeps tracking the un-jittered eps.  Breaks the
                             resonance of the position-based U-turn test when a power of two steps is close to a whole
                             number of oscillation periods. */
  double target_accept;
  uint64_t seed;
  float *theta;        /* [n_chains, D] in/out */
  double *step_size;   /* [n_chains] in/out: eps used by the next iteration */
  double *da_state;    /* [n_chains, 3]: (H_bar, eps_bar, mu as float) in/out */
  int64_t *n_accept;   /* [n_chains] in/out: iterations with mean alpha > 0.5 (nuts.py:341) */
  int64_t *n_leaves;   /* [n_chains] in/out: leapfrog steps actually taken */
  int64_t *n_diverge;  /* [n_chains] in/out */
  float *draws;        /* [n_iter, n_chains, D] or NULL */
  int32_t *depths;     /* [n_iter, n_chains] or NULL */
  float *alphas;       /* [n_iter, n_chains] mean alpha, or NULL */
  /* draw injection (parity check 2); all NULL => Philox */
  const float *inj_normal; /* [n_iter, n_chains, D] */
  const float *inj_slice;  /* [n_iter, n_chains] */
  const float *inj_dir;    /* [n_iter, n_chains, max_tree_depth] */
  const float *inj_take;   /* [n_iter, n_chains, max_tree_depth] */
  const float *inj_merge;  /* [n_iter, n_chains, max_tree_depth, 2^max_tree_depth - 1] post-order merge slots */
  int32_t *trace_doubling; /* [n_iter, n_chains, max_tree_depth, 6] (v, n_sub, s_sub, took, s, n) or NULL */
  float *trace_energy;     /* [n_iter, n_chains] H0 or NULL */
} b2m_nuts_args;

/* ---- entry points ---- */

const char *b2m_last_error(void);
int b2m_abi_version(void);
/* sizeof of the structs above as compiled, so a binding can verify its mirror */
int b2m_struct_sizes(int32_t *out6); /* term, operand, lin_entry, hmc_args, mh_args, nuts_args */

/* Build a model from the traced term table.  Replaces the user log_prob + Distribution.log_prob
 * bodies (distributions/<name>.py) as differentiated by grad_log_prob, kernels/hmc.py:53-67.
 * `arrays[i].data` must stay valid for the life of the model. */
int b2m_model_create(const b2m_term *terms, int32_t n_terms, const b2m_lin_entry *lin, int32_t n_lin,
                     const b2m_array *arrays, int32_t n_arrays, int32_t D, b2m_model **out);
void b2m_model_destroy(b2m_model *m);
int b2m_model_dim(const b2m_model *m);
/* 0 = pointwise class (persistent register-resident kernels), 1 = GLM class (X @ beta, GEMM kernels) */
int b2m_model_class(const b2m_model *m);
/* GLM class: arithmetic of the two contractions -- 0 fp32 FMA tiles, 1 tcgen05 3xTF32, 2 tcgen05 3xFP16 (operands scaled
 * by powers of two into fp16's range, chosen automatically unless the data's dynamic range is too wide); -1 otherwise.
 * The environment variable B2M_GLM_PATH = simt | tc | tc16 forces one. */
int b2m_model_glm_path(const b2m_model *m);

/* log p(theta_c) and d/dtheta for C chains.  Replaces mx.grad(log_prob_flat)(*params) plus the
 * separate value call in hamiltonian() (kernels/hmc.py:53-67,102-111).  grad may be NULL. */
int b2m_logp_grad(b2m_model *m, const float *theta, int64_t n_chains, float *logp, float *grad,
                  int32_t lanes, void *stream);

int b2m_hmc_run(b2m_model *m, const b2m_hmc_args *a, void *stream);
int b2m_mh_run(b2m_model *m, const b2m_mh_args *a, void *stream);
int b2m_nuts_run(b2m_model *m, const b2m_nuts_args *a, void *stream);

/* ---- observation sharding (no reference counterpart: the reference is single device) ----
 * Every rank builds the model on its own row shard of (X, y) and holds every chain.  After b2m_model_set_comm each
 * value+gradient of a GLM-class model sums the [C, D] gradient partial and the [C] sum of squared residuals over
 * the ranks (one NCCL all-reduce on the caller's stream), so all ranks compute bit-identical results.
 * NCCL is bound at run time; without it these calls return an error and everything else keeps working. */
typedef struct b2m_comm b2m_comm;
int b2m_comm_unique_id(uint8_t *out128);                       /* rank 0 creates the 128-byte id; caller broadcasts it */
int b2m_comm_init(const uint8_t *id128, int32_t n_ranks, int32_t rank, b2m_comm **out);   /* collective */
void b2m_comm_destroy(b2m_comm *c);
int b2m_model_set_comm(b2m_model *m, b2m_comm *c, void *stream); /* collective (sums the shard row counts); c may be NULL */
int b2m_comm_allreduce_f32(b2m_comm *c, float *buf, int64_t n, void *stream);   /* in place, sum */

/* ---- forward sampling on the device ----
 * n draws from one library distribution into out[n] (device).  Replaces Distribution.sample (distributions/<name>.py; the
 * reference's Gamma / Beta fall back to numpy, gamma.py:107-117, beta.py:110-119).  dist = B2M_NORMAL (p0 loc, p1 scale),
 * B2M_HALFNORMAL (p0 scale), B2M_EXPONENTIAL (p0 rate), B2M_GAMMA (p0 alpha, p1 rate), B2M_BETA (p0 a, p1 b) or
 * B2M_SAMPLE_CATEGORICAL (cdf = device array of n_cat cumulative probabilities, last = 1; draws are indices as floats).
 * Output i depends only on (seed, i). */
#define B2M_SAMPLE_CATEGORICAL 6
int b2m_sample(int32_t dist, float p0, float p1, const float *cdf, int32_t n_cat, uint64_t seed, int64_t n, float *out,
               void *stream);

/* ---- diagnostics on the device ----
 * draws is [S, C, D] as written by the samplers.
 * b2m_diag_series: per (chain, parameter) series -> mean [C, D], variance (ddof 0) [C, D], effective sample size by
 *   the reference's estimator (ess_mode 0: examples/06_nuts_comparison.py:22-41, 1: examples/02_hmc_comparison.py:111-128)
 *   and by Geyer's initial positive sequence; the two ESS outputs may be NULL.
 * b2m_diag_params: per parameter -> out[d][5] = pooled mean, pooled std (np.mean / np.std of
 *   mlx_mcmc/inference/mcmc.py:219-220 over all chains), Gelman-Rubin R-hat, ESS summed over chains (both estimators). */
int b2m_diag_series(const float *draws, int64_t S, int64_t C, int64_t D, int32_t ess_mode, float *mean, float *var,
                    float *ess_ref, float *ess_geyer, void *stream);
int b2m_diag_params(const float *mean, const float *var, const float *ess_ref, const float *ess_geyer, int64_t S,
                    int64_t C, int64_t D, double *out, void *stream);

/* Instrumentation: CUDA-event timing of every GEMM launch of the GLM class, on the launching stream.
 * b2m_profile(1) resets and starts recording, b2m_profile(0) stops; b2m_profile_read synchronises the device and
 * returns {K5 total ms, K5 launches, K6 total ms, K6 launches}. */
int b2m_profile(int32_t enable);
int b2m_profile_read(double *out4);

/* number of kernel launches this library has issued since load (bench.py's gpu_launches) */
int64_t b2m_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MCMC_H */
