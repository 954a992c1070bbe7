"""Small end-to-end pass over every kernel family, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mlx_mcmc_b200 as B
import mlx_mcmc_b200.core as mx
from mlx_mcmc_b200 import workloads as W
from mlx_mcmc_b200.diagnostics import device_summary
from mlx_mcmc_b200.engine import compile_model

for name in ("c1_normal", "c2_event_rate", "c5_ab_test", "t_vector_normal"):
    fn, init, _ = W.ALL_SMALL[name](B.ns)
    m = compile_model(fn, init, cache=False)
    m.logp_grad(m.pack(init, 70))
    if name != "t_vector_normal":
        B.hmc(fn, init, num_samples=5, num_warmup=5, num_chains=70, key=mx.random.key(1))
    B.nuts(fn, init, num_samples=4, num_warmup=4, num_chains=70, max_tree_depth=4, key=mx.random.key(2))
    B.metropolis_hastings(fn, init, num_samples=6, num_chains=70, random_seed=3)
for path in ("simt", "tc", "tc16"):
    fn, init, meta = W.regression(B.ns, 700, 70, seed=1)
    m = compile_model(fn, init, cache=False, glm_path=path)
    th = torch.from_numpy((meta.beta_true[None] + 0.1 * np.random.default_rng(0).standard_normal((300, 70))).astype(np.float32)).cuda()
    m.logp_grad(th)
    s, _ = B.nuts(fn, init, num_samples=3, num_warmup=3, num_chains=300, max_tree_depth=3, compat="correct",
                  step_size_adaptation="pooled", key=mx.random.key(4), model=m, return_torch=True)
    B.hmc(fn, init, num_samples=2, num_warmup=2, num_leapfrog_steps=3, num_chains=130, key=mx.random.key(5), model=m)
    B.metropolis_hastings(fn, init, num_samples=3, num_chains=130, random_seed=6, model=m)
    device_summary(s["beta"].permute(1, 0, 2).contiguous())
torch.cuda.synchronize()
print("sanitize_smoke done")
