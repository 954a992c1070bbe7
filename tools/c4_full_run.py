"""The user-facing flow at the Target configuration: MCMC(log_prob).run(...) from beta = 0, 4096 chains, NUTS with
pooled step-size adaptation; reports wall time, the tree depth per warm-up iteration and device-side diagnostics."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mlx_mcmc_b200 as B
from mlx_mcmc_b200 import workloads as W

n, d, c = int(os.environ.get("N", 100000)), int(os.environ.get("D", 1000)), int(os.environ.get("C", 4096))
warm, samp = int(os.environ.get("WARM", 60)), int(os.environ.get("SAMP", 40))
fn, init, meta = W.regression(B.ns, n, d, seed=0)
m = B.MCMC(fn)
torch.cuda.synchronize()
t0 = time.perf_counter()
m.run(init, num_samples=samp, num_warmup=warm, method="nuts", step_size=float(os.environ.get("EPS", 0.001)), num_chains=c,
      compat="correct", step_size_adaptation="pooled", step_size_jitter=float(os.environ.get("JITTER", 0.0)),
      adapt_mass_matrix=bool(int(os.environ.get("MASS", "0"))), return_torch=True, return_info=True, verbose=False, random_seed=1)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
info = m.info
diag = m.diagnostics()["beta"]
post_m, post_V = W.regression_posterior(meta) if d <= 1000 else (None, None)
out = {"wall_s": wall, "warmup_depth_per_iter_max": info.warmup_depths.max(axis=1).tolist(),
       "warmup_depth_per_iter_mean": [round(float(v), 2) for v in info.warmup_depths.mean(axis=1)],
       "sampling_depth_mean": float(info.depths.mean()), "step_size": float(info.step_size[0]),
       "grad_evals": info.grad_evals, "grad_evals_per_s": info.grad_evals / wall,
       "min_ess_geyer": float(diag["ess_geyer"].min()), "min_ess_per_s": float(diag["ess_geyer"].min()) / wall,
       "max_rhat": float(np.nanmax(diag["rhat"])), "n_diverge": int(info.n_diverge.sum())}
if post_m is not None:
    out["max_abs_mean_err_in_sd"] = float(np.max(np.abs(diag["mean"] - post_m) / np.sqrt(np.diag(post_V))))
    out["sd_ratio_range"] = [float(np.min(diag["std"] / np.sqrt(np.diag(post_V)))), float(np.max(diag["std"] / np.sqrt(np.diag(post_V))))]
print(json.dumps(out))
