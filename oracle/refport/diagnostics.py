"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's diagnostics for SURVEY.md 8(f) rows 1-2.

The reference ships no diagnostics module (README.md:129-136 only draws one); what exists is
  * `compute_ess` in examples/06_nuts_comparison.py:22-41 and its variant in examples/02_hmc_comparison.py:111-128,
  * `MCMC.summary` in mlx_mcmc/inference/mcmc.py:191-225.
Pinned by oracle/make_golden_diag.py, which executes those reference function bodies on seeded series and asserts
this restatement reproduces them exactly (tests/golden/diagnostics.json).
"""
import numpy as np


def compute_ess(samples):
    """examples/06_nuts_comparison.py:22-41 -- n / (1 + 2 sum rho_k) over lags 1..min(n//2,100)-1, stopping after the
    first rho_k < 0.05 (that lag included); n when the variance is zero."""
    n = len(samples)
    mean = np.mean(samples)
    var = np.var(samples)
    if var == 0:
        return n
    rho = []
    for lag in range(1, min(n // 2, 100)):
        r = np.mean((samples[:-lag] - mean) * (samples[lag:] - mean)) / var
        rho.append(r)
        if r < 0.05:
            break
    return n / (1 + 2 * np.sum(rho))


def compute_ess_example02(samples):
    """examples/02_hmc_comparison.py:111-128 -- same sum, but the stop test only applies from the second lag on."""
    n = len(samples)
    mean = np.mean(samples)
    c0 = np.mean((samples - mean) ** 2)
    rho = []
    for lag in range(1, min(n // 2, 100)):
        c = np.mean((samples[:-lag] - mean) * (samples[lag:] - mean))
        rho.append(c / c0)
        if len(rho) > 1 and rho[-1] < 0.05:
            break
    return n / (1 + 2 * np.sum(rho))


def summary(samples: dict, credible_interval=0.95):
    """mlx_mcmc/inference/mcmc.py:212-225 -- mean / std (ddof 0) / median / lower / upper percentile per parameter."""
    alpha = 1 - credible_interval
    lo, hi = 100 * alpha / 2, 100 * (1 - alpha / 2)
    out = {}
    for name, x in samples.items():
        out[name] = {"mean": float(np.mean(x)), "std": float(np.std(x)), "median": float(np.median(x)),
                     f"{lo:.1f}%": float(np.percentile(x, lo)), f"{hi:.1f}%": float(np.percentile(x, hi))}
    return out


def rhat(chains):
    """Gelman-Rubin potential scale reduction for `chains` [C, S] (not in the reference: README.md:212-216 lists it
    as planned).  W = mean within-chain variance (ddof 1), B/S = variance of the chain means (ddof 1)."""
    x = np.asarray(chains, dtype=np.float64)
    c, s = x.shape
    w = np.mean(np.var(x, axis=1, ddof=1))
    b_over_s = np.var(np.mean(x, axis=1), ddof=1) if c > 1 else 0.0
    return float(np.sqrt(((s - 1) / s * w + b_over_s) / w))
