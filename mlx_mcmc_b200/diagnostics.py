"""Effective-sample-size estimators used for the min-ESS/s metric.

`compute_ess` is the ad-hoc estimator of the reference's examples (examples/06_nuts_comparison.py:22-41),
reported verbatim; `ess_geyer` is the standard initial-positive-sequence estimator, used whenever the
two disagree in sign (the reference's estimator goes negative on antithetic chains)."""
from __future__ import annotations

import numpy as np


def compute_ess(samples) -> float:
    """n / (1 + 2 sum rho_k), lags 1..min(n/2,100)-1, stopping after the first rho_k < 0.05."""
    x = np.asarray(samples, dtype=np.float64)
    n = len(x)
    mean, var = np.mean(x), np.var(x)
    if var == 0:
        return float(n)
    acf = []
    for lag in range(1, min(n // 2, 100)):
        c = np.mean((x[:-lag] - mean) * (x[lag:] - mean)) / var
        acf.append(c)
        if c < 0.05:
            break
    return float(n / (1 + 2 * np.sum(acf)))


def _autocov_fft(x):
    n = len(x)
    m = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(x - np.mean(x), m)
    return np.fft.irfft(f * np.conj(f), m)[:n] / n


def ess_geyer(samples) -> float:
    """Geyer's initial positive sequence: sum autocorrelation pairs while their sum stays positive."""
    x = np.asarray(samples, dtype=np.float64)
    n = len(x)
    if n < 4:
        return float(n)
    acov = _autocov_fft(x)
    if acov[0] <= 0:
        return float(n)
    rho = acov / acov[0]
    tau = -1.0
    for k in range(0, n - 1, 2):
        pair = rho[k] + rho[k + 1]
        if pair < 0:
            break
        tau += 2.0 * pair
    tau = max(tau, 1.0 / n)
    return float(n / tau)


def min_ess(draws: dict, num_chains: int = 1, estimator=ess_geyer, max_chains: int = 256) -> float:
    """Sum over chains (a strided subset of at most `max_chains`, scaled up) of per-chain ESS, minimum over
    scalar parameters.  `draws[name]` is (S,) / (S, n) when num_chains == 1, else (C, S) / (C, S, n)."""
    best = None
    for v in draws.values():
        a = np.asarray(v)
        if num_chains == 1:
            a = a[None]
        if a.ndim == 2:
            a = a[:, :, None]
        C = a.shape[0]
        idx = range(0, C, max(1, C // max_chains))
        for j in range(a.shape[2]):
            tot = sum(estimator(a[c, :, j]) for c in idx) * (C / len(idx))
            best = tot if best is None else min(best, tot)
    return float(best)
