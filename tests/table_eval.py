"""float64 numpy evaluation of a traced term table (test infrastructure).

Independent of the CUDA library: it states, in the plainest possible form, what a `TracedModel` MEANS
(include/b200mcmc.h, SURVEY.md 8(a2')), so the CPU suite can check that tracing a `log_prob` preserves its value
and gradient against the oracle before any kernel is involved."""
import math

import numpy as np

from mlx_mcmc_b200 import tracer as T

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _operand(o, theta, arrays, n_len):
    if o.kind == T.OP_CONST:
        return np.full(n_len, o.c)
    if o.kind == T.OP_PARAM:
        return np.full(n_len, theta[o.a])
    if o.kind == T.OP_DATA:
        return np.broadcast_to(np.asarray(arrays[o.a], dtype=np.float64), (n_len,))
    if o.kind == T.OP_PARAMVEC:
        assert o.b == n_len
        return theta[o.a:o.a + o.b]
    if o.kind == T.OP_LIN:
        out = np.full(n_len, o.c)
        for param, array, coef in o.lin:
            v = coef * (np.asarray(arrays[array], dtype=np.float64) if array >= 0 else 1.0)
            out = out + v * (theta[param] if param >= 0 else 1.0)
        return out
    if o.kind == T.OP_MATVEC:
        X = np.asarray(arrays[o.a], dtype=np.float64)
        return X @ theta[o.b:o.b + X.shape[1]] + o.c
    raise AssertionError(f"unknown operand kind {o.kind}")


def _logpdf(dist, x, p0, p1, k):
    with np.errstate(divide="ignore", invalid="ignore"):
        if dist == T.NORMAL:
            return -_HALF_LOG_2PI - np.log(p1) - 0.5 * (x - p0) ** 2 / p1 ** 2
        if dist == T.HALFNORMAL:
            return np.where(x >= 0, math.log(2.0) - _HALF_LOG_2PI - np.log(p0) - 0.5 * x ** 2 / p0 ** 2, -np.inf)
        if dist == T.EXPONENTIAL:
            return np.where(x >= 0, np.log(p0) - p0 * x, -np.inf)
        if dist == T.GAMMA:      # k = (alpha, lgamma(alpha), -), p0 = rate
            return np.where(x > 0, k[0] * np.log(p0) - k[1] + (k[0] - 1.0) * np.log(x) - p0 * x, -np.inf)
        if dist == T.BETA:       # k = (a, b, log B(a, b))
            return np.where((x > 0) & (x < 1), (k[0] - 1.0) * np.log(x) + (k[1] - 1.0) * np.log1p(-x) - k[2], -np.inf)
        if dist == T.CONSTANT:
            return np.full_like(x, k[0])
    raise AssertionError(f"unknown distribution tag {dist}")


def table_logp(model, theta):
    """log p(theta) of a TracedModel at a flat float64 parameter vector."""
    theta = np.asarray(theta, dtype=np.float64)
    total = 0.0
    for t in model.terms:
        x, p0, p1 = (_operand(o, theta, model.arrays, t.length) for o in (t.x, t.p0, t.p1))
        total += t.weight * float(np.sum(_logpdf(t.dist, x, p0, p1, t.k)))
    return total


def table_grad(model, theta, h=1e-6):
    """central differences of table_logp (float64: good to ~1e-8 relative for these smooth densities)"""
    theta = np.asarray(theta, dtype=np.float64)
    g = np.zeros_like(theta)
    for i in range(theta.size):
        step = h * max(1.0, abs(theta[i]))
        up, dn = theta.copy(), theta.copy()
        up[i] += step
        dn[i] -= step
        g[i] = (table_logp(model, up) - table_logp(model, dn)) / (2.0 * step)
    return g
