"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Restatement of the six log-densities of korentomas/mlx-mcmc on the ``mlx.core``
stand-in (oracle/mlx_shim).  Each ``log_prob`` keeps the reference's operation
order so float32 rounding matches the reference run on the same stand-in
(`oracle/make_golden.py` asserts bit-equality against /root/reference).

Reference lines restated:
  Normal       mlx_mcmc/distributions/normal.py:27-56
  HalfNormal   mlx_mcmc/distributions/halfnormal.py:28-63
  Exponential  mlx_mcmc/distributions/exponential.py:37-71
  Gamma        mlx_mcmc/distributions/gamma.py:40-88
  Beta         mlx_mcmc/distributions/beta.py:37-91
  Categorical  mlx_mcmc/distributions/categorical.py:38-93
"""
from __future__ import annotations

import mlx.core as mx  # the oracle stand-in; callers put oracle/mlx_shim on sys.path
from scipy.special import gammaln

_NEG_INF = -mx.inf


class _Density:
    def log_prob(self, value):  # pragma: no cover - interface
        raise NotImplementedError


class Normal(_Density):
    def __init__(self, loc, scale):
        self.loc, self.scale = mx.array(loc), mx.array(scale)
        self.log_scale = mx.log(self.scale)
        self.half_log_2pi_neg = -0.5 * mx.log(2 * mx.pi)

    def log_prob(self, value):
        x = mx.array(value)
        variance = self.scale ** 2
        return self.half_log_2pi_neg - self.log_scale - 0.5 * ((x - self.loc) ** 2) / variance


class HalfNormal(_Density):
    def __init__(self, scale):
        self.scale = mx.array(scale)
        self.log_scale = mx.log(self.scale)
        self.half_log_2pi_neg = -0.5 * mx.log(2 * mx.pi)
        self.log_two = mx.log(mx.array(2.0))

    def log_prob(self, value):
        x = mx.array(value)
        variance = self.scale ** 2
        inside = self.log_two + self.half_log_2pi_neg - self.log_scale - 0.5 * (x ** 2) / variance
        return mx.where(x >= 0, inside, mx.array(_NEG_INF))


class Exponential(_Density):
    def __init__(self, rate):
        self.rate = mx.array(rate)

    def log_prob(self, value):
        x = mx.array(value)
        inside = mx.log(self.rate) - self.rate * x
        return mx.where(x >= 0, inside, mx.array(_NEG_INF))


class Gamma(_Density):
    def __init__(self, alpha, beta=1.0):
        self.alpha, self.beta = mx.array(alpha), mx.array(beta)
        a = float(self.alpha) if self.alpha.size == 1 else self.alpha  # concrete shape only
        self.log_norm = self.alpha * mx.log(self.beta) - mx.array(gammaln(a))

    def log_prob(self, value):
        x = mx.array(value)
        inside = self.log_norm + (self.alpha - 1) * mx.log(x) - self.beta * x
        return mx.where(x > 0, inside, mx.array(_NEG_INF))


class Beta(_Density):
    def __init__(self, alpha, beta):
        self.alpha, self.beta = mx.array(alpha), mx.array(beta)
        a = float(self.alpha) if self.alpha.size == 1 else self.alpha
        b = float(self.beta) if self.beta.size == 1 else self.beta
        self.log_B = mx.array(gammaln(a) + gammaln(b) - gammaln(a + b))

    def log_prob(self, value):
        x = mx.array(value)
        inside = (self.alpha - 1) * mx.log(x) + (self.beta - 1) * mx.log(1 - x) - self.log_B
        return mx.where((x > 0) & (x < 1), inside, mx.array(_NEG_INF))


class Categorical(_Density):
    def __init__(self, probs=None, logits=None):
        if probs is None and logits is None:
            raise ValueError("Either probs or logits must be specified")
        if probs is not None and logits is not None:
            raise ValueError("Only one of probs or logits can be specified")
        if probs is not None:
            p = mx.array(probs)
            self.probs = p / mx.sum(p)
            self.logits = mx.log(self.probs)
        else:
            self.logits = mx.array(logits)  # raw logits are what log_prob gathers (reference quirk)
            shifted = mx.exp(self.logits - mx.max(self.logits))
            self.probs = shifted / mx.sum(shifted)
        self.num_categories = self.probs.shape[0]

    def log_prob(self, value):
        k = mx.array(value, dtype=mx.int32)
        ok = (k >= 0) & (k < self.num_categories)
        safe = mx.where(ok, k, mx.array(0, dtype=mx.int32))  # gather needs an in-range index
        return mx.where(ok, self.logits[safe], mx.array(_NEG_INF))
