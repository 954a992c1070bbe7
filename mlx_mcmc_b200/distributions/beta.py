"""Beta distribution (reference: mlx_mcmc/distributions/beta.py)."""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln

from .. import core as mx
from ..tracer import BETA
from .base import Distribution, context, f32, require_concrete, traced


class Beta(Distribution):
    """Beta(alpha, beta); log p(x) = (a-1) log x + (b-1) log(1-x) - log B(a,b) for 0 < x < 1, -inf otherwise
    (beta.py:53-91).  Both shape parameters must be concrete (float() at beta.py:53-54)."""

    def __init__(self, alpha, beta):
        self.alpha = f32(require_concrete("Beta alpha", alpha))
        self.beta = f32(require_concrete("Beta beta", beta))
        a, b = np.float64(self.alpha), np.float64(self.beta)
        self._log_beta_const = np.float32(gammaln(a) + gammaln(b) - gammaln(a + b))

    def log_prob(self, value):
        if traced(value):
            return context().log_density(BETA, value, None, None,
                                         k=(float(self.alpha), float(self.beta), float(self._log_beta_const)))
        x = f32(value)
        with np.errstate(divide="ignore", invalid="ignore"):
            inside = (self.alpha - 1) * np.log(x) + (self.beta - 1) * np.log(1 - x) - self._log_beta_const
        return np.where((x > 0) & (x < 1), inside, np.float32(-np.inf)).astype(np.float32)

    def _device_sample_spec(self):
        return BETA, float(self.alpha), float(self.beta), None

    def sample(self, key, shape=()):
        seed = int(mx.random.randint(0, 2 ** 31 - 1, key=key))   # numpy fallback as in beta.py:110-119
        rng = np.random.default_rng(seed)
        return rng.beta(float(self.alpha), float(self.beta), size=shape).astype(np.float32)

    def mean(self):
        return self.alpha / (self.alpha + self.beta)

    def variance(self):
        s = self.alpha + self.beta
        return (self.alpha * self.beta) / (s ** 2 * (s + 1))

    def mode(self):
        return (self.alpha - 1) / (self.alpha + self.beta - 2)

    def __repr__(self):
        return f"Beta(alpha={float(self.alpha):.3f}, beta={float(self.beta):.3f})"
