"""Gamma distribution, shape/rate parameterisation (reference: mlx_mcmc/distributions/gamma.py)."""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln

from .. import core as mx
from ..tracer import GAMMA
from .base import Distribution, context, f32, require_concrete, traced


class Gamma(Distribution):
    """Gamma(alpha, beta); log p(x) = alpha log(beta) - lgamma(alpha) + (alpha-1) log x - beta x for x > 0,
    -inf otherwise (gamma.py:53-88).  `alpha` must be concrete (the reference calls float() on it,
    gamma.py:55); `beta` and the value may be traced."""

    def __init__(self, alpha, beta=1.0):
        self.alpha = f32(require_concrete("Gamma alpha", alpha))
        self.beta = beta if traced(beta) else f32(beta)
        self._lgamma_alpha = np.float32(gammaln(np.float64(self.alpha)))

    def log_prob(self, value):
        if traced(value, self.beta):
            return context().log_density(GAMMA, value, self.beta, None, k=(float(self.alpha), float(self._lgamma_alpha), 0.0))
        x = f32(value)
        with np.errstate(divide="ignore", invalid="ignore"):
            inside = (self.alpha * np.log(self.beta) - self._lgamma_alpha) + (self.alpha - 1) * np.log(x) - self.beta * x
        return np.where(x > 0, inside, np.float32(-np.inf)).astype(np.float32)

    def _device_sample_spec(self):
        require_concrete("Gamma beta", self.beta)
        return GAMMA, float(self.alpha), float(self.beta), None

    def sample(self, key, shape=()):
        # the reference falls back to numpy's sampler seeded from the key (gamma.py:107-117)
        seed = int(mx.random.randint(0, 2 ** 31 - 1, key=key))
        rng = np.random.default_rng(seed)
        return rng.gamma(float(self.alpha), scale=1.0 / float(self.beta), size=shape).astype(np.float32)

    def mean(self):
        return self.alpha / self.beta

    def variance(self):
        return self.alpha / self.beta ** 2

    def mode(self):
        return np.where(self.alpha >= 1, (self.alpha - 1) / self.beta, np.float32(0.0)).astype(np.float32)

    def __repr__(self):
        b = "<traced>" if traced(self.beta) else f"{float(self.beta):.3f}"
        return f"Gamma(alpha={float(self.alpha):.3f}, beta={b})"
