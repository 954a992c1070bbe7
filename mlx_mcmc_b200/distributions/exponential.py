"""Exponential distribution (reference: mlx_mcmc/distributions/exponential.py)."""
from __future__ import annotations

import numpy as np

from .. import core as mx
from ..tracer import EXPONENTIAL
from .base import Distribution, context, f32, require_concrete, traced


class Exponential(Distribution):
    """Exponential(rate); log p(x) = log(rate) - rate x for x >= 0, -inf otherwise (exponential.py:61-71)."""

    def __init__(self, rate):
        self.rate = rate if traced(rate) else f32(rate)

    def log_prob(self, value):
        if traced(value, self.rate):
            return context().log_density(EXPONENTIAL, value, self.rate)
        x = f32(value)
        with np.errstate(divide="ignore", invalid="ignore"):
            inside = np.log(self.rate) - self.rate * x
        return np.where(x >= 0, inside, np.float32(-np.inf)).astype(np.float32)

    def _device_sample_spec(self):
        require_concrete("Exponential rate", self.rate)
        return EXPONENTIAL, float(self.rate), 0.0, None

    def sample(self, key, shape=()):
        u = mx.random.uniform(shape=shape, key=key)
        return (-np.log1p(-u) / self.rate).astype(np.float32)   # inverse CDF (exponential.py:88-91)

    def mean(self):
        return np.float32(1.0) / self.rate

    def variance(self):
        return np.float32(1.0) / self.rate ** 2

    def mode(self):
        return np.float32(0.0)

    def median(self):
        return np.log(np.float32(2.0)) / self.rate

    def __repr__(self):
        return "Exponential(<traced>)" if traced(self.rate) else f"Exponential(rate={float(self.rate):.3f})"
