"""Normal distribution (reference: mlx_mcmc/distributions/normal.py)."""
from __future__ import annotations

import numpy as np

from .. import core as mx
from ..tracer import NORMAL
from .base import Distribution, context, f32, require_concrete, traced

_HALF_LOG_2PI = np.float32(0.5) * np.log(np.float32(2.0 * np.pi), dtype=np.float32)


class Normal(Distribution):
    """Normal(loc, scale); log p(x) = -0.5 log(2 pi) - log(scale) - 0.5 (x - loc)^2 / scale^2
    (normal.py:49-56).  loc, scale and the value may each be traced."""

    def __init__(self, loc, scale):
        self.loc = loc if traced(loc) else f32(loc)
        self.scale = scale if traced(scale) else f32(scale)

    def log_prob(self, value):
        if traced(value, self.loc, self.scale):
            return context().log_density(NORMAL, value, self.loc, self.scale)
        x = f32(value)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (-_HALF_LOG_2PI - np.log(self.scale)) - np.float32(0.5) * ((x - self.loc) ** 2) / (self.scale ** 2)

    def _device_sample_spec(self):
        require_concrete("Normal loc", self.loc), require_concrete("Normal scale", self.scale)
        return NORMAL, float(self.loc), float(self.scale), None

    def sample(self, key, shape=()):
        return mx.random.normal(shape, key=key) * self.scale + self.loc

    def __repr__(self):
        if traced(self.loc, self.scale):
            return "Normal(<traced>)"
        return f"Normal(loc={float(self.loc):.3f}, scale={float(self.scale):.3f})"
