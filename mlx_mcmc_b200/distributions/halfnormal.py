"""Half-normal distribution (reference: mlx_mcmc/distributions/halfnormal.py)."""
from __future__ import annotations

import numpy as np

from .. import core as mx
from ..tracer import HALFNORMAL
from .base import Distribution, context, f32, require_concrete, traced

_HALF_LOG_2PI = np.float32(0.5) * np.log(np.float32(2.0 * np.pi), dtype=np.float32)
_LOG2 = np.log(np.float32(2.0), dtype=np.float32)


class HalfNormal(Distribution):
    """|N(0, scale)|; log p(x) = log 2 - 0.5 log(2 pi) - log(scale) - 0.5 x^2/scale^2 for x >= 0,
    -inf otherwise (halfnormal.py:55-63)."""

    def __init__(self, scale):
        self.scale = scale if traced(scale) else f32(scale)

    def log_prob(self, value):
        if traced(value, self.scale):
            return context().log_density(HALFNORMAL, value, self.scale)
        x = f32(value)
        with np.errstate(divide="ignore", invalid="ignore"):
            inside = (_LOG2 - _HALF_LOG_2PI - np.log(self.scale)) - np.float32(0.5) * (x ** 2) / (self.scale ** 2)
        return np.where(x >= 0, inside, np.float32(-np.inf)).astype(np.float32)

    def _device_sample_spec(self):
        require_concrete("HalfNormal scale", self.scale)
        return HALFNORMAL, float(self.scale), 0.0, None

    def sample(self, key, shape=()):
        return np.abs(mx.random.normal(shape, key=key) * self.scale)

    def __repr__(self):
        if traced(self.scale):
            return "HalfNormal(<traced>)"
        return f"HalfNormal(scale={float(self.scale):.3f})"
