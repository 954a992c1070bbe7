"""Distribution interface (reference: mlx_mcmc/distributions/base.py:6-54) plus the shared dispatch
between the two evaluation modes of ``log_prob``:

  * any argument traced  -> one term of the B200 term table (tracer.TraceContext.log_density);
  * everything concrete  -> host numpy float32, same formula and support mask as the reference
                            (a convenience of the classes; the samplers never take this branch).
"""
from __future__ import annotations

import numpy as np

from ..tracer import Sym, active_context


def f32(x):
    """concrete argument -> numpy float32 (0-d for scalars), as mx.array() would make it"""
    return np.asarray(x, dtype=np.float32)


def traced(*vals) -> bool:
    return any(isinstance(v, Sym) for v in vals)


def require_concrete(name, v):
    """Shape parameters that the reference converts with float() at construction (beta.py:53-54,
    gamma.py:55) cannot be traced there either."""
    if isinstance(v, Sym):
        v._concretise(f"{name} (a shape parameter is evaluated with float() at construction)")
    return v


def context():
    ctx = active_context()
    if ctx is None:  # pragma: no cover - a symbol escaped its trace
        raise RuntimeError("traced value used outside of MCMC.run / trace()")
    return ctx


class Distribution:
    """Base class: subclasses implement ``log_prob(value)`` and ``sample(key, shape)``."""

    def log_prob(self, value):
        raise NotImplementedError(f"{self.__class__.__name__} must implement log_prob()")

    def sample(self, key, shape=()):
        raise NotImplementedError(f"{self.__class__.__name__} must implement sample()")

    def __repr__(self):
        return f"{self.__class__.__name__}()"
