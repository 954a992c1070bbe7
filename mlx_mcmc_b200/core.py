"""``mx``-like namespace for user ``log_prob`` functions: ``import mlx_mcmc_b200.core as mx``.

Mirrors the part of ``mlx.core`` that models written for korentomas/mlx-mcmc call (SURVEY.md 8b item 5:
``mx.array`` incl. a list of traced scalars, ``mx.sum``, ``mx.inf``, arithmetic operators; north-star
additions ``mx.matmul`` / ``@``).  Concrete values are numpy float32 arrays; traced values are the
symbols of :mod:`mlx_mcmc_b200.tracer`.  Nonlinear functions of a *traced* value (``mx.log(mu)`` ...)
are not part of the lowered term-table family and raise ``UnsupportedOpError``.
"""
from __future__ import annotations

import builtins as _b
import math

import numpy as np

from .tracer import Lin, LogProb, LogProbStack, Sym, UnsupportedOpError, active_context

inf = math.inf
pi = math.pi
nan = math.nan
float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool_ = np.bool_


def array(value, dtype=None):
    """mx.array: python/numpy -> float32 (ints -> int32); traced values pass through; a list of traced
    scalar log-probabilities becomes a stack that ``sum`` can reduce."""
    if isinstance(value, Sym):
        return value
    if isinstance(value, (list, tuple)) and _b.any(isinstance(v, Sym) for v in value):
        ctx = active_context()
        items = []
        for v in value:
            if isinstance(v, LogProb):
                items.append(v)
            elif isinstance(v, Sym):
                raise UnsupportedOpError("unsupported op: mx.array([...]) of traced values that are not log-probabilities")
            else:
                items.append(LogProb(ctx, [])._coerce(v))
        return LogProbStack(ctx, items)
    a = np.asarray(value)
    if dtype is not None:
        return a.astype(dtype)
    if a.dtype.kind == "f":
        return a.astype(np.float32)
    if a.dtype.kind in "iu":
        return a.astype(np.int32)
    return a


def _no_trace(name, x):
    if isinstance(x, Sym):
        raise UnsupportedOpError(f"unsupported op: mx.{name} of a traced value (only library distributions and "
                                 "affine expressions of the parameters are lowered to CUDA)")


def _f32(x):
    a = np.asarray(x)
    return a.astype(np.float32) if a.dtype.kind != "f" or a.dtype == np.float64 else a


def log(x):
    _no_trace("log", x)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(_f32(x))


def exp(x):
    _no_trace("exp", x)
    return np.exp(_f32(x))


def sqrt(x):
    _no_trace("sqrt", x)
    with np.errstate(invalid="ignore"):
        return np.sqrt(_f32(x))


def abs(x):  # noqa: A001
    _no_trace("abs", x)
    return np.abs(np.asarray(x))


def square(x):
    _no_trace("square", x)
    a = _f32(x)
    return a * a


def sum(x, axis=None, keepdims=False):  # noqa: A001
    if isinstance(x, (LogProb, LogProbStack)):
        if axis not in (None, 0, -1):
            raise UnsupportedOpError("unsupported op: mx.sum over an axis other than the observation axis")
        return x.total()
    if isinstance(x, Lin):
        raise UnsupportedOpError("unsupported op: mx.sum of a raw traced expression (sum log_prob terms instead)")
    return np.sum(np.asarray(x), axis=axis, keepdims=keepdims)


def mean(x, axis=None):
    _no_trace("mean", x)
    return np.mean(_f32(x), axis=axis)


def var(x, axis=None):
    _no_trace("var", x)
    return np.var(_f32(x), axis=axis)


def std(x, axis=None):
    _no_trace("std", x)
    return np.std(_f32(x), axis=axis)


def max(x, axis=None):  # noqa: A001
    _no_trace("max", x)
    return np.max(np.asarray(x), axis=axis)


def min(x, axis=None):  # noqa: A001
    _no_trace("min", x)
    return np.min(np.asarray(x), axis=axis)


def argmax(x, axis=None):
    _no_trace("argmax", x)
    return np.argmax(np.asarray(x), axis=axis).astype(np.int32)


def cumsum(x, axis=None):
    _no_trace("cumsum", x)
    return np.cumsum(np.asarray(x), axis=axis)


def expand_dims(x, axis):
    _no_trace("expand_dims", x)
    return np.expand_dims(np.asarray(x), axis)


def where(cond, a, b):
    for v in (cond, a, b):
        _no_trace("where", v)
    return np.where(np.asarray(cond), a, b)


def all(x):  # noqa: A001
    _no_trace("all", x)
    return np.all(np.asarray(x))


def any(x):  # noqa: A001
    _no_trace("any", x)
    return np.any(np.asarray(x))


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return bool(np.allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol))


def matmul(a, b):
    if isinstance(b, Sym):
        return b.__rmatmul__(a)
    if isinstance(a, Sym):
        return a.__matmul__(b)
    return np.matmul(np.asarray(a), np.asarray(b))


def stack(xs, axis=0):
    if _b.any(isinstance(x, Sym) for x in xs):
        return array(list(xs))
    return np.stack([np.asarray(x) for x in xs], axis=axis)


def zeros(shape, dtype=np.float32):
    return np.zeros(shape, dtype=dtype)


def ones(shape, dtype=np.float32):
    return np.ones(shape, dtype=dtype)


def eval(*_a):  # noqa: A001 - MLX's lazy-evaluation hook is a no-op here
    return None


# ---------------------------------------------------------------------------------------
class _Key:
    """Splittable key as in ``mx.random.key``: a seed plus a split path.  The samplers derive the
    64-bit Philox seed from it; host-side ``Distribution.sample`` uses it to seed numpy."""

    __slots__ = ("seed", "path")

    def __init__(self, seed, path=()):
        self.seed, self.path = int(seed), tuple(path)

    def philox_seed(self) -> int:
        h = self.seed & 0xFFFFFFFFFFFFFFFF
        for p in self.path:  # splitmix-style fold of the split path
            h = (h ^ (p + 0x9E3779B97F4A7C15 + ((h << 6) & 0xFFFFFFFFFFFFFFFF) + (h >> 2))) & 0xFFFFFFFFFFFFFFFF
        return h

    def _rng(self):
        return np.random.Generator(np.random.Philox(key=self.philox_seed()))

    def __repr__(self):
        return f"key({self.seed}{''.join('/%d' % p for p in self.path)})"


class _Random:
    @staticmethod
    def key(seed):
        return _Key(seed)

    @staticmethod
    def split(key, num=2):
        return [_Key(key.seed, key.path + (i,)) for i in range(num)]

    @staticmethod
    def normal(shape=(), dtype=np.float32, loc=0.0, scale=1.0, key=None):
        rng = key._rng() if key is not None else np.random.default_rng()
        return (rng.standard_normal(tuple(shape)) * scale + loc).astype(np.float32)

    @staticmethod
    def uniform(low=0.0, high=1.0, shape=(), dtype=np.float32, key=None):
        rng = key._rng() if key is not None else np.random.default_rng()
        return (rng.random(tuple(shape)) * (high - low) + low).astype(np.float32)

    @staticmethod
    def randint(low, high, shape=(), dtype=np.int32, key=None):
        rng = key._rng() if key is not None else np.random.default_rng()
        return np.asarray(rng.integers(int(low), int(high), size=tuple(shape)), dtype=np.int32)


random = _Random()
