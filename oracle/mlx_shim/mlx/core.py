"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A small torch-CPU stand-in for the subset of ``mlx.core`` that korentomas/mlx-mcmc
touches (SURVEY.md section 8c lists the symbols).  Apple MLX is not installable in
the build container or on the GPU box, so this module lets

  * the *unmodified* reference source under /root/reference run here
    (``oracle/make_golden.py`` -> fixtures in ``tests/golden/``), and
  * the restatement in ``oracle/refport`` run anywhere (tests, bench cpu_baseline).

Semantics kept from MLX: float32 default dtype, python scalars are weakly typed,
float64 numpy input is down-cast, ``array([traced scalars])`` stacks and keeps the
autograd graph, ``grad(fn, argnums)`` returns zeros for unused arguments,
splittable random keys.  The random stream is NOT MLX's threefry stream (that cannot
be reproduced without MLX); every draw can be recorded so the same numbers can be
injected into the CUDA path (parity check 2).
"""
from __future__ import annotations

import builtins as _b
import hashlib
import math
from typing import Callable, List, Sequence

import numpy as _np
import torch as _torch

_torch.set_num_threads(1)

inf = math.inf
pi = math.pi
nan = math.nan

# dtype tokens -----------------------------------------------------------------
float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool_ = _torch.bool

_FLOAT = [_torch.float32]  # default float dtype; the arbiter switches it to float64


def set_default_float(dtype) -> None:
    """float32 = reference semantics; float64 = arbiter for the 1e-5 checks."""
    _FLOAT[0] = dtype


def default_float():
    return _FLOAT[0]


# array ------------------------------------------------------------------------
def _raw(x):
    """python/numpy/array -> torch tensor with MLX-like dtype defaults."""
    if isinstance(x, array):
        return x.t
    if isinstance(x, _torch.Tensor):
        return x
    if isinstance(x, bool):
        return _torch.tensor(x)
    if isinstance(x, int):
        return _torch.tensor(x, dtype=_torch.int32)
    if isinstance(x, float):
        return _torch.tensor(x, dtype=_FLOAT[0])
    if isinstance(x, _np.ndarray) or isinstance(x, _np.generic):
        a = _np.asarray(x)
        if a.dtype.kind == "f":
            return _torch.tensor(a.astype(_np.float64), dtype=_FLOAT[0])
        if a.dtype.kind in "iu":
            return _torch.tensor(a.astype(_np.int64), dtype=_torch.int32)
        if a.dtype.kind == "b":
            return _torch.tensor(a)
        raise TypeError(f"unsupported numpy dtype {a.dtype}")
    if isinstance(x, (list, tuple)):
        if len(x) and _b.any(isinstance(e, (array, _torch.Tensor)) for e in x):
            return _torch.stack([_raw(e) for e in x])
        return _raw(_np.asarray(x))
    raise TypeError(f"cannot convert {type(x)} to array")


def _weak(other, like: _torch.Tensor):
    """Operand promotion: python scalars adopt the array's dtype (weak typing)."""
    if isinstance(other, array):
        return other.t
    if isinstance(other, bool):
        return _torch.tensor(other)
    if isinstance(other, (int, float)):
        if like.dtype.is_floating_point:
            return _torch.tensor(float(other), dtype=like.dtype)
        if isinstance(other, float):
            return _torch.tensor(other, dtype=_FLOAT[0])
        return _torch.tensor(other, dtype=like.dtype)
    return _raw(other)


class array:
    """Thin wrapper around a CPU torch tensor with the mlx.core.array surface used by the reference."""

    __slots__ = ("t",)
    __array_priority__ = 1000

    def __init__(self, value, dtype=None):
        t = _raw(value)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        self.t = t

    # -- introspection
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def size(self):
        return int(self.t.numel())

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def dtype(self):
        return self.t.dtype

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield array(self.t[i])

    def __float__(self):
        return float(self.t.detach())

    def __int__(self):
        return int(self.t.detach())

    def __bool__(self):
        return bool(self.t.detach())

    def item(self):
        return self.t.detach().item()

    def tolist(self):
        return self.t.detach().tolist()

    def astype(self, dtype):
        return array(self.t.to(dtype))

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"array({self.t.detach().tolist()}, dtype={self.t.dtype})"

    def __getitem__(self, idx):
        if isinstance(idx, array):
            idx = idx.t.long()
        return array(self.t[idx])

    # -- arithmetic
    def __add__(self, o):
        return array(self.t + _weak(o, self.t))

    __radd__ = __add__

    def __sub__(self, o):
        return array(self.t - _weak(o, self.t))

    def __rsub__(self, o):
        return array(_weak(o, self.t) - self.t)

    def __mul__(self, o):
        return array(self.t * _weak(o, self.t))

    __rmul__ = __mul__

    def __truediv__(self, o):
        a, b = self.t, _weak(o, self.t)
        if not a.dtype.is_floating_point and not b.dtype.is_floating_point:
            a, b = a.to(_FLOAT[0]), b.to(_FLOAT[0])
        return array(a / b)

    def __rtruediv__(self, o):
        a, b = _weak(o, self.t), self.t
        if not a.dtype.is_floating_point and not b.dtype.is_floating_point:
            a, b = a.to(_FLOAT[0]), b.to(_FLOAT[0])
        return array(a / b)

    def __pow__(self, o):
        return array(_torch.pow(self.t, _weak(o, self.t)))

    def __rpow__(self, o):
        return array(_torch.pow(_weak(o, self.t), self.t))

    def __neg__(self):
        return array(-self.t)

    def __matmul__(self, o):
        return matmul(self, o)

    # -- comparisons / logic
    def __lt__(self, o):
        return array(self.t < _weak(o, self.t))

    def __le__(self, o):
        return array(self.t <= _weak(o, self.t))

    def __gt__(self, o):
        return array(self.t > _weak(o, self.t))

    def __ge__(self, o):
        return array(self.t >= _weak(o, self.t))

    def __eq__(self, o):  # noqa: D105
        return array(self.t == _weak(o, self.t))

    def __ne__(self, o):
        return array(self.t != _weak(o, self.t))

    __hash__ = None

    def __and__(self, o):
        return array(self.t & _weak(o, self.t))

    def __or__(self, o):
        return array(self.t | _weak(o, self.t))

    def __invert__(self):
        return array(~self.t)


def _f(x) -> _torch.Tensor:
    """tensor for a float-valued elementwise function (python scalars -> default float)."""
    t = _raw(x)
    return t if t.dtype.is_floating_point else t.to(_FLOAT[0])


def log(x):
    return array(_torch.log(_f(x)))


def exp(x):
    return array(_torch.exp(_f(x)))


def sqrt(x):
    return array(_torch.sqrt(_f(x)))


def abs(x):  # noqa: A001
    return array(_torch.abs(_raw(x)))


def square(x):
    t = _raw(x)
    return array(t * t)


def sum(x, axis=None, keepdims=False):  # noqa: A001
    t = _raw(x)
    if axis is None:
        return array(t.sum())
    return array(t.sum(dim=axis, keepdim=keepdims))


def mean(x, axis=None):
    t = _f(x)
    return array(t.mean() if axis is None else t.mean(dim=axis))


def var(x, axis=None):
    t = _f(x)
    return array(t.var(unbiased=False) if axis is None else t.var(dim=axis, unbiased=False))


def std(x, axis=None):
    t = _f(x)
    return array(t.std(unbiased=False) if axis is None else t.std(dim=axis, unbiased=False))


def max(x, axis=None):  # noqa: A001
    t = _raw(x)
    return array(t.max() if axis is None else t.max(dim=axis).values)


def min(x, axis=None):  # noqa: A001
    t = _raw(x)
    return array(t.min() if axis is None else t.min(dim=axis).values)


def argmax(x, axis=None):
    t = _raw(x)
    return array((t.argmax() if axis is None else t.argmax(dim=axis)).to(_torch.int32))


def cumsum(x, axis=None):
    t = _raw(x)
    return array(t.flatten().cumsum(0) if axis is None else t.cumsum(dim=axis))


def expand_dims(x, axis):
    return array(_raw(x).unsqueeze(axis))


def where(cond, a, b):
    c = _raw(cond)
    ta = a.t if isinstance(a, array) else None
    tb = b.t if isinstance(b, array) else None
    like = ta if ta is not None else (tb if tb is not None else _torch.zeros((), dtype=_FLOAT[0]))
    ta = ta if ta is not None else _weak(a, like)
    tb = tb if tb is not None else _weak(b, like)
    if ta.dtype != tb.dtype:
        dt = _torch.promote_types(ta.dtype, tb.dtype)
        ta, tb = ta.to(dt), tb.to(dt)
    return array(_torch.where(c, ta, tb))


def all(x):  # noqa: A001
    return array(_raw(x).all())


def any(x):  # noqa: A001
    return array(_raw(x).any())


def allclose(a, b, rtol=1e-5, atol=1e-8):
    ta, tb = _f(a), _f(b)
    return array(_torch.tensor(bool(_torch.allclose(ta, tb, rtol=rtol, atol=atol))))


def matmul(a, b):
    ta, tb = _raw(a), _raw(b)
    if ta.dtype != tb.dtype:       # observed float32 data against float64 parameters (arbiter mode)
        dt = _torch.promote_types(ta.dtype, tb.dtype)
        ta, tb = ta.to(dt), tb.to(dt)
    return array(ta @ tb)


def stack(xs, axis=0):
    return array(_torch.stack([_raw(x) for x in xs], dim=axis))


def zeros(shape, dtype=None):
    return array(_torch.zeros(shape, dtype=dtype or _FLOAT[0]))


def ones(shape, dtype=None):
    return array(_torch.ones(shape, dtype=dtype or _FLOAT[0]))


# autodiff ---------------------------------------------------------------------
def grad(fn: Callable, argnums=0):
    """mx.grad: gradient of a scalar function w.r.t. positional args (zeros when unused)."""
    single = isinstance(argnums, int)
    nums: Sequence[int] = [argnums] if single else list(argnums)

    def wrapped(*args):
        args = list(args)
        leaves = []
        for i in nums:
            leaf = _f(args[i]).detach().clone().requires_grad_(True)
            leaves.append(leaf)
            args[i] = array(leaf)
        out = fn(*args)
        out_t = _raw(out)
        if not out_t.requires_grad:
            gs = [None] * len(leaves)
        else:
            gs = _torch.autograd.grad(out_t, leaves, allow_unused=True)
        res = tuple(array(_torch.zeros_like(l) if g is None else g) for l, g in zip(leaves, gs))
        return res[0] if single else res

    return wrapped


def value_and_grad(fn: Callable, argnums=0):
    g = grad(fn, argnums)

    def wrapped(*args):
        return fn(*args), g(*args)

    return wrapped


def eval(*_a):  # noqa: A001  (MLX lazy-eval no-op)
    return None


# random -----------------------------------------------------------------------
class _Key:
    """Splittable key = path of integers; draws come from a numpy Philox keyed by a hash of the path."""

    __slots__ = ("path",)

    def __init__(self, path):
        self.path = tuple(int(p) for p in path)

    def _gen(self):
        h = hashlib.blake2b(repr(self.path).encode(), digest_size=16).digest()
        return _np.random.Generator(_np.random.Philox(key=int.from_bytes(h, "little") % (1 << 128)))

    def __repr__(self):
        return f"Key{self.path}"


class _Random:
    def __init__(self):
        self.tape: List[tuple] | None = None  # when a list: every draw is appended as (kind, np.ndarray)

    def key(self, seed: int) -> _Key:
        return _Key((int(seed),))

    def split(self, key: _Key, num: int = 2):
        return [_Key(key.path + (i,)) for i in range(num)]

    def normal(self, shape=(), dtype=None, loc=0.0, scale=1.0, key: _Key | None = None):
        z = key._gen().standard_normal(tuple(shape)).astype(_np.float32)
        if self.tape is not None:
            self.tape.append(("normal", z.copy()))
        out = array(_torch.tensor(z.astype(_np.float64), dtype=_FLOAT[0]))
        if loc != 0.0 or scale != 1.0:
            out = out * scale + loc
        return out

    def uniform(self, low=0.0, high=1.0, shape=(), dtype=None, key: _Key | None = None):
        u = key._gen().random(tuple(shape), dtype=_np.float32)
        # keep strictly inside (0,1) so log(u) is finite, as MLX's uniform never returns exactly 1
        u = _np.clip(u, _np.float32(2.0**-24), _np.float32(1.0 - 2.0**-24))
        if self.tape is not None:
            self.tape.append(("uniform", u.copy()))
        out = array(_torch.tensor(u.astype(_np.float64), dtype=_FLOAT[0]))
        if low != 0.0 or high != 1.0:
            out = out * (high - low) + low
        return out

    def randint(self, low, high, shape=(), dtype=None, key: _Key | None = None):
        v = key._gen().integers(int(low), int(high), size=tuple(shape))
        return array(_torch.tensor(_np.asarray(v), dtype=_torch.int32))


random = _Random()
