"""ORACLE / TEST INFRASTRUCTURE ONLY -- the checker's namespace (mx stand-in + restated classes).

Importing this module puts oracle/mlx_shim on sys.path so that `import mlx.core` resolves to the
torch-CPU stand-in.  If a real `mlx` is installed it is NOT shadowed silently: we refuse, because the
restatement is pinned against the stand-in's float32 semantics.
"""
import os
import sys
from types import SimpleNamespace

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mlx_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

import mlx.core as mx  # noqa: E402

if "mlx_shim" not in os.path.abspath(mx.__file__):  # pragma: no cover
    raise ImportError("oracle expects the mlx.core stand-in (oracle/mlx_shim), found " + mx.__file__)

from oracle.refport import distributions as _d  # noqa: E402
from oracle.refport import samplers as _s  # noqa: E402

ns = SimpleNamespace(
    mx=mx, Normal=_d.Normal, HalfNormal=_d.HalfNormal, Beta=_d.Beta, Gamma=_d.Gamma,
    Exponential=_d.Exponential, Categorical=_d.Categorical,
)
samplers = _s
Tape = _s.Tape


def value_and_grad(log_prob_fn, params: dict, dtype="float32"):
    """(logp, {name: grad}) of a namespace-generic model at `params`, in float32 (reference
    semantics) or float64 (arbiter)."""
    import numpy as np
    import torch
    prev = mx.default_float()
    mx.set_default_float(torch.float64 if dtype == "float64" else torch.float32)
    try:
        names = list(params)
        vals = [mx.array(np.asarray(params[n], dtype=np.float64)) for n in names]
        fn = lambda *a: log_prob_fn(dict(zip(names, a)))  # noqa: E731
        lp = fn(*vals)
        g = mx.grad(fn, argnums=list(range(len(names))))(*vals)
        return np.asarray(lp, dtype=np.float64), {n: np.asarray(gi, dtype=np.float64) for n, gi in zip(names, g)}
    finally:
        mx.set_default_float(prev)
