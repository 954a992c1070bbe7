"""Pieces shared by the three sampler front-ends."""
from __future__ import annotations

import numpy as np
import torch

from .. import core as mx
from ..engine import ChainState, DeviceModel, compile_model


def philox_seed(key, random_seed=None) -> int:
    """64-bit Philox seed from an ``mx.random.key`` (hmc/nuts) or an integer seed (metropolis)."""
    if key is not None:
        if isinstance(key, (int, np.integer)):
            return int(key)
        return key.philox_seed()
    return int(random_seed or 0)


JIT_MIN_CHAINS = 1024      # jit='auto': a ~3 s NVRTC compile per model pays off from about this many lock-step chains


def prepare(log_prob_fn, initial_params, num_chains, step_size, chain_offset, model=None, theta0=None, cache=True,
            transforms=None, jit="auto"):
    """Trace/compile (cached by the content of the trace, engine.compile_model) and build the per-chain state.
    `theta0` ([num_chains, D], in the sampler's coordinates) overrides the common starting point -- used to hand
    the warm-up's final positions to the sampling call."""
    model = model if isinstance(model, DeviceModel) else compile_model(log_prob_fn, initial_params, cache=cache,
                                                                       transforms=transforms)
    if jit not in ("auto", True, False, "on", "off"):
        raise ValueError(f"Unknown jit option: {jit}")
    if not getattr(model, "jit", False) and (jit in (True, "on") or (jit == "auto" and num_chains >= JIT_MIN_CHAINS)):
        # pointwise class: kernels specialised to this model's term table (mlx_mcmc_b200/jit.py); bit-identical results
        from ..jit import specialize
        specialize(model, True if jit in (True, "on") else "auto")
    if theta0 is not None:
        theta0 = torch.as_tensor(theta0)          # a host array is copied to the device here (pinned staging)
        if tuple(theta0.shape) != (num_chains, model.D):
            raise ValueError(f"theta0 has shape {tuple(theta0.shape)}, expected {(num_chains, model.D)}")
        if not theta0.is_cuda:
            staged = torch.empty(theta0.shape, dtype=torch.float32, pin_memory=True)
            staged.copy_(theta0)
            theta = staged.to(model.device, non_blocking=True)
        else:
            theta = theta0.to(device=model.device, dtype=torch.float32).clone()
    else:
        theta = model.pack(initial_params, num_chains)
    return model, ChainState(model, theta, step_size, chain_offset)


def alloc_draws(model, n_iter, n_chains):
    return torch.empty((n_iter, n_chains, model.D), dtype=torch.float32, device=model.device)


class SamplerInfo(dict):
    """Diagnostics returned next to the draws when ``return_info=True`` (per-chain device tensors
    moved to numpy): step sizes, accept counts, tree depths, divergences, gradient-evaluation counts."""
    __getattr__ = dict.__getitem__


# ------------------------------------------------------------------------------------------ mass-matrix adaptation
# SURVEY.md 8(f) row 3.  The reference integrates with the identity mass matrix (hmc.py:102-111, nuts.py:113-117) and
# lists "mass matrix adaptation for HMC/NUTS" as planned (README.md:165,220).  What is built here is the windowed
# scheme of Stan's warm-up on the device: a fast initial buffer (step size only), slow windows of doubling length at
# whose end the diagonal metric is re-estimated from the window's draws (pooled over ALL chains of the call: with
# thousands of chains a 25-iteration window already holds 10^5 draws per coordinate), and a fast final buffer.
def warmup_windows(num_warmup: int, init_buffer: int = 75, term_buffer: int = 50, base_window: int = 25):
    """[(start, end, update_metric_at_end)] covering [0, num_warmup).  Stan's schedule: 75 / 25-50-100-... / 50, shrunk
    to 15 % / 75 % / 10 % when the warm-up is shorter than 150 iterations; under 20 iterations: one fast segment."""
    n = int(num_warmup)
    if n < 20:
        return [(0, n, False)]
    if init_buffer + base_window + term_buffer > n:
        init_buffer, term_buffer = int(0.15 * n), int(0.1 * n)
        base_window = n - init_buffer - term_buffer
    segs = [(0, init_buffer, False)] if init_buffer > 0 else []
    start, size, stop = init_buffer, base_window, n - term_buffer
    while start < stop:
        end = start + size
        if end + 2 * size > stop:        # the next window would not fit: stretch this one to the end of the slow phase
            end = stop
        segs.append((start, end, True))
        start, size = end, 2 * size
    if term_buffer > 0:
        segs.append((stop, n, False))
    return segs


def mass_from_window(model, draws_u):
    """inv_mass [D] (device float32) = regularised pooled variance of the window's draws (sampler coordinates,
    [S, C, D]) -- computed by b2m_mass_from_draws, nothing travels to the host."""
    import ctypes as C
    from .. import _cabi
    S, Cn, D = draws_u.shape
    out = torch.empty(D, dtype=torch.float32, device=draws_u.device)
    with torch.cuda.device(draws_u.device):
        _cabi.check(model.lib.b2m_mass_from_draws(C.c_void_p(draws_u.data_ptr()), S, Cn, D, C.c_void_p(out.data_ptr()),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def rescale_step(step_size, old_inv_mass, new_inv_mass):
    """Keep the position step eps * sqrt(M^-1) (geometric mean over the coordinates) when the metric changes, so the
    restarted dual averaging does not have to climb orders of magnitude through maximum-depth trees."""
    f = float(torch.exp(0.5 * (old_inv_mass.double().log().mean() - new_inv_mass.double().log().mean())).item())
    step_size.mul_(f)
    return f
