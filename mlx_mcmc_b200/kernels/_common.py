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


def prepare(log_prob_fn, initial_params, num_chains, step_size, chain_offset, model=None, theta0=None):
    """Trace/compile (cached) and build the per-chain state.  `theta0` ([num_chains, D] device tensor)
    overrides the common starting point -- used to hand the warm-up's final positions to the sampling call."""
    model = model if isinstance(model, DeviceModel) else compile_model(log_prob_fn, initial_params)
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
