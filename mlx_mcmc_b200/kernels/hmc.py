"""Hamiltonian Monte Carlo front-end with the reference's signature (mlx_mcmc/kernels/hmc.py:7-17).

The reference's Python loops (warm-up :159-176, sampling :185-198) become two kernel launches; every
chain runs whole trajectories inside the persistent kernel `hmc_kernel` (csrc/pointwise.cu).
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np
import torch

from .. import _cabi
from ..engine import launch_hmc
from ._common import (SamplerInfo, alloc_draws, mass_from_window, philox_seed, prepare, rescale_step,
                      warmup_windows)


def hmc(
    log_prob_fn: Callable,
    initial_params: Dict[str, object],
    num_samples: int = 1000,
    num_warmup: int = 1000,
    step_size: float = 0.1,
    num_leapfrog_steps: int = 10,
    adapt_step_size: bool = True,
    target_accept: float = 0.8,
    key=None,
    *,
    num_chains: int = 1,
    adapt: str = "reference",
    chain_offset: int = 0,
    lanes: int = 0,
    return_torch: bool = False,
    return_info: bool = False,
    model=None,
    theta0=None,
    adapt_mass_matrix: bool = False,
    transforms=None,
    cache: bool = True,
    jit="auto",
) -> Tuple[Dict[str, object], float]:
    """Same arguments and return value as the reference's ``hmc``: ``(samples, acceptance_rate)`` with
    ``samples[name]`` of shape ``(num_samples,)`` (sampling-phase acceptance rate, hmc.py:200-206).

    Extensions (keyword only): ``num_chains`` independent chains in lock-step (draws get a leading
    chain axis when > 1); ``adapt='reference'`` is the reference's +-5 % rule on the cumulative
    acceptance rate (hmc.py:164-170), ``adapt='dual_averaging'`` runs Hoffman-Gelman dual averaging on
    device; ``chain_offset`` is the global id of chain 0 (multi-GPU chain sharding);
    ``return_torch`` keeps the draws on the device.  ``adapt_mass_matrix`` / ``transforms='auto'``: windowed
    diagonal-metric adaptation and log / logit coordinates for constrained parameters, as in ``nuts`` (the reference has
    neither: identity mass at hmc.py:102-111; both are on its roadmap, README.md:165,220, PROGRESS.md:119).
    """
    if num_warmup == 0:
        # the reference divides by the number of warm-up iterations (hmc.py:175)
        raise ZeroDivisionError("division by zero")
    if adapt not in ("reference", "dual_averaging"):
        raise ValueError(f"Unknown adapt mode: {adapt}")
    if adapt_mass_matrix and num_warmup < 20:
        raise ValueError("adapt_mass_matrix=True needs num_warmup >= 20 (the metric is estimated in warm-up windows)")
    seed = philox_seed(key, 0)
    model, st = prepare(log_prob_fn, initial_params, num_chains, step_size, chain_offset, model, theta0, cache, transforms, jit)
    mode = _cabi.ADAPT_NONE
    if adapt_step_size:
        mode = _cabi.ADAPT_REFERENCE if adapt == "reference" else _cabi.ADAPT_DUAL_AVERAGING
        if mode == _cabi.ADAPT_DUAL_AVERAGING:
            st.da_state[:, 1] = 0.0                           # log eps_bar
            st.da_state[:, 2] = float(np.log(10.0 * step_size))  # mu
    inv_mass = None
    if adapt_mass_matrix and num_warmup >= 20:
        inv_mass = torch.ones(model.D, dtype=torch.float32, device=model.device)
        origin = 0
        for (w0, w1, update) in warmup_windows(num_warmup):
            n_store = min(w1 - w0, max(8, int(4e9 // max(num_chains * model.D * 4, 1)))) if update else 0
            if w1 - w0 - n_store > 0:
                launch_hmc(st, w1 - w0 - n_store, num_leapfrog_steps, mode, target_accept, seed, w0, lanes=lanes,
                           inv_mass=inv_mass, adapt_origin=origin)
            if n_store:
                window = torch.empty((n_store, num_chains, model.D), dtype=torch.float32, device=model.device)
                launch_hmc(st, n_store, num_leapfrog_steps, mode, target_accept, seed, w1 - n_store, draws=window,
                           lanes=lanes, inv_mass=inv_mass, adapt_origin=origin, draws_unconstrained=True)
                new_mass = mass_from_window(model, window)
                del window
                rescale_step(st.step_size, inv_mass, new_mass)
                inv_mass = new_mass
                if mode == _cabi.ADAPT_DUAL_AVERAGING:      # restart around the rescaled step size
                    st.da_state[:, 0] = 0.0
                    st.da_state[:, 1] = st.step_size.log()
                    st.da_state[:, 2] = (10.0 * st.step_size).log()
                    origin = w1
    else:
        launch_hmc(st, num_warmup, num_leapfrog_steps, mode, target_accept, seed, 0, lanes=lanes)
    if mode == _cabi.ADAPT_DUAL_AVERAGING:
        st.step_size.copy_(st.da_state[:, 1].exp())
    warm_accept = (st.n_accept.double().sum() / st.n_total.double().sum().clamp(min=1)).item() if return_info else None
    st.reset_counters()
    draws = alloc_draws(model, num_samples, num_chains)
    launch_hmc(st, num_samples, num_leapfrog_steps, _cabi.ADAPT_NONE, target_accept, seed, num_warmup,
               draws=draws, lanes=lanes, inv_mass=inv_mass)
    rate = float((st.n_accept.double().sum() / st.n_total.double().sum().clamp(min=1)).item())
    samples = model.unpack(draws, squeeze_chain=(num_chains == 1), to_numpy=not return_torch)
    if return_info:
        info = SamplerInfo(step_size=st.step_size.cpu().numpy(), n_accept=st.n_accept.cpu().numpy(),
                           inv_mass=None if inv_mass is None else inv_mass.cpu().numpy(),
                           warmup_accept_rate=warm_accept,
                           grad_evals=int(num_chains) * (num_warmup + num_samples) * num_leapfrog_steps,
                           state=st, model=model)
        return samples, rate, info
    return samples, rate
