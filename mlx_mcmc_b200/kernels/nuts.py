"""No-U-Turn sampler front-end with the reference's signature (mlx_mcmc/kernels/nuts.py:16-26)."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np
import torch

from .. import _cabi
from ..engine import launch_nuts
from ._common import (SamplerInfo, alloc_draws, mass_from_window, philox_seed, prepare, rescale_step,
                      warmup_windows)


def _sliced_world(model, num_chains: int, schedule: str) -> int:
    """World size when a sliced observation-sharded schedule applies to `model`, else an error (never a silent
    fall-back: the per-chain outputs of a sliced call must be merged, those of a replicated call must not)."""
    comm = getattr(model, "_comm", None)
    if comm is None or comm.world < 2:
        raise ValueError("slice_state needs an observation-sharded model (dist.compile_obs_sharded) on > 1 ranks")
    if model.model_class != 1:
        raise ValueError("slice_state needs a GLM-class model")
    if num_chains % comm.world or num_chains % 256:
        raise ValueError("slice_state: num_chains must be a multiple of 256 and of the number of ranks")
    if schedule == "sync":
        raise ValueError("slice_state needs the asynchronous NUTS schedule")
    return comm.world


def nuts(
    log_prob_fn: Callable,
    initial_params: Dict[str, object],
    num_samples: int = 1000,
    num_warmup: int = 1000,
    step_size: float = 0.1,
    max_tree_depth: int = 10,
    adapt_step_size: bool = True,
    target_accept: float = 0.65,
    key=None,
    *,
    num_chains: int = 1,
    compat: str = "reference",
    step_size_adaptation: str = "per_chain",
    step_size_jitter: float = 0.0,
    chain_offset: int = 0,
    lanes: int = 0,
    return_torch: bool = False,
    return_info: bool = False,
    model=None,
    theta0=None,
    slice_state=False,
    schedule: str = "async",
    adapt_mass_matrix: bool = False,
    transforms=None,
    cache: bool = True,
    jit="auto",
) -> Tuple[Dict[str, object], float]:
    """Same arguments and return value as the reference's ``nuts``: ``(samples, rate)`` where rate is the
    fraction of sampling iterations whose mean acceptance statistic exceeded 0.5 (nuts.py:341,353) and
    vector parameters come back as ``(num_samples, D)`` (nuts.py:338-339).

    ``compat='reference'`` reproduces the reference's float32 slice round trip and NaN handling
    (SURVEY.md F6/F7); ``compat='correct'`` keeps the slice in log space (see csrc/nuts_pointwise.cu).
    Dual averaging follows nuts.py:62-68,298-310 per chain on device; after warm-up every chain
    switches to its averaged step size (nuts.py:317-320).

    ``step_size_adaptation='pooled'`` (an extension for ``num_chains > 1``; the reference has one chain):
    all chains share one step size, so a lock-step batch stays at one tree depth.  GLM-class models run ONE
    dual-averaging recurrence on the mean acceptance statistic over the chains, on device, every iteration;
    pointwise models adapt per chain and every chain then takes the median of the averaged step sizes.

    ``step_size_jitter=j`` (extension, 0 = the reference's behaviour): each chain integrates iteration t with
    ``eps * (1 + j * (2u - 1))``.  The reference's U-turn test looks at positions only (nuts.py:119-135); on a nearly
    isotropic Gaussian posterior it cannot see a turn when 2^k - 1 steps are just over a whole number of oscillation
    periods, and trees then run to ``max_tree_depth`` (measured at the 1000 x 100K regression: mean depth 7.7 instead
    of 4 at eps = 1.34e-3).  A jitter of 0.1-0.2 removes the resonance.

    ``slice_state`` (observation-sharded GLM models only, see dist.py): during the sampling phase rank r advances only
    the chains of its slice and draws / counters of the slices are merged over ``torch.distributed`` before returning.
    ``'peer'`` (or True when the model has a peer window attached): K6's epilogue stores every finished gradient tile
    straight into its owner's window over NVLink and the owner publishes the next leaf's packed rows to every rank --
    no collective on the critical path.  ``'nccl'``: reduce-scatter + all-gather calls between the kernels.

    ``schedule='sync'`` (GLM class) selects the synchronous lock-step schedule instead of the iteration-asynchronous one.

    ``adapt_mass_matrix=True`` (extension; the reference's mass matrix is the identity, nuts.py:113-117, adaptation is on
    its roadmap, README.md:165,220): windowed warm-up as in Stan -- a diagonal metric re-estimated on the device from
    the pooled draws of all chains at the end of slow windows of doubling length, dual averaging restarted after every
    update.  ``transforms='auto'`` (extension, PROGRESS.md:119): parameters that are the value of a HalfNormal /
    Exponential / Gamma (Beta) term are sampled in log (logit) coordinates with the log-Jacobian added, which removes
    the hard walls that freeze the reference's NUTS; draws come back in the model's own coordinates.

    ``jit`` (pointwise-class models): 'auto' compiles kernels specialised to this model's term table with NVRTC when the
    call runs >= 1024 chains (mlx_mcmc_b200/jit.py; ~3 s once per model, cached on disk), True forces it, False keeps the
    generic interpreter kernels.  The draws are bit-identical either way."""
    if num_warmup == 0:
        raise ZeroDivisionError("division by zero")   # nuts.py:322-323
    if compat not in ("reference", "correct"):
        raise ValueError(f"Unknown compat mode: {compat}")
    if step_size_adaptation not in ("per_chain", "pooled"):
        raise ValueError(f"Unknown step_size_adaptation: {step_size_adaptation}")
    if not 0.0 <= step_size_jitter < 1.0:
        raise ValueError("step_size_jitter must be in [0, 1)")
    if not 1 <= max_tree_depth <= _cabi.MAX_TREE_DEPTH:
        raise ValueError(f"max_tree_depth must be in 1..{_cabi.MAX_TREE_DEPTH}")
    if schedule not in ("async", "sync"):
        raise ValueError(f"Unknown schedule: {schedule}")
    if slice_state not in (False, True, "peer", "nccl"):
        raise ValueError(f"Unknown slice_state: {slice_state}")
    if adapt_mass_matrix and (not adapt_step_size or num_warmup < 20):
        raise ValueError("adapt_mass_matrix=True needs adapt_step_size=True and num_warmup >= 20 (the metric is estimated in "
                         "warm-up windows and the step size re-adapted after every update)")
    sched = _cabi.SCHED_SYNC if schedule == "sync" else _cabi.SCHED_ASYNC
    cmode = _cabi.COMPAT_REFERENCE if compat == "reference" else _cabi.COMPAT_CORRECT
    seed = philox_seed(key, 0)
    model, st = prepare(log_prob_fn, initial_params, num_chains, step_size, chain_offset, model, theta0, cache, transforms, jit)
    st.da_state[:, 0] = 0.0                                           # H_bar
    st.da_state[:, 1] = 1.0                                           # eps_bar
    st.da_state[:, 2] = float(np.log(np.float32(10.0 * step_size)))   # mu, a float32 in the reference
    amode = _cabi.ADAPT_DUAL_AVERAGING if adapt_step_size else _cabi.ADAPT_NONE
    pooled = adapt_step_size and step_size_adaptation == "pooled"
    if pooled and model.model_class == 1:
        amode = _cabi.ADAPT_POOLED
    warm_depths = torch.empty((num_warmup, num_chains), dtype=torch.int32, device=model.device) if return_info else None
    inv_mass = None
    if adapt_mass_matrix and amode != _cabi.ADAPT_NONE and num_warmup >= 20:
        # windowed warm-up (kernels/_common.py): step size only / metric at the end of each slow window / step size only
        inv_mass = torch.ones(model.D, dtype=torch.float32, device=model.device)
        origin = 0
        for (w0, w1, update) in warmup_windows(num_warmup):
            n_store = min(w1 - w0, max(8, int(4e9 // max(num_chains * model.D * 4, 1)))) if update else 0
            wd = warm_depths[w0:w1] if return_info else None
            if w1 - w0 - n_store > 0:
                launch_nuts(st, w1 - w0 - n_store, max_tree_depth, amode, cmode, target_accept, seed, w0,
                            depths=None if wd is None else wd[: w1 - w0 - n_store], lanes=lanes,
                            step_size_jitter=step_size_jitter, inv_mass=inv_mass, schedule=sched, adapt_origin=origin)
            if n_store:
                window = torch.empty((n_store, num_chains, model.D), dtype=torch.float32, device=model.device)
                launch_nuts(st, n_store, max_tree_depth, amode, cmode, target_accept, seed, w1 - n_store, draws=window,
                            depths=None if wd is None else wd[w1 - w0 - n_store:], lanes=lanes,
                            step_size_jitter=step_size_jitter, inv_mass=inv_mass, schedule=sched, adapt_origin=origin,
                            draws_unconstrained=True)
                new_mass = mass_from_window(model, window)
                del window
                rescale_step(st.step_size, inv_mass, new_mass)
                inv_mass = new_mass
                # restart dual averaging around the rescaled step size (Stan: mu = log(10 eps), H_bar = 0)
                st.da_state[:, 0] = 0.0
                st.da_state[:, 1] = 1.0
                st.da_state[:, 2] = (10.0 * st.step_size).log().float().double()
                origin = w1
    else:
        launch_nuts(st, num_warmup, max_tree_depth, amode, cmode, target_accept, seed, 0, depths=warm_depths, lanes=lanes,
                    step_size_jitter=step_size_jitter, schedule=sched)
    if adapt_step_size:
        st.step_size.copy_(st.da_state[:, 1])
        if pooled and model.model_class != 1:
            st.step_size.fill_(float(st.da_state[:, 1].median().item()))
    warm_leaves = st.n_leaves.clone()
    st.n_accept.zero_()
    draws = alloc_draws(model, num_samples, num_chains)
    depths = torch.empty((num_samples, num_chains), dtype=torch.int32, device=model.device)
    smode = _cabi.SLICE_OFF
    if slice_state:
        _sliced_world(model, num_chains, schedule)
        if slice_state == "peer" or (slice_state is True and getattr(model, "_peer", None) is not None):
            if getattr(model, "_peer", None) is None:
                raise ValueError("slice_state='peer' needs a peer window (dist.compile_obs_sharded(..., peer_chains=num_chains))")
            smode = _cabi.SLICE_PEER
        else:
            smode = _cabi.SLICE_NCCL
    if smode != _cabi.SLICE_OFF:
        # rank r writes draws / depths / counters of its slice only: start from zeros and sum the slices afterwards
        from ..dist import merge_slices
        draws.zero_()
        depths.zero_()
        base = [t.clone() for t in (st.n_leaves, st.n_diverge)]
    launch_nuts(st, num_samples, max_tree_depth, _cabi.ADAPT_NONE, cmode, target_accept, seed, num_warmup,
                draws=draws, depths=depths, lanes=lanes, step_size_jitter=step_size_jitter, inv_mass=inv_mass,
                schedule=sched, slice_state=smode)
    if smode != _cabi.SLICE_OFF:
        merge_slices((st.n_accept, draws, depths), zip((st.n_leaves, st.n_diverge), base))
    rate = float(st.n_accept.double().sum().item() / max(num_samples * num_chains, 1))
    samples = model.unpack(draws, squeeze_chain=(num_chains == 1), to_numpy=not return_torch)
    if return_info:
        info = SamplerInfo(step_size=st.step_size.cpu().numpy(), depths=depths.cpu().numpy(),
                           inv_mass=None if inv_mass is None else inv_mass.cpu().numpy(),
                           warmup_depths=warm_depths.cpu().numpy(), n_diverge=st.n_diverge.cpu().numpy(),
                           grad_evals=int(st.n_leaves.sum().item()), warmup_grad_evals=int(warm_leaves.sum().item()),
                           state=st, model=model)
        return samples, rate, info
    return samples, rate
