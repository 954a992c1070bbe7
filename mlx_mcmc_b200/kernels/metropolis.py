"""Random-walk Metropolis front-end with the reference's signature (mlx_mcmc/kernels/metropolis.py:6-13)."""
from __future__ import annotations

from ..engine import launch_mh
from ._common import SamplerInfo, alloc_draws, philox_seed, prepare


def metropolis_hastings(
    log_prob_fn,
    initial_params,
    num_samples=1000,
    proposal_scale=0.1,
    random_seed=0,
    verbose=False,
    *,
    num_chains: int = 1,
    chain_offset: int = 0,
    lanes: int = 0,
    return_torch: bool = False,
    return_info: bool = False,
    model=None,
    theta0=None,
    cache: bool = True,
    jit="auto",
):
    """Same arguments and return value as the reference: ``(samples, acceptance_rate)`` where
    ``samples[name]`` holds ``num_samples`` values (a numpy array here, a python list there) and the
    rate is accepted / num_samples (metropolis.py:99).  One launch of `mh_kernel` runs all steps for
    all chains with the current log-prob cached on device (metropolis.py:55,87)."""
    seed = philox_seed(None, random_seed)
    model, st = prepare(log_prob_fn, initial_params, num_chains, 0.0, chain_offset, model, theta0, cache, None, jit)
    draws = alloc_draws(model, num_samples, num_chains)
    launch_mh(st, num_samples, float(proposal_scale), seed, 0, draws=draws, lanes=lanes)
    rate = float(st.n_accept.double().sum().item() / max(num_samples * num_chains, 1))
    if verbose:
        print(f"Running {num_samples} Metropolis-Hastings iterations x {num_chains} chain(s): accept rate {rate:.2%}")
    samples = model.unpack(draws, squeeze_chain=(num_chains == 1), to_numpy=not return_torch)
    if return_info:
        return samples, rate, SamplerInfo(n_accept=st.n_accept.cpu().numpy(), state=st, model=model,
                                          value_evals=int(num_chains) * num_samples)
    return samples, rate
