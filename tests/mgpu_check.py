"""Multi-GPU parity check, run as   torchrun --nproc-per-node 2 tests/mgpu_check.py   (NCCL, one rank per GPU).

1. chain sharding: the gathered draws of a sharded run equal the draws of the same global chains run on one GPU,
   bit for bit (Philox streams are keyed by global chain id);
2. observation sharding: log p / gradient of the row-sharded model agree with the unsharded model to float32
   rounding and are bit-identical on every rank;
3. observation-sharded NUTS: every rank produces bit-identical draws, and they agree with the closed-form posterior;
4. the same with sliced state, both forms -- 'nccl' (reduce-scatter / all-gather inside the library) and 'peer' (the
   peer window: K6's epilogue stores gradient tiles into the owner's window over NVLink, release / acquire flags): the
   merged draws are identical on every rank, start out equal to the replicated schedule's (same Philox streams,
   rounding-level differences only) and agree with the closed-form posterior; counters match the depth record.
Prints one line `MGPU_CHECK OK ...` from rank 0 on success; any failure raises on the failing rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mlx_mcmc_b200 as B  # noqa: E402
import mlx_mcmc_b200.core as mx  # noqa: E402
from mlx_mcmc_b200 import dist as D, workloads as W  # noqa: E402
from mlx_mcmc_b200.engine import compile_model  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    report = {}

    # ---- 1. chain sharding (pointwise HMC and GLM NUTS)
    fn, init, meta = W.c2_event_rate(B.ns)
    kw = dict(num_samples=40, num_warmup=60, step_size=0.1, num_leapfrog_steps=10, key=mx.random.key(5))
    sharded, rate, _ = D.run_sharded(fn, init, method="hmc", num_chains=37, shard="chains", **kw)
    whole, rate1 = B.hmc(fn, init, num_chains=37, **kw)
    assert sharded["rate"].shape == whole["rate"].shape == (37, 40)
    assert np.array_equal(sharded["rate"], whole["rate"]), "chain-sharded HMC draws differ from the single-GPU run"
    assert abs(rate - rate1) < 1e-12
    report["chains_hmc"] = "bit-equal"
    # the same through the user-facing call: MCMC.run(..., shard='chains') gathers every rank's draws
    mc = B.MCMC(fn)
    got = mc.run(init, num_samples=40, num_warmup=60, method="hmc", step_size=0.1, num_leapfrog_steps=10, random_seed=5,
                 num_chains=37, shard="chains", verbose=False)
    ref1 = B.MCMC(fn).run(init, num_samples=40, num_warmup=60, method="hmc", step_size=0.1, num_leapfrog_steps=10,
                          random_seed=5, num_chains=37, verbose=False)
    assert got["rate"].shape == (37, 40) and np.array_equal(got["rate"], ref1["rate"]), "MCMC.run(shard='chains') differs"
    report["mcmc_run_shard_chains"] = "bit-equal"

    # ---- 2. observation sharding: value + gradient
    fr, initr, mr = W.regression(B.ns, 6000, 48, seed=2)
    full = compile_model(fr, initr, cache=False)
    shard = D.compile_obs_sharded(fr, initr)
    rng = np.random.default_rng(0)
    theta = torch.from_numpy((mr.beta_true[None] + 0.05 * rng.standard_normal((256, 48))).astype(np.float32)).cuda()
    lp_f, g_f = full.logp_grad(theta)
    lp_s, g_s = shard.logp_grad(theta)
    torch.cuda.synchronize()
    e_lp = float(((lp_s - lp_f).abs() / lp_f.abs()).max())
    e_g = float((g_s - g_f).abs().max() / g_f.abs().max())
    assert e_lp < 2e-6 and e_g < 5e-6, (e_lp, e_g)
    both = [torch.empty_like(g_s) for _ in range(world)]
    td.all_gather(both, g_s)
    assert all(torch.equal(both[0], b) for b in both), "ranks disagree on the all-reduced gradient"
    lps = [torch.empty_like(lp_s) for _ in range(world)]
    td.all_gather(lps, lp_s)
    assert all(torch.equal(lps[0], b) for b in lps)
    report["obs_logp_rel"], report["obs_grad_normwise"] = e_lp, e_g

    # ---- 3. observation-sharded NUTS
    s, r, info = D.run_sharded(fr, initr, method="nuts", num_chains=128, shard="obs", num_samples=60, num_warmup=120,
                               step_size=0.05, compat="correct", key=mx.random.key(3), return_torch=True)
    draws = s["beta"].contiguous()
    allr = [torch.empty_like(draws) for _ in range(world)]
    td.all_gather(allr, draws)
    assert all(torch.equal(allr[0], a) for a in allr), "ranks diverged during observation-sharded NUTS"
    m, V = W.regression_posterior(mr)
    mean = draws.double().mean(dim=(0, 1)).cpu().numpy()
    z = np.abs(mean - m) / np.sqrt(np.diag(V) / (128 * 60 / 4))      # generous ESS guess: 1/4 of the draws
    assert z.max() < 6.0, z.max()
    report["obs_nuts_max_z"] = float(z.max())
    report["obs_nuts_grad_evals"] = info.grad_evals

    # ---- 4. sliced state: NCCL reduce-scatter / all-gather, and the peer window (K6 push epilogue over NVLink)
    kw = dict(method="nuts", num_chains=256 * world, shard="obs", num_samples=40, num_warmup=80, step_size=0.05,
              compat="correct", key=mx.random.key(4), return_torch=True)
    s_rep, r_rep, i_rep = D.run_sharded(fr, initr, **kw)
    d_rep = s_rep["beta"].contiguous()
    for mode in ("nccl", "peer"):
        s_sl, r_sl, i_sl = D.run_sharded(fr, initr, slice_state=mode, **kw)
        d_sl = s_sl["beta"].contiguous()
        assert d_sl.shape == d_rep.shape == (256 * world, 40, 48)
        allr = [torch.empty_like(d_sl) for _ in range(world)]
        td.all_gather(allr, d_sl)
        assert all(torch.equal(allr[0], a) for a in allr), f"ranks hold different merged draws after a sliced run ({mode})"
        # first kept draw: same Philox streams.  The NCCL forms sum in the same order as the replicated all-reduce (equal
        # to rounding); the peer form sums per-source slots in rank order, so a chain whose slice / multinomial decision
        # sits on its float32 boundary may take another -- equally valid -- branch: nearly all chains must agree.
        per_chain = (d_sl[:, 0] - d_rep[:, 0]).abs().amax(dim=1)
        first = float(per_chain.median())
        agree = float((per_chain < 5e-4).float().mean())
        assert agree >= (0.97 if mode == "peer" else 1.0) and first < 5e-5, (mode, agree, first)
        assert torch.isfinite(d_sl).all()
        mean = d_sl.double().mean(dim=(0, 1)).cpu().numpy()
        z = np.abs(mean - m) / np.sqrt(np.diag(V) / (256 * world * 40 / 4))
        assert z.max() < 6.0, (mode, z.max())
        assert i_sl.depths.shape == i_rep.depths.shape and i_sl.depths.min() >= 1
        leaves = int((2 ** i_sl.depths.astype(np.int64) - 1).sum())   # leaves of the sampling phase follow from the depths
        assert i_sl.grad_evals - i_sl.warmup_grad_evals <= leaves + i_sl.depths.size * 2, (mode, i_sl.grad_evals, leaves)
        assert abs(i_sl.grad_evals - i_rep.grad_evals) < 0.05 * i_rep.grad_evals, (mode, i_sl.grad_evals, i_rep.grad_evals)
        assert abs(r_sl - r_rep) < 0.05, (mode, r_sl, r_rep)
        same_depth = float((i_sl.depths == i_rep.depths).mean())
        report[f"sliced_{mode}_first_draw_median_maxdiff"] = first
        report[f"sliced_{mode}_first_draw_chains_agreeing"] = agree
        report[f"sliced_{mode}_max_z"] = float(z.max())
        report[f"sliced_{mode}_same_depths"] = same_depth
        report[f"sliced_{mode}_grad_evals"] = (i_sl.grad_evals, i_rep.grad_evals)
    # a second peer call on a fresh model: sequence numbers and flags of the windows start over cleanly
    s2, _, _ = D.run_sharded(fr, initr, slice_state="peer", **{**kw, "num_samples": 10, "num_warmup": 20})
    assert torch.isfinite(s2["beta"]).all()

    td.barrier()
    if rank == 0:
        print("MGPU_CHECK OK", world, report)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
