#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: leapfrog grad-evals/sec, min-ESS/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c5|...] [--impl reference]

Default workload (N=1) is BASELINE.json configs[1]: examples/04_event_rates.py Gamma/Exponential rate
model, 65,536 independent chains, HMC (step 0.1, 10 leapfrog steps), chains sharded over the GPUs.

One *step* = one launch of the persistent HMC kernel covering ITERS whole HMC iterations (momentum
draw, L leapfrog steps with a fused log-density+gradient each, Metropolis accept, draw written to HBM)
for every chain of the rank.  A grad-eval = one value-and-gradient of log_prob for one chain; a step
performs chains x ITERS x L of them (the gradient at the trajectory start is the cached one).

The JSON line also carries
  e2e          the same metric through the public API (`hmc(...)`: host initial values -> device, warm-up
               and sampling launches, draws copied back to pinned host memory), per step;
  roofline     HBM roofline of the dominant kernel (algorithmic bytes = draws written + chain state
               read/written per launch) -- this kernel is FP32-issue bound, see `issue`;
  cpu_baseline the oracle restatement of the reference's HMC (oracle/refport, torch-CPU stand-in for MLX)
               timed on one host core on a bounded sample of the same workload.
`--impl reference` times that same restatement on all host cores (one independent chain per core).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (model factory, method, chains at N=1, sampler kwargs, iterations per step)
    "c2": dict(model="c2_event_rate", method="hmc", chains=65536, step_size=0.1, L=10, iters=100,
               desc="examples/04_event_rates Gamma/Exponential rate model, 65536 chains, HMC eps0=0.1 L=10"),
    "c1": dict(model="c1_normal", method="hmc", chains=65536, step_size=0.01, L=10, iters=50,
               desc="examples/01-02 Normal(mu,sigma) posterior, 100 obs, HMC L=10"),
    "c5": dict(model="c5_ab_test", method="metropolis", chains=1048576, proposal_scale=0.02, iters=100,
               desc="examples/03_ab_testing Beta A/B model, 1M Metropolis chains"),
}
METRIC = "leapfrog_grad_evals_per_sec"
UNIT = "grad-evals/s"


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx_ = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx_) if mx_ else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU baseline
def _cpu_chain(args):
    """one chain of the oracle restatement (runs in a worker process)"""
    wl_name, seed, n_warm, n_samp = args
    from oracle.ns import Tape, ns as ons, samplers
    from mlx_mcmc_b200 import workloads as W
    wl = WORKLOADS[wl_name]
    fn, init, _ = W.ALL_SMALL[wl["model"]](ons)
    tape = Tape()
    t0 = time.perf_counter()
    if wl["method"] == "hmc":
        s, _, _ = samplers.hmc_port(fn, init, num_samples=n_samp, num_warmup=n_warm, step_size=wl["step_size"],
                                    num_leapfrog_steps=wl["L"], key=ons.mx.random.key(seed), tape=tape)
        evals = tape.leapfrogs
    else:
        s, _ = samplers.run_port(fn, init, num_samples=n_samp, num_warmup=n_warm, method="metropolis",
                                 proposal_scale=wl["proposal_scale"], random_seed=seed, tape=tape)
        evals = n_samp + n_warm
    dt = time.perf_counter() - t0
    return evals, dt, tape.grad_evals


def cpu_baseline(wl_name: str, cores: int, n_warm=150, n_samp=350):
    """Oracle port on `cores` host cores, one independent chain per core (the reference is a single
    Python thread per chain).  Returns the cpu_baseline object."""
    jobs = [(wl_name, 1000 + i, n_warm, n_samp) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_chain(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_chain, jobs)
    wall = time.perf_counter() - t0
    evals = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return {"value": evals / busy, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} chain(s) x ({n_warm} warm-up + {n_samp} draws) of the same model/sampler settings on the "
                      f"oracle restatement (reference source semantics on a torch-CPU stand-in for MLX); one leapfrog "
                      f"step counted as one grad-eval (the reference spends 2 mx.grad + value traces per step: "
                      f"{sum(r[2] for r in res)} mx.grad calls here); wall {wall:.1f}s"}


# ----------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ess", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": wl["desc"], "chains_per_gpu": args.chains or wl["chains"], "iters_per_step": wl["iters"],
              "sharding": "chains (no data-path collective)", "l2": "flushed between timed steps (256 MiB write)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        W_, K = max(args.warmup, 0), max(args.steps, 1)
        # each step is a bounded sample: one short chain per core
        for _ in range(min(W_, 1)):
            cpu_baseline(args.workload, cores, 20, 30)
        vals, t0 = [], time.perf_counter()
        per = None
        for _ in range(K):
            per = cpu_baseline(args.workload, cores, 30, 70)
            vals.append(per["value"])
            if time.perf_counter() - t0 > 150:
                break
        v = float(np.mean(vals))
        per["value"] = v
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
                "warmup": min(W_, 1), "ms_per_step": 1e3 * (time.perf_counter() - t0) / len(vals), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": per, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import mlx_mcmc_b200 as B
    import mlx_mcmc_b200.core as mx
    from mlx_mcmc_b200 import _cabi, workloads as W
    from mlx_mcmc_b200.diagnostics import compute_ess, ess_geyer, min_ess
    from mlx_mcmc_b200.engine import ChainState, compile_model, launch_hmc, launch_mh

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _cabi.load(build_if_missing=False)

    C = args.chains or wl["chains"]            # weak scaling: per-GPU work fixed
    chain_offset = rank * C
    fn, init, meta = W.ALL_SMALL[wl["model"]](B.ns)
    model = compile_model(fn, init)
    D, ITERS = model.D, wl["iters"]
    st = ChainState(model, model.pack(init, C), wl.get("step_size", 0.0), chain_offset)
    draws = torch.empty((ITERS, C, D), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    seed = 1234

    # untimed: adapt the step size exactly as run() would (reference rule, 300 iterations)
    if wl["method"] == "hmc":
        launch_hmc(st, 300, wl["L"], _cabi.ADAPT_REFERENCE, 0.8, seed, 0)
        st.reset_counters()

    it_count = [300]

    def one_step():
        if wl["method"] == "hmc":
            launch_hmc(st, ITERS, wl["L"], _cabi.ADAPT_NONE, 0.8, seed, it_count[0], draws=draws)
        else:
            launch_mh(st, ITERS, wl["proposal_scale"], seed, it_count[0], draws=draws)
        it_count[0] += ITERS

    evals_per_step = C * ITERS * (wl["L"] if wl["method"] == "hmc" else 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        one_step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    n0 = lib.b2m_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.fill_(1)                                  # evict L2 between timed steps (not timed)
        a.record()
        one_step()
        b.record()
    barrier()
    launches = lib.b2m_launch_count() - n0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(kernel_ms))

    # ---- e2e: the public API call with host buffers, per step
    e2e_steps = max(2, min(args.steps, 5))
    n_warm_e2e = ITERS
    h2d = C * D * 4
    d2h = ITERS * C * D * 4

    def api_call(k):
        if wl["method"] == "hmc":
            return B.hmc(fn, init, num_samples=ITERS, num_warmup=n_warm_e2e, step_size=wl["step_size"],
                         num_leapfrog_steps=wl["L"], key=mx.random.key(k), num_chains=C, chain_offset=chain_offset)
        return B.metropolis_hastings(fn, init, num_samples=ITERS, proposal_scale=wl["proposal_scale"], random_seed=k,
                                     num_chains=C, chain_offset=chain_offset)

    api_call(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        out = api_call(k + 1)
    barrier()
    e2e_s = time.perf_counter() - t0
    clock_info = clocks.stop()
    e2e_evals = e2e_steps * C * ((ITERS + n_warm_e2e) * wl["L"] if wl["method"] == "hmc" else ITERS)

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    value = world * evals_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = world * e2e_evals / e2e_s

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        launch_ms = total_ms / max(launches, 1)
        # algorithmic bytes per launch: draws written + chain state read and written back
        state_bytes = C * (D * 4 + 8 + 16) * 2
        alg_bytes = ITERS * C * D * 4 + state_bytes
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": None, "kernel": "hmc_kernel" if wl["method"] == "hmc" else "mh_kernel",
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                    "note": "observations live in shared memory; the kernel is FP32-issue bound, not HBM bound -- see `issue`"}
        n_obs = getattr(meta, "N", 0)
        issue = {"grad_evals_per_s_per_gpu": value / world, "obs_terms_per_s_per_gpu": value / world * n_obs,
                 "avg_launch_ms": launch_ms}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clock_info, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "call": f"hmc(num_warmup={n_warm_e2e}, num_samples={ITERS}, num_chains={C})" if wl["method"] == "hmc"
                        else f"metropolis_hastings(num_samples={ITERS}, num_chains={C})"},
                "roofline": roofline, "issue": issue}
        if not args.no_ess and wl["method"] == "hmc":
            # min-ESS/s of a full run() through the API (1000 warm-up + 1000 draws), wall clock incl. D2H
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s, rate = B.hmc(fn, init, num_samples=1000, num_warmup=1000, step_size=wl["step_size"], num_leapfrog_steps=wl["L"],
                            key=mx.random.key(99), num_chains=C)
            wall = time.perf_counter() - t0
            line["ess"] = {"min_ess_per_s_geyer": min_ess(s, C, ess_geyer) / wall,
                           "min_ess_per_s_reference_estimator": min_ess(s, C, compute_ess) / wall,
                           "accept_rate": rate, "run": "1000 warm-up + 1000 draws", "wall_s": wall, "chains": C,
                           "note": "ESS summed over a strided subset of 256 chains scaled to all chains; per GPU"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload, 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
