"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.json (run in the build container).

    python oracle/make_golden.py            # needs /root/reference (read-only) -- not available on the GPU box
    python oracle/make_golden.py --check hmc_c2 nuts_c2 ...   # re-run the reference and compare with the committed files

What it does
  1. imports the UNMODIFIED reference package from /root/reference on the mlx.core stand-in
     (oracle/mlx_shim) and runs its hmc / nuts / MCMC.run(metropolis) on the parity models;
  2. runs the restatement (oracle/refport) with the same stand-in keys and ASSERTS bit-equal draws and
     acceptance rates -- this is what pins the oracle to the reference;
  3. evaluates log p and its gradient at fixed points with the reference's own distribution classes
     (float32, torch autograd standing in for mx.grad) and with the restatement in float64 (arbiter);
  4. writes everything, including the slot-addressed random draws of each run, as small JSON fixtures.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("B2M_REFERENCE", "/root/reference")

from oracle.ns import Tape, ns as port_ns, samplers, value_and_grad  # noqa: E402  (puts the stand-in on sys.path)

sys.path.insert(0, REFERENCE)
import mlx.core as mx  # noqa: E402
import mlx_mcmc as ref  # noqa: E402

from mlx_mcmc_b200 import workloads as W  # noqa: E402

ref_ns = SimpleNamespace(mx=mx, Normal=ref.Normal, HalfNormal=ref.HalfNormal, Beta=ref.Beta, Gamma=ref.Gamma,
                         Exponential=ref.Exponential, Categorical=ref.Categorical)
OUT = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def tolist(x):
    return np.asarray(x, dtype=np.float64).tolist()


def tape_json(tape: Tape):
    return {
        "normals": [{"it": k[0], "param": k[1], "z": tolist(v)} for k, v in tape.normals.items()],
        "uniforms": [{"slot": list(k), "u": float(v)} for k, v in tape.uniforms.items()],
        "iters": json.loads(json.dumps(tape.iters, default=lambda o: tolist(o))),
        "grad_evals": tape.grad_evals, "value_evals": tape.value_evals, "leapfrogs": tape.leapfrogs,
    }


POINTS = {
    "c1_normal": [{"mu": 0.0, "sigma": 1.0}, {"mu": 4.8, "sigma": 1.85}, {"mu": 5.5, "sigma": 0.7}, {"mu": -2.0, "sigma": 9.0},
                  {"mu": 1.0, "sigma": -0.5}],
    "c2_event_rate": [{"rate": 2.0}, {"rate": 3.44}, {"rate": 0.3}, {"rate": 11.0}, {"rate": -1.0}],
    "c5_ab_test": [{"p_A": 0.1, "p_B": 0.1}, {"p_A": 0.1158, "p_B": 0.1527}, {"p_A": 0.5, "p_B": 0.9}, {"p_A": -0.1, "p_B": 0.2},
                   {"p_A": 0.3, "p_B": 1.5}],
    "t_normal_1d": [{"mu": 0.0}, {"mu": 5.0}, {"mu": -3.25}],
    "t_normal_2d": [{"mu1": 0.0, "mu2": 0.0}, {"mu1": 0.7, "mu2": 4.0}],
    "t_halfnormal_scale": [{"sigma": 1.0}, {"sigma": 0.3}, {"sigma": 6.0}, {"sigma": -0.4}],
    "t_halfnormal": [{"sigma": 1.0}, {"sigma": 3.0}, {"sigma": -1.0}],
    "t_vector_normal": [{"x": [0.0, 0.0, 0.0]}, {"x": [0.5, -1.0, 2.5]}],
    "t_regression_small": [{"beta": [0.0, 0.0, 0.0]}, {"beta": [0.3, -1.2, 0.8]}, {"beta": [2.0, 2.0, -2.0]}],
    "t_regression_sigma": [{"beta": [0.0, 0.0, 0.0, 0.0], "sigma": 1.0}, {"beta": [0.5, -0.5, 1.0, 0.1], "sigma": 0.7},
                           {"beta": [1.0, 1.0, 1.0, 1.0], "sigma": 3.0}],
}

RUNS = [
    # (fixture name, model, method, kwargs)
    ("hmc_c1", "c1_normal", "hmc", dict(num_samples=40, num_warmup=40, step_size=0.01, num_leapfrog_steps=10, target_accept=0.8, seed=42)),
    ("hmc_c2", "c2_event_rate", "hmc", dict(num_samples=60, num_warmup=30, step_size=0.1, num_leapfrog_steps=10, target_accept=0.8, seed=0)),
    ("hmc_normal2d", "t_normal_2d", "hmc", dict(num_samples=60, num_warmup=40, step_size=0.3, num_leapfrog_steps=8, target_accept=0.8, seed=5)),
    ("hmc_halfnormal", "t_halfnormal", "hmc", dict(num_samples=60, num_warmup=40, step_size=0.4, num_leapfrog_steps=5, target_accept=0.8, seed=9)),
    ("mh_c5", "c5_ab_test", "metropolis", dict(num_samples=150, num_warmup=50, proposal_scale=0.02, seed=3)),
    ("mh_c1", "c1_normal", "metropolis", dict(num_samples=100, num_warmup=50, proposal_scale=0.3, seed=42)),
    ("nuts_normal1d", "t_normal_1d", "nuts", dict(num_samples=60, num_warmup=60, step_size=0.5, max_tree_depth=6, seed=42)),
    ("nuts_normal2d", "t_normal_2d", "nuts", dict(num_samples=60, num_warmup=60, step_size=0.15, max_tree_depth=6, seed=123)),
    ("nuts_halfnormal_scale", "t_halfnormal_scale", "nuts", dict(num_samples=60, num_warmup=60, step_size=0.1, max_tree_depth=6, seed=456)),
    ("nuts_vector", "t_vector_normal", "nuts", dict(num_samples=50, num_warmup=50, step_size=0.2, max_tree_depth=6, seed=7)),
    ("nuts_regression", "t_regression_small", "nuts", dict(num_samples=40, num_warmup=40, step_size=0.05, max_tree_depth=6, seed=21)),
    ("nuts_regression_sigma", "t_regression_sigma", "nuts", dict(num_samples=30, num_warmup=40, step_size=0.05, max_tree_depth=6, seed=22)),
    ("nuts_c2", "c2_event_rate", "nuts", dict(num_samples=40, num_warmup=40, step_size=0.1, max_tree_depth=5, seed=11)),
    # 64 coefficients x 2048 observations: the decision-level fixture of the tensor-core path (round 2)
    ("nuts_regression_mid", "t_regression_mid", "nuts", dict(num_samples=8, num_warmup=8, step_size=0.002, max_tree_depth=5, seed=31)),
]


def golden_points():
    out = {}
    for name, pts in POINTS.items():
        f_ref, _, _ = W.ALL_SMALL[name](ref_ns)
        f_port, _, _ = W.ALL_SMALL[name](port_ns)
        rows = []
        for p in pts:
            names = list(p)
            vals = [mx.array(np.asarray(p[n], dtype=np.float64)) for n in names]
            fn = lambda *a: f_ref(dict(zip(names, a)))  # noqa: E731
            lp32 = float(fn(*vals))
            g32 = mx.grad(fn, argnums=list(range(len(names))))(*vals)
            lp_p32, g_p32 = value_and_grad(f_port, p, "float32")
            assert np.array_equal(np.float32(lp32), np.float32(lp_p32), equal_nan=True), (name, p)
            for n, g in zip(names, g32):
                assert np.array_equal(np.asarray(g, dtype=np.float32), g_p32[n].astype(np.float32), equal_nan=True), (name, p, n)
            lp64, g64 = value_and_grad(f_port, p, "float64")
            rows.append({"params": p, "logp32": lp32, "grad32": {n: tolist(g) for n, g in zip(names, g32)},
                         "logp64": float(lp64), "grad64": {n: tolist(g) for n, g in g64.items()}})
        out[name] = rows
    return out


def golden_run(model, method, kw):
    kw = dict(kw)
    seed = kw.pop("seed")
    f_ref, init, _ = W.ALL_SMALL[model](ref_ns)
    f_port, _, _ = W.ALL_SMALL[model](port_ns)
    tape = Tape()
    if method == "hmc":
        s_ref, a_ref = quiet(ref.hmc, f_ref, init, key=mx.random.key(seed), **kw)
        s_port, a_port, eps = samplers.hmc_port(f_port, init, key=mx.random.key(seed), tape=tape, **kw)
    elif method == "nuts":
        s_ref, a_ref = quiet(ref.nuts, f_ref, init, key=mx.random.key(seed), **kw)
        s_port, a_port, eps = samplers.nuts_port(f_port, init, key=mx.random.key(seed), tape=tape, **kw)
    else:
        m = ref.MCMC(f_ref)
        s_ref = quiet(m.run, init, method="metropolis", random_seed=seed, verbose=False, **kw)
        a_ref = m.acceptance_rate
        s_port, a_port = samplers.run_port(f_port, init, method="metropolis", random_seed=seed, tape=tape, **kw)
        # the tape holds the sampling phase; also record where warm-up ended so the CUDA path can start there
        warm, _ = samplers.metropolis_port(f_port, init, num_samples=kw["num_warmup"], proposal_scale=kw["proposal_scale"],
                                           random_seed=seed)
        eps = None
        tape.iters.insert(0, {"start": {k: float(v[-1]) for k, v in warm.items()}})
    for k in s_ref:
        a, b = np.asarray(s_ref[k], dtype=np.float32), np.asarray(s_port[k], dtype=np.float32)
        assert np.array_equal(a, b), f"restatement differs from the reference: {model}/{method}/{k}"
    assert a_ref == a_port, (model, method, a_ref, a_port)
    return {"model": model, "method": method, "kwargs": kw, "seed": seed, "accept_rate": a_ref, "final_step_size": eps,
            "draws": {k: tolist(v) for k, v in s_ref.items()}, "tape": tape_json(tape)}


def check(names):
    """Re-run the unmodified reference (and the restatement next to it) for the named run fixtures and compare with
    the committed JSON: draws, acceptance rate and the recorded random draws must be identical."""
    runs = {n: (m, meth, kw) for n, m, meth, kw in RUNS}
    for name in names:
        g = json.loads(json.dumps(golden_run(*runs[name])))
        with open(os.path.join(OUT, f"{name}.json")) as f:
            have = json.load(f)
        for key in ("draws", "accept_rate", "final_step_size", "kwargs", "seed"):
            assert g[key] == have[key], f"{name}: committed fixture differs from the reference in {key!r}"
        assert g["tape"]["normals"] == have["tape"]["normals"] and g["tape"]["uniforms"] == have["tape"]["uniforms"], name
        print(f"{name}: committed fixture reproduces")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--check":
        return check(sys.argv[2:])
    only = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None   # (re)write the named run fixtures only
    os.makedirs(OUT, exist_ok=True)
    if only is None:
        with open(os.path.join(OUT, "logp_grad.json"), "w") as f:
            json.dump(golden_points(), f)
        print("logp_grad.json written")
    for name, model, method, kw in RUNS:
        if only is not None and name not in only:
            continue
        g = golden_run(model, method, kw)
        with open(os.path.join(OUT, f"{name}.json"), "w") as f:
            json.dump(g, f)
        print(f"{name}.json written  accept={g['accept_rate']:.3f}")


if __name__ == "__main__":
    main()
