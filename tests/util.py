"""Shared helpers for the tests: golden fixtures, injection arrays, tolerances."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

RUN_FIXTURES = ["hmc_c1", "hmc_c2", "hmc_normal2d", "hmc_halfnormal", "mh_c5", "mh_c1", "nuts_normal1d",
                "nuts_normal2d", "nuts_halfnormal_scale", "nuts_vector", "nuts_c2", "nuts_regression",
                "nuts_regression_sigma"]


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def flat_params(layout, params):
    """dict -> flat float32 vector in layout order"""
    D = sum(n for _, n, _ in layout.values())
    out = np.zeros(D, dtype=np.float32)
    for name, (off, n, _) in layout.items():
        out[off:off + n] = np.asarray(params[name], dtype=np.float32).reshape(-1)
    return out


def tape_normals(tape, layout, n_iter, first_iter=0):
    """[n_iter, 1, D] momentum/proposal draws from a golden tape, iterations first_iter.."""
    D = sum(n for _, n, _ in layout.values())
    z = np.zeros((n_iter, 1, D), dtype=np.float32)
    for rec in tape["normals"]:
        it = rec["it"] - first_iter
        if 0 <= it < n_iter:
            off, n, _ = layout[rec["param"]]
            z[it, 0, off:off + n] = np.asarray(rec["z"], dtype=np.float32).reshape(-1)
    return z


def tape_uniforms(tape, role, n_iter, first_iter=0, fill=0.5):
    u = np.full((n_iter, 1), fill, dtype=np.float32)
    for rec in tape["uniforms"]:
        s = rec["slot"]
        if s[1] == role and len(s) == 2 and 0 <= s[0] - first_iter < n_iter:
            u[s[0] - first_iter, 0] = rec["u"]
    return u


def tape_nuts(tape, layout, n_iter, max_depth, first_iter=0):
    """Injection arrays for b2m_nuts_run from a golden NUTS tape (unused slots hold 0.5)."""
    M = (1 << max_depth) - 1
    inj = {
        "normal": tape_normals(tape, layout, n_iter, first_iter),
        "slice": tape_uniforms(tape, "slice", n_iter, first_iter),
        "dir": np.full((n_iter, 1, max_depth), 0.5, dtype=np.float32),
        "take": np.full((n_iter, 1, max_depth), 0.5, dtype=np.float32),
        "merge": np.full((n_iter, 1, max_depth, M), 0.5, dtype=np.float32),
    }
    for rec in tape["uniforms"]:
        s = rec["slot"]
        it = s[0] - first_iter
        if not 0 <= it < n_iter:
            continue
        if s[1] == "dir":
            inj["dir"][it, 0, s[2]] = rec["u"]
        elif s[1] == "take":
            inj["take"][it, 0, s[2]] = rec["u"]
        elif s[1] == "merge":
            inj["merge"][it, 0, s[2], s[3]] = rec["u"]
    return inj


def rel_err(a, b):
    """norm-wise relative error max|a-b| / max(|b|, tiny)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
