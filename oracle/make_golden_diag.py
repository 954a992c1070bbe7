"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/diagnostics.json (needs /root/reference).

Executes the reference's own `compute_ess` bodies (cut out of the example scripts with `ast`, because the scripts
import matplotlib and run at import) and the reference's `MCMC.summary` on seeded series, asserts the restatement in
oracle/refport/diagnostics.py reproduces them exactly, and stores series + answers as a fixture.
"""
import ast
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("B2M_REFERENCE", "/root/reference")

from oracle.ns import ns as _ns  # noqa: E402,F401  (puts the mlx.core stand-in on sys.path)
from oracle.refport import diagnostics as port  # noqa: E402

sys.path.insert(0, REFERENCE)
import mlx_mcmc as ref  # noqa: E402


def reference_function(path, name):
    """compile one function definition of a reference script without running the script"""
    tree = ast.parse(open(os.path.join(REFERENCE, path)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            scope = {"np": np}
            exec(compile(mod, path, "exec"), scope)
            return scope[name]
    raise KeyError(name)


def series(kind, n, seed):
    rng = np.random.default_rng(seed)
    e = rng.standard_normal(n)
    if kind == "iid":
        x = e
    elif kind == "ar1":            # slowly mixing, like random-walk Metropolis
        x = np.zeros(n)
        for i in range(1, n):
            x[i] = 0.9 * x[i - 1] + e[i]
    elif kind == "antithetic":     # HMC near half a period (the estimator goes negative here)
        x = np.zeros(n)
        for i in range(1, n):
            x[i] = -0.7 * x[i - 1] + e[i]
    elif kind == "sticky":         # rejected proposals repeat values
        x = np.repeat(e[: n // 5 + 1], 5)[:n]
    else:                          # constant
        x = np.full(n, 1.25)
    return (3.0 + 2.0 * x).astype(np.float32)


def main():
    ess06 = reference_function("examples/06_nuts_comparison.py", "compute_ess")
    ess02 = reference_function("examples/02_hmc_comparison.py", "compute_ess")
    rows = []
    for kind in ("iid", "ar1", "antithetic", "sticky", "constant"):
        for n in (8, 57, 400, 1500):
            x = series(kind, n, seed=n + len(kind))
            a, b = ess06(x), (ess02(x) if kind != "constant" else None)
            assert a == port.compute_ess(x), (kind, n)
            if b is not None:
                assert b == port.compute_ess_example02(x), (kind, n)
            rows.append({"kind": kind, "n": n, "x": x.astype(np.float64).tolist(), "ess06": float(a),
                         "ess02": None if b is None else float(b)})
    m = ref.MCMC(lambda p: 0.0)
    m.samples = {"mu": series("ar1", 700, 1), "beta": np.stack([series("iid", 300, 2), series("ar1", 300, 3)], axis=1)}
    summ = m.summary(0.9)
    assert summ == port.summary(m.samples, 0.9)
    out = {"ess": rows, "summary": {"samples": {k: np.asarray(v, dtype=np.float64).tolist() for k, v in m.samples.items()},
                                    "credible_interval": 0.9, "table": summ}}
    path = os.path.join(ROOT, "tests", "golden", "diagnostics.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, "written:", len(rows), "series")


if __name__ == "__main__":
    main()
