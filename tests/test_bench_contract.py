"""CPU: the shape of bench.py's one JSON line -- the reference arm run for real on a small workload, and the last
committed device line under profiles/ (the device arm itself needs a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _check_common(line):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["metric"] == "leapfrog_grad_evals_per_sec" and line["unit"] == "grad-evals/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert isinstance(line["config"], dict) and "workload" in line["config"] and "model" not in line["config"]
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])


def test_reference_arm_prints_exactly_one_json_line():
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    _check_common(line)
    assert line["impl"] == "reference"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "NUTS transition" in cb["sample"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 == line["e2e"]["d2h_bytes_per_step"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, PYTHONPATH=ROOT, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_device_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r01_bench_default_v9_skip.json")) as f:
        line = json.loads(f.read())
    _check_common(line)
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["scaling"] == "weak"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    cb = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] != line["value"]          # measured separately, not a copy of the device-timed value
