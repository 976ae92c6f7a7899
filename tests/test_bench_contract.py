"""bench.py's JSON contract, checked where it can run without a GPU (the reference arm) and on the GPU (native arm,
tiny config): one line on stdout, the keys the driver reads, the reference arm's e2e with zero copy bytes."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "cpu_baseline"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout  # ONE JSON line on stdout, everything else goes to stderr
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--config", "0", "--steps", "1", "--warmup", "1", "--cpu-budget", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "input triangles voxelized/sec" and d["unit"] == "triangles/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


@pytest.mark.gpu
def test_native_arm_line():
    d = _run("--config", "0", "--steps", "2", "--warmup", "3", "--cpu-budget", "1")
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d) and "impl" not in d
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["dtype"] and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] == "port"


def test_committed_headline_profile_is_a_contract_line():
    """profiles/ holds the bench line the round's numbers are quoted from: it must be a full native-arm line of config 2."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_bench.json")))  # one headline line per round from round 2 on
    assert files
    d = json.load(open(files[-1]))
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["config"]["workload"].startswith("config2") and d["gpu_launches"] > 0
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["parity_check"]["ok"] and d["issue"]["steps_rebuilt"] == 0 and r["issue_roofline"]["kernels"]
    assert 0 < d["e2e"]["value"] < d["value"] and d["e2e"]["d2h_bytes_per_step"] < d["e2e"]["full_layout"]["d2h_bytes_per_step"]
    assert d["steps"] * d["config"]["totals_rank0"]["triangles"] / (d["ms_per_step"] * d["steps"] * 1e-3) == pytest.approx(d["value"], rel=1e-6)
