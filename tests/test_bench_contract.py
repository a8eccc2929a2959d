"""The CPU-runnable part of the bench.py contract: the reference arm prints ONE JSON line with the agreed keys
(the GPU arm is exercised on the B200 box; its line has the same base keys plus roofline / clocks / gpu_launches)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-mf",
                        "--frames-per-gpu", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert (BASE_KEYS - {"cpu_baseline"}) | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert "impl" not in d and d["value"] > 0 and d["gpu_launches"] > 0 and d["dtype"] == "f32"
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 6 * 4 * 512 * 432 * 64 and d["e2e"]["d2h_bytes_per_step"] == 4
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    # round 2: the headline is the module path, the fused entry sits beside it; the roofline names its compute limiters
    assert d["value_and_grad"]["value"] > 0 and d["value_and_grad"]["ms_per_step"] > 0
    assert rf["traffic"] is None or rf["traffic"] > 0.5 * rf["algorithmic_bytes"]
    assert rf["xu_frac"] is None or 0 < rf["xu_frac"] < 1
    assert rf["issue_frac"] is None or 0 < rf["issue_frac"] < 1
    lo = d["e2e"]["loader_inputs_only"]
    assert lo["value"] >= d["e2e"]["value"] * 0.9 and lo["h2d_bytes_per_step"] == 2 * 4 * 512 * 432 * 64
    assert d["strong"] is None and "forward() + backward()" in d["config"]["workload"]
