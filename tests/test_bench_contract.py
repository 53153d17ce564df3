"""CPU checks of bench.py's contract: the reference arm (`--impl reference`, the CPU port of the reference
algorithm, the only bench leg that may run without a GPU) prints ONE JSON line with the keys the driver
reads, for each workload; the default arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("workload,n,batch", [("c64", 256, 64), ("ordered", 2048, 8), ("f128", 64, 16)])
def test_reference_arm_prints_one_json_line(workload, n, batch):
    r = run_bench("--impl", "reference", "--workload", workload, "--n", str(n), "--batch", str(batch), "--steps", "3", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "transforms/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 3 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "transforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["n"] == n and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "256", "--batch", "8",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
