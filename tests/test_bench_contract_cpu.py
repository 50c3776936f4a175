"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the driver reads
(here on a small sample so that it takes seconds), and the B200 arm refuses to run without a CUDA device instead of
falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], cwd=ROOT, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--size", "32", "--levels", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mcells/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert "Mcells/s" in d["metric"] and "3D Poisson" in d["config"]["workload"]
    assert d["value"] > 0 and abs(d["ms_per_step"] - 32 ** 3 / d["value"] / 1e3) < 1e-6 * d["ms_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "Mcells/s"
    assert "32x32x32" in cb["sample"] and "UNMODIFIED reference" in cb["sample"] and cb["same_grid_as_workload"]
    e = d["e2e"]
    assert e == {"value": d["value"], "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present")
def test_b200_arm_refuses_cpu():
    r = run_bench("--steps", "1", "--warmup", "1", "--no_cpu_baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
