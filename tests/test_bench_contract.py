"""bench.py prints ONE JSON line with the keys the driver reads.  The reference arm (CPU oracle) runs anywhere; the
B200 arm needs a GPU.  Both on the smoke-sized workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1", "--cpu-sample", "1")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "experts/sec (optimise+predict)" and d["unit"] == "experts/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run("--workload", "tiny", "--steps", "2", "--warmup", "3", "--experts-per-step", "16", "--cpu-sample", "1")
    assert BASE <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["gpu_launches"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f64"
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == d["unit"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 30 < r["peak"] < 45
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]
