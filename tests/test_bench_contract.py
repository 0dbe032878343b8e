"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, cwd=REPO, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                     # stdout carries the JSON line and nothing else
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    d = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--num-envs", "256")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"] == "env-steps/sec incl. obs" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "lockstep steps" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["timed_lockstep_steps"] >= 512          # never a run so short that it times env creation


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run("--steps", "64", "--warmup", "3", "--cpu-seconds", "1")
    assert BASE_KEYS | {"clocks", "roofline"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 64 and d["warmup"] == 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["value"] > 1e8 and d["gpu_launches"] >= 1 and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 and e["d2h_bytes_per_step"] == 4096 * 372
    assert e["value"] < d["value"]                    # host buffers and PCIe inside the timed region
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["single_thread"]["cores"] == 1
