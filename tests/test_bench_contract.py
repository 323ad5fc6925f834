"""bench.py's driver-facing contract, as far as it can be checked without a GPU: the reference arm runs the oracle's
port of the reference CPU path and prints one JSON line with the agreed keys; our arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("megapixels/sec raw->sRGB")
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_do_nothing():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
