"""bench.py contract on the CPU side: the reference arm prints ONE JSON line with the agreed keys (timed on the unmodified reference built into
oracle/_ref), and the engine arm refuses to run without a CUDA device (no CPU fallback on the product path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--kkt-n", "4000", "--kkt-m", "20", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "systems/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("KKT systems/sec")
    assert d["steps"] == 1 and d["warmup"] == 0 and d["steps_requested"] == 3          # one full-size system, whatever was asked for
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) <= 1e-9 * d["value"]
    assert d["config"]["sampled"] is False and "n=4000 m=20" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] >= 1 and "nothing extrapolated" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "systems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline_optimised"]["kind"] == "port" and d["cpu_baseline_optimised"]["value"] > 0


def test_engine_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert p.returncode != 0
    assert not any(ln.startswith("{") for ln in p.stdout.splitlines())              # no number without the CUDA path
