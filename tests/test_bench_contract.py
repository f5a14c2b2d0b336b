"""bench.py's contract on a box without a GPU: the product arm fails loudly (no CPU path, no JSON line), the reference arm
(`--impl reference`: the unmodified reference's eval_epoch on the host cores) prints exactly ONE JSON line on stdout with the
keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU path" in (r.stdout + r.stderr)
    assert not any(line.strip().startswith("{") for line in r.stdout.splitlines())


def test_reference_arm_prints_one_json_line():
    from oracle import ref_harness as RH
    if not RH.reference_available():
        pytest.skip("reference files not placed (oracle/vendor_ref.py)")
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-queries", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grounding_queries_per_sec" and d["unit"] == "queries/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
