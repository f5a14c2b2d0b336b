"""Two GPUs (skipped on a single-GPU box): the sharded evaluation driver under torchrun + NCCL must print exactly the
tables of the single-process run — videos are independent units, the only exchange is the all-reduce of the counters."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, out, flavour):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + world), "-m", "cone_b200.tools.sharded_eval", "--videos", "10",
           "--queries", "4", "--flavour", flavour, "--out", out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.load(open(out))


@pytest.mark.parametrize("flavour", ["mad", "ego4d"])
def test_sharded_eval_equals_single_process(tmp_path, flavour):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = _run(1, str(tmp_path / "one.json"), flavour)
    two = _run(2, str(tmp_path / "two.json"), flavour)
    assert one["n_queries"] == two["n_queries"] == 40
    assert one["window_recall"] == two["window_recall"]
    for k in ("fusion", "proposal", "matching"):
        if flavour == "mad":
            assert one[k] == two[k], k
        else:  # hit counts are integers; the mean of the top-1 IoUs is summed in a different order across ranks
            assert one[k]["recall"] == two[k]["recall"], k
            assert one[k]["mIoU"] == pytest.approx(two[k]["mIoU"], abs=1e-12), k
