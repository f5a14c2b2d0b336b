"""The tcgen05 encoder attention (csrc/enc_attn_tc.cu) is opt-in (CONE_ATTN_TC=1; the mma.sync kernel is faster on the
benchmark windows, profiles/r02_notes.md §5).  The switch is read once per process, so the tensor-core parity tests are re-run
in a child process with the kernel enabled: dense layers against the oracle, the end-to-end cases and the MAD-768 case."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tc_parity_with_the_tcgen05_attention():
    env = dict(os.environ, CONE_ATTN_TC="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_tc.py"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-30:])
    assert r.returncode == 0, "tensor-core parity tests fail with CONE_ATTN_TC=1:\n" + tail
    assert " passed" in r.stdout
