"""Window slicing inside the fused encoder tail (TMA gather4 of the residual rows, csrc/enc_tail.cu) against the gathered copy
of the window rows it replaces (`CONE_TAIL_GATHER=0`): the same values travel by a different route, so every output must be
bit-identical.  The switch is read once per process: the gathered-copy run happens in a child process."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = """
import sys, numpy as np
sys.path.insert(0, {root!r})
from cone_b200.config import MAD512
from cone_b200.engine import ConeEngine
from cone_b200.inference import ground_dataset
from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict
cfg = MAD512.replace(eval_bsz=4)
sd = init_state_dict(cfg, 11)
ds = make_dataset(cfg, 3, [700, 333, 61], [3, 2, 2], seed=19)
eng = ConeEngine(cfg, sd, device="cuda:0", precision="tc", workspace_bytes=2 << 30)
res = ground_dataset(eng, ds.videos, ds.queries)
out = {{}}
for q in ds.queries:
    r = res[q.query_id]
    for k in ("pred_spans", "prob_fg", "match"):
        out[q.query_id + "/" + k] = np.asarray(r[k])
    out[q.query_id + "/ranklist"] = np.asarray(r["ranklist"])
np.savez({path!r}, **out)
"""


def _run(path, gather):
    env = dict(os.environ, CONE_TAIL_GATHER="1" if gather else "0")
    r = subprocess.run([sys.executable, "-c", _SCRIPT.format(root=ROOT, path=path)], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    return dict(np.load(path))


@pytest.mark.gpu
def test_tail_gather4_equals_the_gathered_copy_bit_for_bit():
    with tempfile.TemporaryDirectory() as tmp:
        a = _run(os.path.join(tmp, "gather.npz"), True)
        b = _run(os.path.join(tmp, "copy.npz"), False)
    assert set(a) == set(b) and len(a) > 0
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), (k, float(np.nanmax(np.abs(a[k].astype(np.float64) - b[k]))))
