"""CPU-side checks of the boundary: the library builds, loads and exports every symbol the header declares;
host-side packing logic.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

from cone_b200 import _lib, build
from cone_b200.config import EGO4D, MAD512
from cone_b200.engine import pack_queries
from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict, state_dict_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cone_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cone_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cone_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"


def test_weight_blob_size_matches_state_dict(lib):
    import ctypes as C
    import math
    for cfg in (EGO4D, MAD512):
        d = _lib.ConeDims(cfg.v_feat_dim, cfg.t_feat_dim, cfg.hidden_dim, cfg.nheads, cfg.dim_feedforward,
                          cfg.enc_layers, cfg.dec_layers, cfg.num_queries, cfg.max_v_l, cfg.max_q_l)
        want = sum(math.prod(s) for s in state_dict_shapes(cfg).values())
        assert lib.cone_weights_expected_floats(C.byref(d)) == want
        assert lib.cone_workspace_bytes(C.byref(d), 64, cfg.max_v_l, cfg.max_q_l) > 0
    assert sum(math.prod(s) for s in state_dict_shapes(EGO4D).values()) == 4355589


def test_bad_dims_are_refused(lib):
    import ctypes as C
    d = _lib.ConeDims(250, 768, 256, 8, 1024, 2, 2, 5, 90, 20)  # Dv not a multiple of 16
    assert lib.cone_weights_expected_floats(C.byref(d)) == 0
    assert b"multiples of 16" in lib.cone_last_error()


def test_state_dict_is_deterministic():
    a, b = init_state_dict(EGO4D, 3), init_state_dict(EGO4D, 3)
    assert all((a[k] == b[k]).all() for k in a)
    assert list(a) == list(state_dict_shapes(EGO4D))


def test_pack_queries_groups_by_video_and_keeps_eval_batches():
    cfg = EGO4D.replace(eval_bsz=4)
    ds = make_dataset(cfg, 3, [900, 40, 300], [5, 2, 3], seed=1)
    qs = list(ds.queries)
    qs[1], qs[6] = qs[6], qs[1]  # dataset order not grouped by video
    qb = pack_queries(cfg, [len(v) for v in ds.videos], qs)
    vid = [qs[i].video_idx for i in qb.order]
    assert vid == sorted(vid)
    assert qb.q_first.tolist() == [0, 5, 7, 10]
    assert qb.q_batch.tolist() == [int(i) // 4 for i in qb.order]
    assert qb.n_batches == 3
    assert qb.total_scores == 5 * 900 + 2 * 40 + 3 * 300
    assert qb.tok_len.max().item() <= cfg.max_q_l
    # truncation to max_q_l and zero padding
    j = int(np.argmax([len(qs[i].tokens) for i in qb.order]))
    q = qs[qb.order[j]]
    n = min(len(q.tokens), cfg.max_q_l)
    assert np.array_equal(qb.tokens[j, :n].numpy(), q.tokens[:n])
    assert not qb.tokens[j, n:].any()


def test_bench_algorithmic_flops_match_the_survey():
    """transformer + heads FLOPs per query, 9.72 / 20.4 / 21.05 GFLOP (Ego4D / MAD-512 / MAD-768)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from cone_b200.config import MAD768
    for cfg, want in ((EGO4D, 9.72e9), (MAD512, 20.4e9), (MAD768, 21.05e9)):
        assert abs(bench.algorithmic_flops_per_query(cfg) / want - 1) < 0.01, cfg.name
