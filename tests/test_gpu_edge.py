"""Edge cases of the hot path against the CPU oracle: videos of 1-3 frames and videos shorter than one
window (fewer windows than `topk_window`), queries of one token and of `max_q_l` and more tokens (truncation), a video nobody
asks about, an empty query list, and a workspace so small that the windows of one step are processed in many chunks."""
import dataclasses

import numpy as np
import pytest
import torch

from cone_b200.config import EGO4D
from cone_b200.engine import ConeEngine
from cone_b200.inference import ground_dataset
from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict
from oracle import cone_oracle as O
from helpers import FP32_TOL, TC_TOL, Hatch, assert_close, oracle_window_scores, ranklist_near_tie

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _edge_dataset(cfg):
    ds = make_dataset(cfg, 7, [1, 2, 3, 44, 89, 90, 300], [2, 1, 2, 3, 0, 2, 3], seed=77)
    rng = np.random.default_rng(5)
    qs = list(ds.queries)
    # token-count extremes: one token, exactly max_q_l, more than max_q_l (the reader truncates, dataloader:262)
    for i, n_tok in ((0, 1), (3, cfg.max_q_l), (5, cfg.max_q_l + 7), (len(qs) - 1, 2)):
        qs[i] = dataclasses.replace(qs[i], tokens=rng.standard_normal((n_tok, cfg.t_feat_dim), dtype=np.float32))
    ds.queries[:] = qs
    return ds


def _compare(cfg, sd, ds, res, ora, tol, tag):
    hatch = Hatch(tag, "top-k window list differs from the oracle (near-tie audited)", 0)
    assert set(res) == {q.query_id for q in ds.queries}
    for q in ds.queries:
        r, o = res[q.query_id], ora[q.query_id]
        nw = cfg.num_window(len(ds.videos[q.video_idx]))
        assert sorted(r["ranklist"]) == list(range(nw)), q.query_id
        if r["ranklist"][: cfg.topk_window] != o["ranklist"][: cfg.topk_window]:
            assert ranklist_near_tie(oracle_window_scores(sd, cfg, ds, q), r["ranklist"], o["ranklist"], cfg.topk_window), q.query_id
            hatch.use(q.query_id)
            continue
        assert r["windows"] == o["windows"], q.query_id
        assert len(r["windows"]) == min(nw, cfg.topk_window)
        assert_close(r["pred_spans"], np.stack(o["pred_spans"]), tol, "pred_spans " + q.query_id)
        assert_close(r["prob_fg"], np.stack(o["prob_fg"]), tol, "prob_fg " + q.query_id)
        for mode in ("fusion", "proposal", "matching"):
            assert len(r[mode]) <= cfg.max_after_nms and len(r[mode]) >= 1
    hatch.close(len(ds.queries))


def test_degenerate_videos_and_token_counts_fp32():
    cfg = EGO4D.replace(eval_bsz=4)
    sd = init_state_dict(cfg, 3)
    ds = _edge_dataset(cfg)
    eng = ConeEngine(cfg, sd, device=DEV, precision="fp32", workspace_bytes=1 << 30)
    res = ground_dataset(eng, ds.videos, ds.queries)
    ora = O.eval_pipeline(sd, cfg, ds.videos, ds.queries)
    _compare(cfg, sd, ds, res, ora, FP32_TOL, "edge_fp32")
    # final lists: identical lengths and, where the raw outputs agree to 1e-5, identical top-1 spans within rounding
    for q in ds.queries:
        for mode in ("fusion", "proposal", "matching"):
            assert len(res[q.query_id][mode]) == len(ora[q.query_id][mode]), (q.query_id, mode)


def test_degenerate_videos_and_token_counts_tc():
    cfg = EGO4D.replace(eval_bsz=4)
    sd = init_state_dict(cfg, 3)
    ds = _edge_dataset(cfg)
    eng = ConeEngine(cfg, sd, device=DEV, precision="tc", workspace_bytes=1 << 30)
    res = ground_dataset(eng, ds.videos, ds.queries)
    ora = O.eval_pipeline(sd, cfg, ds.videos, ds.queries)
    _compare(cfg, sd, ds, res, ora, TC_TOL, "edge_tc")


def test_no_queries_and_unasked_videos():
    cfg = EGO4D.replace(eval_bsz=4)
    sd = init_state_dict(cfg, 3)
    eng = ConeEngine(cfg, sd, device=DEV, precision="fp32", workspace_bytes=1 << 30)
    ds = make_dataset(cfg, 3, [120, 200, 95], [0, 0, 0], seed=1)
    assert ground_dataset(eng, ds.videos, ds.queries) == {}
    assert ground_dataset(eng, [], []) == {}
    torch.cuda.synchronize()


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_small_workspace_chunks_are_bit_identical(precision):
    """The windows of a step are processed in chunks sized to the caller's workspace; the chunking must not change a bit."""
    cfg = EGO4D.replace(eval_bsz=8)
    sd = init_state_dict(cfg, 9)
    ds = make_dataset(cfg, 3, [900, 455, 700], 6, seed=13)
    big = ConeEngine(cfg, sd, device=DEV, precision=precision, workspace_bytes=2 << 30)
    want = ground_dataset(big, ds.videos, ds.queries)
    del big
    small = ConeEngine(cfg, sd, device=DEV, precision=precision, workspace_bytes=96 << 20)
    got = ground_dataset(small, ds.videos, ds.queries)
    for q in ds.queries:
        a, b = got[q.query_id], want[q.query_id]
        assert a["ranklist"] == b["ranklist"]
        for k in ("pred_spans", "prob_fg", "match"):
            x, y = np.asarray(a[k]), np.asarray(b[k])
            assert np.array_equal(x, y, equal_nan=True), (q.query_id, k, float(np.nanmax(np.abs(x - y))))
        for mode in ("fusion", "proposal", "matching"):
            assert a[mode] == b[mode]
