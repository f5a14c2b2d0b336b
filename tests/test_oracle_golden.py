"""Pins the CPU oracle against outputs of the unmodified reference (tests/golden/, made by
`python -m oracle.make_golden` from /root/reference).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from cone_b200.config import EGO4D, MAD512
from cone_b200.weights import init_state_dict
from oracle import cone_oracle as O
from helpers import E2E_CASES, FP32_TOL, GOLDEN, assert_close, assert_match_close, dense_case, load_e2e


@pytest.fixture(scope="module")
def host_ops():
    with open(os.path.join(GOLDEN, "host_ops.json")) as f:
        return json.load(f)


def test_span_cxw_to_xx_docstring_vector(host_ops):
    # the one intact known-answer vector in the reference: cone/span_utils.py:30-33
    case = host_ops["span_cxw_to_xx"]
    got = O.span_cxw_to_xx(torch.Tensor(case["inp"])).tolist()
    assert got == case["out"]
    assert np.allclose(got, [[0.0, 1.0], [0.2, 0.4]], atol=1e-7)


def test_temporal_nms_exact(host_ops):
    for c in host_ops["temporal_nms"]:
        got = O.temporal_nms([list(x) for x in c["inp"]], c["thd"], c["max_after"])
        assert got == c["out"], (c["thd"], c["max_after"], len(c["inp"]))


def test_temporal_nms_survey_vector():
    # SURVEY.md §8c derived-by-probe vector; IoU exactly 0.5 is NOT suppressed (strict >)
    out = O.temporal_nms([[0, 10, .9], [1, 11, .8], [20, 30, .7], [0, 20, .6], [5, 15, .5]], 0.5, 5)
    assert out == [[0, 10, .9], [20, 30, .7], [0, 20, .6], [5, 15, .5]]


def test_normalize_score_exact(host_ops):
    for c in host_ops["normalize_score"]:
        assert O.normalize_score(list(c["inp"])) == c["out"]


@pytest.mark.parametrize("name,cfg,wseed", [("dense_ego4d", EGO4D, 3), ("dense_mad512", MAD512, 4)])
def test_forward_and_matching_dense(name, cfg, wseed):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = init_state_dict(cfg, wseed)
    vid, vm, txt, tm, cls = dense_case(cfg, 100 + wseed)
    with torch.no_grad():
        out = O.cone_forward(sd, txt, tm, vid, vm)
        # operator boundary: matching is fed the REFERENCE's spans, so floor/ceil bounds are identical
        match = O.clip_matching(sd, cls, vid, vm, torch.from_numpy(g["pred_spans"]))
        adapted = O.adapter(sd, vid[0])
    assert_close(out["pred_logits"], g["pred_logits"], FP32_TOL, "pred_logits")
    assert_close(out["pred_spans"], g["pred_spans"], FP32_TOL, "pred_spans")
    assert_close(out["saliency_scores"], g["saliency"], FP32_TOL, "saliency")
    assert_close(out["aux_outputs"][0]["pred_spans"], g["aux_spans"], FP32_TOL, "aux spans")
    assert_close(match, g["match"], FP32_TOL, "match")
    assert_close(adapted, g["adapted0"], FP32_TOL, "adapter")


@pytest.mark.parametrize("name", list(E2E_CASES))
def test_eval_pipeline_vs_reference(name):
    cfg, sd, ds, arrays, lists = load_e2e(name)
    res = O.eval_pipeline(sd, cfg, ds.videos, ds.queries)
    for q in ds.queries:
        r, g = res[q.query_id], lists[q.query_id]
        # integer outputs: bit-exact
        assert r["ranklist"] == g["ranklist"], q.query_id
        assert len(r["windows"]) == min(cfg.topk_window, len(g["ranklist"]))
        assert_close(np.stack(r["pred_spans"]), arrays[f"{q.query_id}/pred_spans"], FP32_TOL, "spans")
        assert_close(np.stack(r["prob_fg"]), arrays[f"{q.query_id}/prob_fg"], FP32_TOL, "prob")
        assert_match_close(np.stack(r["match"]), arrays[f"{q.query_id}/match"], arrays[f"{q.query_id}/pred_spans"],
                           [n for _, n in r["windows"]], FP32_TOL)
        # operator boundary for stages A10-A13: the reference's raw outputs in, exact lists out
        pp = O.postprocess_query(cfg, r["windows"], arrays[f"{q.query_id}/pred_spans"],
                                 arrays[f"{q.query_id}/prob_fg"], arrays[f"{q.query_id}/match"])
        for k in ("rows", "fusion", "proposal", "matching"):
            assert pp[k] == g[k], (q.query_id, k)


@pytest.mark.parametrize("name", list(E2E_CASES))
def test_window_ranker_operator_boundary(name):
    # fed the reference's own frame scores the rank-list must be identical (ties -> lower index)
    cfg, sd, ds, arrays, lists = load_e2e(name)
    n = 0
    for q in ds.queries:
        key = f"{q.query_id}/frame_score"
        if key in arrays:
            assert O.window_ranklist(torch.from_numpy(arrays[key]), cfg.max_v_l) == lists[q.query_id]["ranklist"]
            n += 1
    assert n >= 2


def test_recall_matches_reference_metric():
    cfg, sd, ds, arrays, lists = load_e2e("e2e_ego4d")
    pred = {q.query_id: lists[q.query_id]["fusion"] for q in ds.queries}
    gt = {q.query_id: list(q.timestamps) for q in ds.queries}
    rec = O.recall_at_k_iou(pred, gt, thresholds=(0.1, 0.3, 0.5), topk=(1, 5, 10, 50, 100))
    assert np.allclose(rec, np.asarray(lists["_metrics"]["fusion_recall"]), atol=1e-6)
