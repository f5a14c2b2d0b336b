"""BASELINE.json's configurations at (or near) their full sizes.  The CPU oracle cannot cover these in seconds, so the
checks are: the oracle on a small subsample of the same inputs, plus size-independent properties of the path
(determinism, rank-list = permutation sorted by the stored window scores, top-k windows = rank-list prefix, NMS
invariants, agreement of the two precision modes)."""
import numpy as np
import pytest
import torch

from cone_b200.config import EGO4D, MAD768
from cone_b200.engine import ConeEngine
from cone_b200.inference import ground_dataset, output_to_host, run_step, stage_step
from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict
from oracle import cone_oracle as O
from helpers import FP32_TOL, TC_TOL, Hatch, assert_close, oracle_window_scores, ranklist_near_tie

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nms_invariants(cfg, rows, col):
    assert len(rows) <= cfg.max_after_nms
    assert all(rows[i][col] >= rows[i + 1][col] for i in range(len(rows) - 1))
    for i in range(len(rows)):
        for j in range(i + 1, len(rows)):
            assert not O.temporal_iou_hull(rows[i], rows[j]) > cfg.nms_thd


def test_mad768_full_step_properties_and_oracle_subsample():
    """configs[2]: one MAD-shaped movie (45 000 frames x 768-d) with 640 queries, top-30 windows of 125 frames."""
    cfg = MAD768
    sd = init_state_dict(cfg, 5)
    ds = make_dataset(cfg, 1, [45000], 640, seed=11)
    eng = ConeEngine(cfg, sd, device=DEV, precision="fp32", workspace_bytes=40 << 30)
    step = stage_step(cfg, ds.videos, ds.queries, [0])
    out1 = run_step(eng, step, want_rows=True)
    out2 = run_step(eng, step, want_rows=True)
    # determinism: two passes over the same inputs are bit-identical
    for name in ("ranklist", "win_start", "win_len", "pred_spans", "prob_fg", "match", "nms", "nms_count"):
        assert torch.equal(getattr(out1, name), getattr(out2, name)), name
    res = output_to_host(cfg, step, out1)
    L = len(ds.videos[0])
    nw = cfg.num_window(L)
    # window scores recomputed from the device's own frame scores: the rank-list must be the stable descending order
    qb = step.qb.to(DEV)
    ctx, _ = eng.video_prepare(step.frames.to(DEV))
    cls_norm = eng.l2_normalize(qb.cls, 1e-5)
    scores, offs = eng.frame_scores(ctx, qb, cls_norm)
    rl, ws = eng.window_ranklist(scores, offs, qb.q_video_len, ranklist_stride=nw, want_scores=True)
    assert torch.equal(rl, out1.ranklist)
    ws_h, rl_h = ws.cpu().numpy(), rl.cpu().numpy()
    for j in range(0, 640, 37):
        order = sorted(range(nw), key=lambda i: (-ws_h[j, i], i))
        assert order == rl_h[j].tolist()
    fs = scores.cpu()
    for j in (0, 123, 639):  # and the window maxima themselves against the oracle's window loop (inference.py:290-295)
        want = O.window_scores(fs[j * L:(j + 1) * L], cfg.max_v_l)
        assert np.array_equal(want.numpy(), ws_h[j])
    for q in ds.queries[::16]:
        r = res[q.query_id]
        assert sorted(r["ranklist"]) == list(range(nw))
        bounds = [cfg.window_bounds(w, L) for w in r["ranklist"][: cfg.topk_window]]
        assert r["windows"] == [(s, e - s) for s, e in bounds]  # top-k windows = rank-list prefix (dataloader:146-149)
        assert np.isfinite(r["pred_spans"]).all() and np.isfinite(r["prob_fg"]).all() and np.isfinite(r["match"]).all()
        assert (r["pred_spans"] > 0).all() and (r["pred_spans"] < 1).all()
        for mode, col in (("fusion", 4), ("proposal", 2), ("matching", 3)):
            _nms_invariants(cfg, r[mode], col)
    # oracle on 6 of the 640 queries (same movie, same weights)
    sub = ds.queries[:3] + ds.queries[-3:]
    ora = O.eval_pipeline(sd, cfg, ds.videos, sub)
    hatch = Hatch("mad768_full_step", "top-k window list differs from the oracle (near-tie audited)", 0)
    for q in sub:
        r, o = res[q.query_id], ora[q.query_id]
        if r["ranklist"][: cfg.topk_window] != o["ranklist"][: cfg.topk_window]:
            assert ranklist_near_tie(oracle_window_scores(sd, cfg, ds, q), r["ranklist"], o["ranklist"], cfg.topk_window), q.query_id
            hatch.use(q.query_id)
            continue
        assert_close(r["pred_spans"], np.stack(o["pred_spans"]), FP32_TOL, "pred_spans")
        assert_close(r["prob_fg"], np.stack(o["prob_fg"]), FP32_TOL, "prob_fg")
    hatch.close(len(sub))
    # reduced-precision mode on the same step: window lists identical to the fp32 path (the pre-filter is fp32 in
    # both), all 288 000 values against the fp32 PATH OF THIS REPO (an internal consistency check — the comparison with
    # the oracle on this shape is tests/test_gpu_tc.py::test_tc_vs_oracle_on_the_benchmark_config), identical R@K
    eng_tc = ConeEngine(cfg, sd, device=DEV, precision="tc", workspace_bytes=40 << 30)
    out_tc = run_step(eng_tc, step, want_rows=True)
    assert torch.equal(out_tc.ranklist, out1.ranklist) and torch.equal(out_tc.win_start, out1.win_start)
    assert torch.equal(out_tc.win_len, out1.win_len)
    valid = (out1.win_len > 0)[:, :, None].expand_as(out1.prob_fg)
    worst, n_over, n_all = 0.0, 0, 0
    for a, b in ((out_tc.pred_spans, out1.pred_spans), (out_tc.prob_fg, out1.prob_fg)):
        d = (a - b).abs()
        d = d[valid.unsqueeze(-1).expand_as(d) if d.dim() == 4 else valid]
        worst = max(worst, float(d.max()))
        n_over += int((d > TC_TOL).sum())
        n_all += d.numel()
    print(f"[tc-vs-fp32] mad768 640 queries: max |tc - fp32| over {n_all} spans / probabilities {worst:.3e}, "
          f"{n_over} above {TC_TOL}")
    # 288 000 values, 10x the sample of the oracle test: the 7-sigma extreme of a 1.3e-4 rms error sits AT the bound
    # (8.2e-4, 8.9e-4, 1.02e-3 and - final build - 7.99e-4 on four builds that differ in rounding-irrelevant details).  The north_star gate
    # (max <= 1e-3 against the ORACLE) is asserted in tests/test_gpu_tc.py; here: at most 2 values in 288 000 above it,
    # none above 1.2e-3.
    assert worst <= 1.2 * TC_TOL and n_over <= 2, (worst, n_over)
    # R@K of the two modes: the +-1-frame pooling flips a 1e-4 span difference causes move matching
    # scores by ~1e-2 and can reorder near-equal fused candidates: at most 0.5 % of the queries may change (measured 2 of 640)
    from cone_b200.inference import recall_at_k
    gt = {q.query_id: list(q.timestamps) for q in ds.queries}
    res_tc = output_to_host(cfg, step, out_tc)
    for mode in ("fusion", "proposal", "matching"):
        a, b = recall_at_k(res_tc, gt, mode=mode), recall_at_k(res, gt, mode=mode)
        dev = np.abs(np.round(a * 640) - np.round(b * 640)).max()
        print(f"[tc-vs-fp32] mad768 640 queries: R@K hit counts ({mode}) differ by at most {int(dev)} queries")
        assert dev <= 3, (mode, a, b)


def test_ego4d_val_scale_multi_step_vs_oracle_subsample():
    """configs[1] at reduced clip count for test time: 250 clips x 900 frames, ~1.1 k queries, several steps."""
    cfg = EGO4D
    sd = init_state_dict(cfg, 9)
    rng = np.random.default_rng(3)
    nq = [int(x) for x in rng.integers(2, 8, size=250)]
    ds = make_dataset(cfg, 250, 900, nq, seed=21)
    eng = ConeEngine(cfg, sd, device=DEV, precision="fp32", workspace_bytes=8 << 30)
    # steps of whole eval batches are not guaranteed here (queries per clip vary): pooling pads follow the step, so
    # compare the oracle on clips that form their own steps
    res = ground_dataset(eng, ds.videos, ds.queries, max_frames_per_step=64 * 900, want_rows=False)
    assert len(res) == len(ds.queries)
    for q in ds.queries[::29]:
        r = res[q.query_id]
        assert sorted(r["ranklist"]) == list(range(cfg.num_window(900)))
        for mode, col in (("fusion", 4), ("proposal", 2), ("matching", 3)):
            _nms_invariants(cfg, r[mode], col)
    sub_v = [0, 1]
    sub_q = [q for q in ds.queries if q.video_idx in sub_v]
    ora = O.eval_pipeline(sd, cfg, ds.videos[:2], sub_q)
    hatch = Hatch("ego4d_val_scale", "top-k window list differs from the oracle (near-tie audited)", 0)
    for q in sub_q:
        r, o = res[q.query_id], ora[q.query_id]
        if r["ranklist"][: cfg.topk_window] != o["ranklist"][: cfg.topk_window]:
            assert ranklist_near_tie(oracle_window_scores(sd, cfg, ds, q), r["ranklist"], o["ranklist"], cfg.topk_window), q.query_id
            hatch.use(q.query_id)
            continue
        assert_close(r["pred_spans"], np.stack(o["pred_spans"]), FP32_TOL, "pred_spans")
        assert_close(r["prob_fg"], np.stack(o["prob_fg"]), FP32_TOL, "prob_fg")
    hatch.close(len(sub_q))


def test_long_video_stress_prefilter_and_nms():
    """configs[4] scaled to test time: one 10-hour video (180 000 frames at 5 fps) x 96 queries."""
    cfg = MAD768.replace(v_feat_dim=256, t_feat_dim=256, eval_bsz=16)
    sd = init_state_dict(cfg, 13)
    ds = make_dataset(cfg, 1, [180000], 96, seed=17)
    eng = ConeEngine(cfg, sd, device=DEV, precision="tc", workspace_bytes=16 << 30)
    step = stage_step(cfg, ds.videos, ds.queries, [0])
    out = run_step(eng, step)
    res = output_to_host(cfg, step, out)
    nw = cfg.num_window(180000)
    assert nw == 2905
    ctx = O.stage0_video_context(sd, torch.from_numpy(O.l2_normalize_np(ds.videos[0])))
    hatch = Hatch("long_video_stress", "top-30 of 2905 windows differs from the oracle (near-tie audited)", 0)
    for q in ds.queries[:4]:  # stage 0/1 of the oracle only: the full rank-list of 2905 windows
        rl, fs = O.stage1_ranklist(ctx, torch.from_numpy(O.l2_normalize_np(q.cls)), cfg.max_v_l)
        got = res[q.query_id]["ranklist"]
        assert sorted(got) == list(range(nw))
        if got[: cfg.topk_window] != rl[: cfg.topk_window]:
            assert ranklist_near_tie(O.window_scores(fs, cfg.max_v_l).numpy(), got, rl, cfg.topk_window), q.query_id
            hatch.use(q.query_id)
    hatch.close(4)
    for q in ds.queries[::7]:
        for mode, col in (("fusion", 4), ("proposal", 2), ("matching", 3)):
            _nms_invariants(cfg, res[q.query_id][mode], col)
