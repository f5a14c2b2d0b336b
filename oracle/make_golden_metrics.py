"""Freeze outputs of the reference's own metric scripts (standalone_eval/) as tests/golden/metrics.json.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_metrics
The fixture holds the (small, seeded) predictions / rank-lists / ground truths AND what
`evaluate_mad.evaluate_nlq_performance`, `evaluate_ego4d_nlq.evaluate_nlq_performance` and
`evaluate_pre_filtered_window.windows_selection` return for them.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as R  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "metrics.json")


def synth_predictions(rng, n_queries, max_after, video_s):
    """Per query: ground truth + three NMS'd lists with rows [st, ed, score, match, fusion] (4-decimal seconds as
    `inference.py:88` writes them), some of them placed to hit IoU thresholds exactly."""
    out = []
    for i in range(n_queries):
        g0 = float(np.round(rng.uniform(0, video_s - 40), 1 if i % 3 else 0))
        g1 = g0 + float(rng.integers(2, 30))
        lists = []
        for m in range(3):
            n = int(rng.integers(1, max_after + 1))
            rows = []
            for j in range(n):
                kind = rng.integers(0, 6)
                if kind == 0:  # IoU exactly 0.5 with the ground truth (hull = 2 x intersection): must NOT count at 0.5
                    st, ed = g0, g0 + 2 * (g1 - g0)
                elif kind == 1:  # disjoint
                    st = g1 + float(rng.uniform(0.1, 50))
                    ed = st + float(rng.uniform(1, 20))
                elif kind == 2:  # touching end point: intersection 0
                    st, ed = g1, g1 + float(rng.uniform(1, 20))
                else:  # jittered copy of the ground truth
                    st = g0 + float(rng.normal(0, 4))
                    ed = max(st + 0.2, g1 + float(rng.normal(0, 4)))
                rows.append([float(f"{st:.4f}"), float(f"{ed:.4f}"), float(f"{rng.uniform():.4f}"),
                             float(f"{rng.uniform():.4f}"), float(rng.uniform(0, 2))])
            lists.append(rows)
        out.append(dict(query_id=f"ann{i // 4:03d}_{i % 4}", gt=[g0, g1] if i % 5 else [int(g0), int(g1)], lists=lists))
    return out


def main():
    assert R.reference_available(), "needs /root/reference"
    R.import_reference()
    import standalone_eval.evaluate_ego4d_nlq as ego4d_eval
    import standalone_eval.evaluate_mad as mad_eval
    import standalone_eval.evaluate_pre_filtered_window as window_eval

    rng = np.random.default_rng(11)
    cases = []
    for name, n, max_after, thr, topk in (("mad_default", 60, 5, [0.1, 0.3, 0.5], [1, 5, 10, 50, 100]),
                                          ("north_star", 41, 5, [0.3, 0.5], [1, 5]),
                                          ("deep_lists", 25, 20, [0.3, 0.5, 0.7], [1, 3, 10])):
        items = synth_predictions(rng, n, max_after, 600.0)
        case = dict(name=name, thresholds=thr, topk=topk, max_after=max_after, items=items, mad=[], ego4d=[])
        for m in range(3):
            # --- MAD: inference.py:333-375 passes torch tensors for thresholds / topK
            sub = [dict(query_id=it["query_id"], predicted_times=it["lists"][m]) for it in items]
            gt = [dict(query_id=it["query_id"], timestamps=it["gt"]) for it in items]
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                r = mad_eval.evaluate_nlq_performance(sub, gt, torch.tensor(thr), torch.tensor(topk))
            case["mad"].append(r.tolist())
            # --- Ego4D: inference.py:419-445 passes Python lists
            anns = {}
            for it in items:
                a, qi = it["query_id"].split("_")
                anns.setdefault(a, {})[int(qi)] = it["gt"]
            ground_truth = {"videos": [{"clips": [{"clip_uid": "clip0", "annotations": [
                {"annotation_uid": a, "language_queries": [
                    {"clip_start_sec": float(q[i][0]), "clip_end_sec": float(q[i][1])} for i in range(len(q))]}
                for a, q in anns.items()]}]}]}
            preds = [dict(clip_uid="clip0", annotation_uid=it["query_id"].split("_")[0],
                          query_idx=int(it["query_id"].split("_")[1]), predicted_times=it["lists"][m]) for it in items]
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                res, miou = ego4d_eval.evaluate_nlq_performance(preds, ground_truth, thr, topk)
            case["ego4d"].append(dict(recall=np.asarray(res).tolist(), mIoU=float(miou)))
        cases.append(case)

    # --- window pre-filtering recall
    wcases = []
    for name, max_v_l, clip_length, n_frames, topk in (("ego4d", 90, 0.53333, 900, [1, 5, 10, 30, 50]),
                                                       ("mad", 125, 0.2, 40000, [1, 5, 10, 30, 50, 100, 200])):
        opt = SimpleNamespace(clip_length=clip_length, max_v_l=max_v_l)
        stride = int(max_v_l / 2)
        n_win = int(np.ceil(n_frames / stride)) + 1
        q2w, gt = {}, []
        for i in range(50):
            perm = rng.permutation(n_win).tolist()
            if i % 7 == 0:
                perm = perm[: max(1, n_win // 3)]  # shorter video
            g0 = float(rng.uniform(0, n_frames * clip_length - 35))
            g1 = g0 + float(rng.uniform(1, 30))
            if i % 4 == 0:  # integer frame boundaries: floor / ceil edge
                g0 = float(int(g0 / clip_length / stride) * stride * clip_length)
            q2w[f"q{i}"] = [int(x) for x in perm]
            gt.append(dict(query_id=f"q{i}", timestamps=[g0, g1]))
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = window_eval.windows_selection(q2w, gt, torch.tensor(topk), opt)
        wcases.append(dict(name=name, max_v_l=max_v_l, clip_length=clip_length, topk=topk, ranklists=q2w,
                           gt={g["query_id"]: g["timestamps"] for g in gt}, recall=r.tolist()))

    with open(GOLDEN, "w") as f:
        json.dump(dict(cases=cases, window_cases=wcases), f)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
