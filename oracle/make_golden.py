"""Freeze outputs of the UNMODIFIED reference as golden fixtures under tests/golden/.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Inputs are regenerated from seeds by `cone_b200.synth` / `cone_b200.weights`, so the fixtures hold
only the reference's OUTPUTS (plus, for the window ranker, the reference's own frame scores, which
are that operator's input).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cone_b200.config import EGO4D, MAD512  # noqa: E402
from cone_b200.synth import make_dataset  # noqa: E402
from cone_b200.weights import init_state_dict  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (config, dataset kwargs, weight seed, perturb)
E2E_CASES = {
    "e2e_ego4d": (EGO4D.replace(eval_bsz=4), dict(n_videos=3, frames=[900, 40, 300], queries_per_video=[5, 2, 3],
                                                  seed=1), 0, True),
    "e2e_ego4d_freshinit": (EGO4D.replace(eval_bsz=3), dict(n_videos=2, frames=[450, 133], queries_per_video=[2, 2],
                                                           seed=2), 1, False),
    "e2e_mad512": (MAD512.replace(eval_bsz=2), dict(n_videos=2, frames=[2000, 700], queries_per_video=[3, 2],
                                                    seed=3), 2, True),
}


def dense_case(cfg, seed):
    """A dense padded batch as `start_end_collate` would emit it (ragged lengths, masks)."""
    g = torch.Generator().manual_seed(seed)
    B, Lv, Lt = 6, cfg.max_v_l, 14
    vlen = [Lv, Lv // 2, Lv, 7, Lv - 1, 1]
    tlen = [14, 5, 9, 14, 4, 11]
    vid = torch.randn(B, Lv, cfg.v_feat_dim, generator=g)
    txt = torch.randn(B, Lt, cfg.t_feat_dim, generator=g)
    txt = txt / (txt.norm(dim=-1, keepdim=True) + 1e-5)
    cls = torch.randn(B, cfg.v_feat_dim, generator=g)
    cls = cls / (cls.norm(dim=-1, keepdim=True) + 1e-5)
    vm = torch.zeros(B, Lv)
    tm = torch.zeros(B, Lt)
    for i in range(B):
        vm[i, : vlen[i]] = 1
        tm[i, : tlen[i]] = 1
        vid[i, vlen[i]:] = 0
        txt[i, tlen[i]:] = 0
    return vid, vm, txt, tm, cls


def main():
    assert R.reference_available(), "needs /root/reference"
    os.makedirs(GOLDEN, exist_ok=True)
    R.import_reference()
    from cone.span_utils import span_cxw_to_xx
    from utils.temporal_nms import temporal_nms
    from utils.basic_utils import normalize_score

    # ---- end-to-end cases: the reference's eval_epoch ------------------------------------------
    for name, (cfg, dkw, wseed, perturb) in E2E_CASES.items():
        sd = init_state_dict(cfg, wseed, perturb=perturb)
        ds = make_dataset(cfg, **dkw)
        ref = R.run_reference_eval_epoch(cfg, sd, ds, capture_frame_scores=3)
        arrays, lists = {}, {}
        for q in ds.queries:
            r = ref[q.query_id]
            for k in ("pred_spans", "prob_fg", "logits", "match", "frame_score"):
                if k in r:
                    arrays[f"{q.query_id}/{k}"] = np.asarray(r[k], dtype=np.float32)
            lists[q.query_id] = {k: r[k] for k in ("ranklist", "rows", "fusion", "proposal", "matching")}
        lists["_metrics"] = ref["_metrics"]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **arrays)
        with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
            json.dump(lists, f)
        print("wrote", name, len(ds.queries), "queries")

    # ---- operator-level: CONE.forward / forward_clip_matching on a dense ragged batch ------------
    for cname, cfg, wseed in (("dense_ego4d", EGO4D, 3), ("dense_mad512", MAD512, 4)):
        sd = init_state_dict(cfg, wseed)
        model = R.build_reference_model(cfg, sd)
        vid, vm, txt, tm, cls = dense_case(cfg, 100 + wseed)
        with torch.no_grad():
            out = model(src_txt=txt, src_txt_mask=tm, src_vid_motion=vid, src_vid_motion_mask=vm)
            match = model.forward_clip_matching(src_cls_txt=cls, src_vid_appear=vid, src_vid_appear_mask=vm,
                                                proposal=out["pred_spans"])
            adapted = model.adapter_layer(vid[0]) + vid[0]
        np.savez_compressed(os.path.join(GOLDEN, cname + ".npz"),
                            pred_logits=out["pred_logits"].numpy(), pred_spans=out["pred_spans"].numpy(),
                            saliency=out["saliency_scores"].numpy(), match=match.numpy(),
                            aux_logits=out["aux_outputs"][0]["pred_logits"].numpy(),
                            aux_spans=out["aux_outputs"][0]["pred_spans"].numpy(), adapted0=adapted.numpy())
        print("wrote", cname)

    # ---- operator-level: host ops ----------------------------------------------------------------
    rng = np.random.default_rng(7)
    nms_cases = []
    fixed = [[0, 10, .9], [1, 11, .8], [20, 30, .7], [0, 20, .6], [5, 15, .5]]
    for thd, mx in ((0.5, 5), (0.5, 2), (0.3, 100), (0.7, 3)):
        nms_cases.append(dict(inp=fixed, thd=thd, max_after=mx, out=temporal_nms([list(x) for x in fixed], thd, mx)))
    for n in (0, 1, 2, 3, 17, 150, 200):
        st = np.round(rng.uniform(0, 100, n), 4)
        ed = np.round(st + rng.uniform(0.5, 30, n), 4)
        sc = np.round(rng.uniform(0, 1, n), 2 if n > 10 else 4)  # 2 decimals: many score ties
        inp = [[float(a), float(b), float(c)] for a, b, c in zip(st, ed, sc)]
        if n >= 3:
            inp[2][0], inp[2][1] = inp[0][0], inp[0][1]  # identical span
            inp[1][1] = inp[1][0]  # zero-length span
        for thd, mx in ((0.5, 5), (0.5, 100), (0.0, 10), (1.0, 10)):
            nms_cases.append(dict(inp=inp, thd=thd, max_after=mx, out=temporal_nms([list(x) for x in inp], thd, mx)))
    host = dict(
        span_cxw_to_xx=dict(inp=[[0.5, 1.0], [0.3, 0.2]],
                            out=span_cxw_to_xx(torch.Tensor([[0.5, 1.0], [0.3, 0.2]])).tolist()),
        temporal_nms=nms_cases,
        normalize_score=[dict(inp=x, out=normalize_score(list(x))) for x in
                         ([0.1, 0.5, 0.3], [0.2, 0.2, 0.2], [1.0], [-1.0, 0.0, 2.5, 2.5])],
    )
    with open(os.path.join(GOLDEN, "host_ops.json"), "w") as f:
        json.dump(host, f)
    print("wrote host_ops")


if __name__ == "__main__":
    main()
