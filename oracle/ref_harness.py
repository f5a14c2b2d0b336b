"""Runs the UNMODIFIED reference (`/root/reference`, houzhijian/CONE) on in-memory synthetic data.

TEST / BENCH INFRASTRUCTURE ONLY.  Usable where `/root/reference` exists (the build container) or
where `oracle/vendor_ref.py` has placed the reference's files under `oracle/_ref` (they travel to the
GPU box with the snapshot).  It is how the oracle is pinned: `python -m oracle.make_golden`
calls `run_reference_eval_epoch` and freezes the reference's own outputs under `tests/golden/`; it is also
what `bench.py --impl reference` times and what `tests/test_gpu_dropin.py` drives.

Mechanics:
* `lmdb` and `terminaltables` are absent here and there is no network: three-line stub modules are
  put in `sys.modules` before the import (LMDB I/O is out of scope; the tables are cosmetic).
* `PreFilteringDataset` / `StartEndDataset` are subclassed with `__init__` and the three `_get_*`
  readers replaced by in-memory numpy; the reference's own `__getitem__`, `start_end_collate`,
  `prepare_batch_inputs`, `eval_epoch`, post-processing and metric code run verbatim.
* the one mandatory patch: `torch.sort` inside `cone.inference` is made stable.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import tempfile
import types
from types import SimpleNamespace
from typing import Dict, List

import numpy as np
import torch

def _reference_root() -> str:
    """`/root/reference` in the build container; on the GPU box the byte-identical files that `oracle/vendor_ref.py`
    placed under `oracle/_ref` (hash-checked against their manifest)."""
    env = os.environ.get("CONE_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/cone/inference.py"):
        return "/root/reference"
    from oracle import vendor_ref
    if vendor_ref.verify():
        return vendor_ref.DEST
    return "/root/reference"


REFERENCE_ROOT = _reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cone", "inference.py"))


def _install_stubs() -> None:
    if "lmdb" not in sys.modules:
        sys.modules["lmdb"] = types.ModuleType("lmdb")
    if "terminaltables" not in sys.modules:
        m = types.ModuleType("terminaltables")

        class AsciiTable:  # cosmetic stand-in
            def __init__(self, data, title=None):
                self.data, self.title, self.justify_columns = data, title, {}

            @property
            def table(self):
                return "\n".join(" | ".join(str(c).replace("\n", " ") for c in row) for row in self.data)

        m.AsciiTable = AsciiTable
        sys.modules["terminaltables"] = m
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")

        class EasyDict(dict):  # attribute access is all run_on_video/cone_localizator.py:52 needs
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__

        m.EasyDict = EasyDict
        sys.modules["easydict"] = m


def import_reference():
    """Returns the reference's `cone.inference` module (and makes `cone.*`, `utils.*` importable)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import cone.inference as ref_inf  # noqa: E402
    return ref_inf


def make_opt(cfg, results_dir: str, eval_path: str, device: str = "cpu") -> SimpleNamespace:
    """The fields `eval_epoch` and `build_model` read, defaults from cone/config.py."""
    return SimpleNamespace(
        device=torch.device(device), pin_memory=False, num_workers=0, adapter_module="linear",
        max_v_l=cfg.max_v_l, max_q_l=cfg.max_q_l, eval_bsz=cfg.eval_bsz, span_loss_type="l1",
        clip_length=cfg.clip_length, no_sort_results=False, debug=False, dset_name="mad",
        results_dir=results_dir, save_all=True, eval_modality="both", eval_split_name="val",
        eval_path=eval_path, nms_thd=cfg.nms_thd, max_before_nms=cfg.max_before_nms,
        max_after_nms=cfg.max_after_nms, topk_window=cfg.topk_window,
        hidden_dim=cfg.hidden_dim, dropout=0.1, nheads=cfg.nheads, dim_feedforward=cfg.dim_feedforward,
        enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers, pre_norm=False, position_embedding="sine",
        input_dropout=0.5, t_feat_dim=cfg.t_feat_dim, v_motion_feat_dim=cfg.v_feat_dim,
        v_appear_feat_dim=cfg.v_feat_dim, num_queries=cfg.num_queries, aux_loss=True, use_txt_pos=False,
        n_input_proj=cfg.n_input_proj, set_cost_span=10, set_cost_giou=1, set_cost_class=4,
        span_loss_coef=10, giou_loss_coef=1, label_loss_coef=4, lw_saliency=1.0, adapter_loss=True,
        adapter_loss_coef=1, eos_coef=0.1, temperature=0.07, saliency_margin=0.2, neg_loss=True,
    )


def build_reference_model(cfg, sd):
    """Reference `build_model` + `load_state_dict` of a state dict made by `cone_b200.weights`."""
    import_reference()
    from cone.model import build_model
    with tempfile.TemporaryDirectory() as td:
        opt = make_opt(cfg, td, os.path.join(td, "val.jsonl"))
        model, _ = build_model(opt)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model


def _make_datasets(ref_inf, ds, opt):
    from cone.ego4d_mad_dataloader import PreFilteringDataset, StartEndDataset
    vids = ds.video_ids
    vid2feat = {vids[i]: ds.videos[i] for i in range(len(vids))}
    qid2q = {q.query_id: q for q in ds.queries}
    ann = ds.annotations()
    cfg = ds.cfg

    class MemPre(PreFilteringDataset):
        def __init__(self):  # bypass LMDB (dataloader:410-431)
            self.dset_name, self.data_ratio, self.data_mode, self.use_video = "mad", 1, "context", True
            self.query_data = ann
            # keep dataset order deterministic (the reference uses list(set(...)))
            self.video_data = list(dict.fromkeys(item["clip_id"] for item in ann))
            self.video2idx = {v: i for i, v in enumerate(self.video_data)}

        def _get_video_appearance_feat_by_vid(self, vid):  # dataloader:453-460
            from utils.basic_utils import l2_normalize_np_array
            return torch.from_numpy(l2_normalize_np_array(vid2feat[vid]))

        def _get_query_feat_by_qid(self, qid):  # dataloader:462-473
            from utils.basic_utils import l2_normalize_np_array
            return l2_normalize_np_array(qid2q[qid].cls)

    class MemSE(StartEndDataset):
        def __init__(self):  # bypass LMDB (dataloader:31-96)
            self.dset_name, self.data_ratio = "mad", 1.0
            self.q_feat_type, self.max_q_l, self.max_v_l = "last_hidden_state", cfg.max_q_l, cfg.max_v_l
            self.ctx_mode, self.use_video = "video", True
            self.normalize_t, self.normalize_v, self.load_labels = True, True, True
            self.clip_len, self.max_windows, self.span_loss_type = cfg.clip_length, 5, "l1"
            self.txt_drop_ratio, self.topk_window = 0, cfg.topk_window
            self.slide_window_size = int(cfg.max_v_l / 2)
            self.eval, self.same_visual_path = True, True
            self.data = ann
            self.query_id2windowidx = None
            self.videofeat = {v: self._get_video_appearance_feat_by_vid(v) for v in vid2feat}

        def _get_video_appearance_feat_by_vid(self, vid):  # dataloader:294-302 returns the RAW rows
            return torch.from_numpy(vid2feat[vid])

        def _get_query_feat_by_qid(self, qid):  # dataloader:258-282
            from utils.basic_utils import l2_normalize_np_array
            q = qid2q[qid]
            tok = l2_normalize_np_array(q.tokens[: self.max_q_l])
            return torch.from_numpy(np.ascontiguousarray(tok)), l2_normalize_np_array(q.cls)

    return MemPre(), MemSE()


@contextlib.contextmanager
def _stable_sort(ref_inf):
    """make the rank-list sort at inference.py:298 stable."""
    real_torch = ref_inf.torch

    class _TorchProxy:
        def __getattr__(self, name):
            return getattr(real_torch, name)

        @staticmethod
        def sort(x, *a, **kw):
            kw.setdefault("stable", True)
            return real_torch.sort(x, *a, **kw)

    ref_inf.torch = _TorchProxy()
    try:
        yield
    finally:
        ref_inf.torch = real_torch


def run_reference_eval_epoch(cfg, sd, ds, capture_frame_scores: int = 0) -> Dict[str, dict]:
    """Reference `eval_epoch` (inference.py:227-499) on dataset `ds` with weights `sd`.

    Returns {query_id: {ranklist, pred_spans, prob_fg, match, rows, fusion, proposal, matching}} in the
    same layout as `oracle.cone_oracle.eval_pipeline`, plus "_metrics" (the reference's R@K table).
    """
    ref_inf = import_reference()
    model = build_reference_model(cfg, sd)
    out: Dict[str, dict] = {q.query_id: {} for q in ds.queries}
    with tempfile.TemporaryDirectory() as td:
        eval_path = os.path.join(td, "val.jsonl")
        with open(eval_path, "w") as f:
            for row in ds.annotations():
                f.write(json.dumps(row) + "\n")
        opt = make_opt(cfg, td, eval_path)
        pre, se = _make_datasets(ref_inf, ds, opt)

        # record the raw model outputs per batch without touching the reference code
        raw_batches: List[dict] = []
        orig_forward, orig_match = model.forward, model.forward_clip_matching

        def fwd(**kw):
            o = orig_forward(**kw)
            raw_batches.append({"pred_spans": o["pred_spans"].clone(),
                                "prob": torch.softmax(o["pred_logits"], -1)[..., 0].clone(),
                                "logits": o["pred_logits"].clone()})
            return o

        def mat(**kw):
            m = orig_match(**kw)
            raw_batches[-1]["match"] = m.clone()
            return m

        model.forward = fwd
        model.forward_clip_matching = mat
        captured = {}
        orig_get = ref_inf.get_eval_res

        def get_eval_res(*a, **kw):
            r = orig_get(*a, **kw)
            captured["mr_res"] = r[0]
            return r

        ref_inf.get_eval_res = get_eval_res
        frame_scores = {}
        orig_einsum = torch.einsum
        try:
            with _stable_sort(ref_inf), torch.no_grad():
                results, _, _, _ = ref_inf.eval_epoch(model, pre, se, opt, "inference_mad_val_test_preds.jsonl",
                                                      epoch_i=0, criterion=None, tb_writer=None)
        finally:
            ref_inf.get_eval_res = orig_get
            torch.einsum = orig_einsum
        ranklists = se.query_id2windowidx
        # raw outputs: batches follow dataset order, eval_bsz queries each, min(k, num_window) windows per query
        qi = 0
        queries = ds.queries
        for rb in raw_batches:
            row = 0
            for q in queries[qi: qi + cfg.eval_bsz]:
                n = min(cfg.topk_window, len(ranklists[q.query_id]))
                o = out[q.query_id]
                o["pred_spans"] = rb["pred_spans"][row: row + n].numpy()
                o["prob_fg"] = rb["prob"][row: row + n].numpy()
                o["logits"] = rb["logits"][row: row + n].numpy()
                o["match"] = rb["match"][row: row + n].numpy()
                row += n
            assert row == rb["pred_spans"].shape[0]
            qi += cfg.eval_bsz
        for q in queries:
            out[q.query_id]["ranklist"] = list(ranklists[q.query_id])
            out[q.query_id]["rows"] = []
        for item in captured["mr_res"]:
            out[item["query_id"]]["rows"].extend(item["pred_relevant_windows"])
        base = os.path.join(td, "inference_mad_val_test_preds.jsonl")
        for name, path in (("fusion", base), ("proposal", base.replace("preds", "proposal_preds")),
                           ("matching", base.replace("preds", "matching_preds"))):
            with open(path) as f:
                for line in f:
                    item = json.loads(line)
                    out[item["query_id"]][name] = item["predicted_times"]
        out["_metrics"] = {"fusion_recall": (results / 100.0).tolist() if hasattr(results, "tolist") else results}
        if capture_frame_scores:
            # operator-boundary input for the window ranker: the reference's own frame scores
            from oracle import cone_oracle as O
            for q in queries[:capture_frame_scores]:
                x = torch.from_numpy(O.l2_normalize_np(ds.videos[q.video_idx]))[None]
                a = model.adapter_layer(x) + x
                a = (a / a.norm(dim=2, keepdim=True))[0]
                c = torch.from_numpy(O.l2_normalize_np(q.cls))
                out[q.query_id]["frame_score"] = torch.einsum("db,b->d", a, c).detach().numpy()
    return out


def run_eval_epoch_files(cfg, ds, model, device: str = "cpu", timer=None) -> Dict[str, object]:
    """The reference's own `eval_epoch` (inference.py:227-499), unmodified, driven with ANY model object that has the
    reference's surface (the reference `CONE`, or `cone_b200.CONE` substituted as INTEGRATION.md §1 describes), on
    in-memory data.  Returns what the reference writes and returns: {"ranklists", "fusion" / "proposal" / "matching"
    (query_id -> predicted_times of the three prediction files), "recall" (its R@K table, fractions), "seconds"}."""
    import time
    ref_inf = import_reference()
    with tempfile.TemporaryDirectory() as td:
        eval_path = os.path.join(td, "val.jsonl")
        with open(eval_path, "w") as f:
            for row in ds.annotations():
                f.write(json.dumps(row) + "\n")
        opt = make_opt(cfg, td, eval_path, device=device)
        pre, se = _make_datasets(ref_inf, ds, opt)
        t0 = time.perf_counter()
        with _stable_sort(ref_inf), torch.no_grad():
            results, _, _, _ = ref_inf.eval_epoch(model, pre, se, opt, "inference_mad_val_test_preds.jsonl",
                                                  epoch_i=0, criterion=None, tb_writer=None)
        seconds = time.perf_counter() - t0
        out: Dict[str, object] = {"ranklists": {k: list(v) for k, v in se.query_id2windowidx.items()}, "seconds": seconds}
        base = os.path.join(td, "inference_mad_val_test_preds.jsonl")
        for name, path in (("fusion", base), ("proposal", base.replace("preds", "proposal_preds")),
                           ("matching", base.replace("preds", "matching_preds"))):
            with open(path) as f:
                out[name] = {json.loads(line)["query_id"]: json.loads(line)["predicted_times"] for line in f}
        out["recall"] = (results / 100.0).tolist() if hasattr(results, "tolist") else results
    return out


def run_reference_localizer(cfg, sd, video_feats, text_token_feats, text_cls_feat):
    """Reference `CONELocalizator.predict_moment` (run_on_video/cone_localizator.py:121-221) on one video / query.
    The class is instantiated without its `__init__` (which builds the Ego4D model from a checkpoint file): the
    model is the reference `CONE` with `sd` loaded, the module-level `args` are set from `cfg`, and the rank-list
    sort is made stable.  Returns (moments, ranklist)."""
    import_reference()
    import run_on_video.cone_localizator as L
    loc = object.__new__(L.CONELocalizator)
    loc.device = "cpu"
    loc.localizator = build_reference_model(cfg, sd)
    loc.slide_window_size = int(cfg.max_v_l / 2)
    loc.max_v_l = cfg.max_v_l
    saved = dict(L.args)
    L.args.update(max_v_l=cfg.max_v_l, max_q_l=cfg.max_q_l, topk_window=cfg.topk_window, clip_length=cfg.clip_length,
                  v_appear_feat_dim=cfg.v_feat_dim, v_motion_feat_dim=cfg.v_feat_dim, t_feat_dim=cfg.t_feat_dim)
    try:
        with _stable_sort(L):
            v = torch.as_tensor(video_feats)
            ranklist = loc.compute_window_ranklist(
                loc.localizator.adapter_layer(torch.nn.functional.normalize(v, dim=-1, eps=1e-5))
                + torch.nn.functional.normalize(v, dim=-1, eps=1e-5), torch.as_tensor(text_cls_feat))
            moments = loc.predict_moment(v, (torch.as_tensor(text_token_feats), torch.as_tensor(text_cls_feat)))
    finally:
        L.args.clear()
        L.args.update(saved)
    return moments, ranklist
