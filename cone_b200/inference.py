"""The reference's `eval_epoch` stages 0-3 (`cone/inference.py:227-322`) over in-memory features, driven through
`ConeEngine`: host features -> pinned staging -> HBM -> kernels -> per-query predictions in the reference's
submission format (`postprocessing_format_mad`, inference.py:169-202)."""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .config import ConeConfig
from .engine import ConeEngine, GroundingOutput, QueryBatch, pack_queries

MODES = ("fusion", "proposal", "matching")  # order of GroundingOutput.nms[:, m]


@dataclasses.dataclass
class HostStep:
    """One step's inputs staged in pinned host memory."""
    frames: torch.Tensor  # [n_frames, Dv] fp32 pinned
    qb: QueryBatch  # pinned
    video_ids: List[int]

    def h2d_bytes(self) -> int:
        return self.frames.numel() * 4 + self.qb.h2d_bytes()


def plan_steps(video_lengths: Sequence[int], queries, max_frames_per_step: int) -> List[List[int]]:
    """Group consecutive videos into steps of at most `max_frames_per_step` frames (>= 1 video each)."""
    steps, cur, tot = [], [], 0
    for v, L in enumerate(video_lengths):
        if cur and tot + L > max_frames_per_step:
            steps.append(cur)
            cur, tot = [], 0
        cur.append(v)
        tot += L
    if cur:
        steps.append(cur)
    return steps


def stage_step(cfg: ConeConfig, videos: Sequence[np.ndarray], queries, video_ids: Sequence[int], pin: bool = True) -> HostStep:
    """Concatenate the step's videos and pack its queries (dataset order preserved within the step)."""
    vid_set = {v: i for i, v in enumerate(video_ids)}
    sel = [(i, q) for i, q in enumerate(queries) if q.video_idx in vid_set]
    local = [dataclasses.replace(q, video_idx=vid_set[q.video_idx]) for _, q in sel]
    lens = [len(videos[v]) for v in video_ids]
    frames = torch.from_numpy(np.concatenate([videos[v] for v in video_ids], axis=0))
    first = sel[0][0] if sel else 0
    qb = pack_queries(cfg, lens, local, first_dataset_index=first)
    if pin and torch.cuda.is_available():
        frames = frames.pin_memory()
        qb = qb.pin()
    return HostStep(frames, qb, list(video_ids))


def run_step(engine: ConeEngine, step: HostStep, want_rows: bool = False) -> GroundingOutput:
    """H2D of the step's inputs (async from pinned memory) + the whole device path."""
    dev = engine.device
    with torch.cuda.device(dev):
        frames = step.frames.to(dev, non_blocking=True)
        qb = step.qb.to(dev)
        return engine.ground(frames, qb, want_rows=want_rows)


def output_to_host(cfg: ConeConfig, step: HostStep, out: GroundingOutput, full: bool = True) -> Dict[str, dict]:
    """D2H + conversion to per-query Python structures (same layout as the oracle's `eval_pipeline`)."""
    nms = out.nms.cpu().numpy()
    cnt = out.nms_count.cpu().numpy()
    res: Dict[str, dict] = {}
    if full:
        rl = out.ranklist.cpu().numpy()
        ws, wl = out.win_start.cpu().numpy(), out.win_len.cpu().numpy()
        sp, pr, mt = out.pred_spans.cpu().numpy(), out.prob_fg.cpu().numpy(), out.match.cpu().numpy()
        rows = out.rows.cpu().numpy() if out.rows is not None else None
        rcnt = out.rows_count.cpu().numpy() if out.rows_count is not None else None
    for j, qid in enumerate(step.qb.query_ids):
        r = {m: nms[j, i, : cnt[j, i]].tolist() for i, m in enumerate(MODES)}
        if full:
            n_win = int((wl[j] > 0).sum())
            r["ranklist"] = [int(x) for x in rl[j] if x >= 0]
            r["windows"] = [(int(ws[j, t]), int(wl[j, t])) for t in range(n_win)]
            r["pred_spans"] = sp[j, :n_win]
            r["prob_fg"] = pr[j, :n_win]
            r["match"] = mt[j, :n_win]
            if rows is not None:
                r["rows"] = rows[j, : rcnt[j]].tolist()
        res[qid] = r
    return res


def ground_dataset(engine: ConeEngine, videos: Sequence[np.ndarray], queries, max_frames_per_step: int = 1 << 20,
                   full: bool = True, want_rows: bool = True) -> Dict[str, dict]:
    """Stages 0-3 over a whole in-memory dataset.  Videos are processed in steps of consecutive videos; the
    reference pools proposals over windows padded to the longest window of each `eval_bsz` batch of queries
    (SURVEY.md §8 A9), which is reproduced exactly when a step holds whole eval batches (always true for a
    single step)."""
    res: Dict[str, dict] = {}
    lens = [len(v) for v in videos]
    for ids in plan_steps(lens, queries, max_frames_per_step):
        step = stage_step(engine.cfg, videos, queries, ids)
        if step.qb.tok_len.numel() == 0:
            continue
        out = run_step(engine, step, want_rows=want_rows)
        res.update(output_to_host(engine.cfg, step, out, full=full))
    return res


def to_submission(results: Dict[str, dict], annotations: Sequence[dict], mode: str = "fusion") -> List[dict]:
    """The reference's MAD submission rows (inference.py:169-202): one dict per query, dataset order."""
    return [dict(query_id=a["query_id"], video_id=a["video_id"], predicted_times=results[a["query_id"]][mode])
            for a in annotations]


def recall_at_k(results: Dict[str, dict], gt: Dict[str, Sequence[float]], mode: str = "fusion",
                thresholds=(0.3, 0.5), topk=(1, 5)) -> np.ndarray:
    """R@K at IoU thresholds over the NMS'd predictions — the metric of standalone_eval/evaluate_mad.py:60-104
    (hull IoU in fp32, strict `>`), computed from the kernel outputs on the host."""
    rec = np.zeros((len(topk), len(thresholds)))
    thr = np.asarray(thresholds, dtype=np.float32)
    for qid, r in results.items():
        rows = np.asarray(r[mode], dtype=np.float64)
        if rows.size == 0:
            continue
        st, ed = rows[: max(topk), 0].astype(np.float32), rows[: max(topk), 1].astype(np.float32)
        g0, g1 = np.float32(gt[qid][0]), np.float32(gt[qid][1])
        inter = np.maximum(np.minimum(ed, g1) - np.maximum(st, g0), np.float32(0))
        hull = np.maximum(ed, g1) - np.minimum(st, g0)
        hit = (inter / hull)[:, None] > thr[None, :]
        for i, k in enumerate(topk):
            rec[i] += hit[:k].any(axis=0)
    return rec / max(len(results), 1)
