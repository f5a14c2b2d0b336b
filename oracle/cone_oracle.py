"""CPU oracle: a restatement of the reference's coarse-to-fine inference path in torch-CPU fp32.

TEST INFRASTRUCTURE ONLY.  Nothing under `cone_b200/` imports this module; it is used by
`tests/`, by `__graft_entry__.smoke()` as the checker, and by `bench.py` for the `cpu_baseline`
/ `--impl reference` legs.  The product path is the CUDA library and fails loudly without it.

Parity pinning: the reference (houzhijian/CONE) has no tests or golden vectors for this path
(SURVEY.md §4), apart from one intact docstring example (`cone/span_utils.py:30-33`).  The
oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container:
`oracle/ref_harness.py` imports `/root/reference`, loads the same random state dict into the
reference `CONE`, runs the reference's own `eval_epoch` on seeded synthetic inputs and commits
the results under `tests/golden/`; `tests/test_oracle_golden.py` checks every function below
against those files.  The only deviation from the reference is the mandatory stable sort of the
window rank-list (`cone/inference.py:298` calls `torch.sort` without `stable=True`, whose tie
order depends on the torch build — SURVEY.md §7 H2); the contract is "score descending, then
window index ascending".

Arithmetic the reference delegates to PyTorch (nn.MultiheadAttention, LayerNorm, Linear,
softmax, sigmoid) is restated from torch's documented definitions and executed with torch-CPU
fp32 ops, i.e. the same library the reference runs on.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# A1  host-side L2 normalisation                                utils/basic_utils.py:97-99
# --------------------------------------------------------------------------------------
def l2_normalize_np(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """x / (||x||_2 + eps) over the last dim, numpy, dtype-preserving for float32 input."""
    return x / (np.linalg.norm(x, axis=-1, keepdims=True) + eps)


# --------------------------------------------------------------------------------------
# small building blocks                              cone/model.py:428-465 (MLP, LinearLayer)
# --------------------------------------------------------------------------------------
def mlp(sd: Dict[str, Tensor], prefix: str, x: Tensor, num_layers: int) -> Tensor:
    """`MLP.forward` (model.py:437-440): Linear+ReLU ... Linear."""
    for i in range(num_layers):
        x = F.linear(x, sd[f"{prefix}.layers.{i}.weight"], sd[f"{prefix}.layers.{i}.bias"])
        if i < num_layers - 1:
            x = F.relu(x)
    return x


def linear_layer(sd: Dict[str, Tensor], prefix: str, x: Tensor, relu: bool) -> Tensor:
    """`LinearLayer.forward` (model.py:458-465): LayerNorm(in) -> Linear -> optional ReLU."""
    w, b = sd[prefix + ".LayerNorm.weight"], sd[prefix + ".LayerNorm.bias"]
    x = F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
    x = F.linear(x, sd[prefix + ".net.1.weight"], sd[prefix + ".net.1.bias"])
    return F.relu(x) if relu else x


def input_proj(sd: Dict[str, Tensor], name: str, x: Tensor, n_input_proj: int = 2) -> Tensor:
    """`input_vid_proj` / `input_txt_proj` (model.py:55-72): ReLU on all but the last layer."""
    for i in range(n_input_proj):
        x = linear_layer(sd, f"{name}.{i}", x, relu=(i != n_input_proj - 1))
    return x


def adapter(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """`adapter_layer(x) + x` (model.py:80, 203-204; inference.py:255)."""
    return mlp(sd, "adapter_layer", x, 2) + x


# --------------------------------------------------------------------------------------
# A2  stage 0: adapted, normalised context features                 cone/inference.py:250-260
# --------------------------------------------------------------------------------------
def stage0_video_context(sd: Dict[str, Tensor], video_norm: Tensor) -> Tensor:
    """video_norm: (L, Dv) host-normalised features (A1).  Returns (L, Dv)."""
    a = adapter(sd, video_norm[None])  # the reference keeps the batch dim of its bs=1 loader
    a = a / a.norm(dim=2, keepdim=True)  # inference.py:257 — no eps
    return a[0]


# --------------------------------------------------------------------------------------
# A3  stage 1: frame scores -> window scores -> rank-list           cone/inference.py:276-299
# --------------------------------------------------------------------------------------
def window_scores(frame_score: Tensor, max_v_l: int) -> Tensor:
    """Max frame score inside each sliding window (inference.py:286-296)."""
    ctx_l = len(frame_score)
    stride = int(max_v_l / 2)
    num_window = math.ceil(ctx_l / stride) + 1
    out = []
    for i in range(num_window):
        s = max((i - 1) * stride, 0)
        e = min((i - 1) * stride + max_v_l, ctx_l)
        out.append(torch.max(frame_score[s:e]))
    return torch.Tensor(out)


def window_ranklist(frame_score: Tensor, max_v_l: int) -> List[int]:
    """Window ids by descending window score, ties by ascending id (inference.py:297-299 with
    the stable-sort patch, see module docstring)."""
    ws = window_scores(frame_score, max_v_l)
    _, idx = torch.sort(ws, descending=True, stable=True)
    return idx.tolist()


def stage1_ranklist(video_ctx: Tensor, cls_norm: Tensor, max_v_l: int) -> Tuple[List[int], Tensor]:
    """One query: einsum('db,b->d') then the window loop (inference.py:284-299)."""
    fs = torch.einsum("db,b->d", video_ctx, cls_norm)
    return window_ranklist(fs, max_v_l), fs


# --------------------------------------------------------------------------------------
# A4/A5  window slicing, collate, padding    cone/ego4d_mad_dataloader.py:144-159, 305-358
#                                            utils/tensor_utils.py:5-53
# --------------------------------------------------------------------------------------
def pad_sequences(seqs: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """Zero-pad to the longest sequence; float mask, 1 = valid (tensor_utils.py:38-53)."""
    n = max(len(s) for s in seqs)
    out = torch.zeros((len(seqs), n) + tuple(seqs[0].shape[1:]), dtype=torch.float32)
    mask = torch.zeros((len(seqs), n), dtype=torch.float32)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = s
        mask[i, : len(s)] = 1
    return out, mask


def slice_query_windows(video_raw: Tensor, ranklist: Sequence[int], topk_window: int, max_v_l: int):
    """The eval branch of `StartEndDataset.__getitem__` (dataloader:144-159): raw (un-normalised)
    rows of the first min(k, num_window) ranked windows; returns [(start, length, rows)]."""
    ctx_l = len(video_raw)
    stride = int(max_v_l / 2)
    out = []
    for i in list(ranklist)[:topk_window]:
        s = max((i - 1) * stride, 0)
        e = min((i - 1) * stride + max_v_l, ctx_l)
        out.append((s, e - s, video_raw[s:e, :]))
    return out


def prepare_query_text(tokens: np.ndarray, cls: np.ndarray, max_q_l: int) -> Tuple[Tensor, np.ndarray]:
    """`_get_query_feat_by_qid` (dataloader:258-282): truncate tokens to max_q_l, L2-normalise
    every token and the CLS vector on the host (eps 1e-5)."""
    q = l2_normalize_np(tokens[:max_q_l])
    c = l2_normalize_np(cls)
    return torch.from_numpy(np.ascontiguousarray(q)), c


# --------------------------------------------------------------------------------------
# A7  sine position embedding                               cone/position_encoding.py:51-72
# --------------------------------------------------------------------------------------
def position_embedding_sine(mask: Tensor, num_pos_feats: int = 256, temperature: float = 10000.0) -> Tensor:
    x_embed = mask.cumsum(1, dtype=torch.float32)
    x_embed = x_embed / (x_embed[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos = x_embed[:, :, None] / dim_t
    return torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=3).flatten(2)


# --------------------------------------------------------------------------------------
# nn.MultiheadAttention (torch; called at cone/transformer.py:239, 304, 308)
# --------------------------------------------------------------------------------------
def multihead_attention(sd: Dict[str, Tensor], prefix: str, query: Tensor, key: Tensor, value: Tensor,
                        nheads: int, key_padding_mask: Optional[Tensor]) -> Tensor:
    """Sequence-first (L, B, E) attention as torch defines it: separate q/k/v projections from the
    packed in_proj, q scaled by sqrt(1/head_dim), additive -inf key-padding mask, softmax, out_proj."""
    E = query.shape[-1]
    hd = E // nheads
    w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    Lq, B, _ = q.shape
    Lk = k.shape[0]
    q = q.reshape(Lq, B * nheads, hd).transpose(0, 1) * math.sqrt(1.0 / hd)
    k = k.reshape(Lk, B * nheads, hd).transpose(0, 1)
    v = v.reshape(Lk, B * nheads, hd).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2))  # (B*h, Lq, Lk)
    if key_padding_mask is not None:
        add = torch.zeros(key_padding_mask.shape, dtype=torch.float32)
        add.masked_fill_(key_padding_mask, float("-inf"))
        scores = scores + add[:, None, None, :].expand(B, nheads, 1, Lk).reshape(B * nheads, 1, Lk)
    attn = torch.softmax(scores, dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(Lq, B, E)
    return F.linear(out, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


# --------------------------------------------------------------------------------------
# A8  transformer                                  cone/transformer.py:49-73, 233-246, 296-317
# --------------------------------------------------------------------------------------
def encoder_layer(sd, p: str, src: Tensor, pad_mask: Tensor, pos: Tensor, nheads: int) -> Tensor:
    """`TransformerEncoderLayer.forward_post` (transformer.py:233-246)."""
    qk = src + pos
    src2 = multihead_attention(sd, p + ".self_attn", qk, qk, src, nheads, pad_mask)
    src = _ln(sd, p + ".norm1", src + src2)
    src2 = F.linear(F.relu(F.linear(src, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                    sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return _ln(sd, p + ".norm2", src + src2)


def decoder_layer(sd, p: str, tgt: Tensor, memory: Tensor, pad_mask: Tensor, pos: Tensor,
                  query_pos: Tensor, nheads: int) -> Tensor:
    """`TransformerDecoderLayer.forward_post` (transformer.py:296-317)."""
    qk = tgt + query_pos
    tgt2 = multihead_attention(sd, p + ".self_attn", qk, qk, tgt, nheads, None)
    tgt = _ln(sd, p + ".norm1", tgt + tgt2)
    tgt2 = multihead_attention(sd, p + ".multihead_attn", tgt + query_pos, memory + pos, memory, nheads, pad_mask)
    tgt = _ln(sd, p + ".norm2", tgt + tgt2)
    tgt2 = F.linear(F.relu(F.linear(tgt, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                    sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return _ln(sd, p + ".norm3", tgt + tgt2)


def transformer(sd, src: Tensor, pad_mask: Tensor, query_embed: Tensor, pos: Tensor, nheads: int,
                enc_layers: int, dec_layers: int) -> Tuple[Tensor, Tensor]:
    """`Transformer.forward` (transformer.py:49-73) with `return_intermediate_dec=True` (:459):
    returns hs (#dec_layers, B, nq, d) after the shared decoder norm (:135-141) and memory (B, S, d)."""
    B = src.shape[0]
    src = src.permute(1, 0, 2)
    pos = pos.permute(1, 0, 2)
    qpos = query_embed.unsqueeze(1).repeat(1, B, 1)
    tgt = torch.zeros_like(qpos)
    memory = src
    for i in range(enc_layers):
        memory = encoder_layer(sd, f"transformer.encoder.layers.{i}", memory, pad_mask, pos, nheads)
    inter = []
    out = tgt
    for i in range(dec_layers):
        out = decoder_layer(sd, f"transformer.decoder.layers.{i}", out, memory, pad_mask, pos, qpos, nheads)
        inter.append(_ln(sd, "transformer.decoder.norm", out))
    hs = torch.stack(inter).transpose(1, 2)
    return hs, memory.transpose(0, 1)


# --------------------------------------------------------------------------------------
# A6  CONE.forward                                                     cone/model.py:82-128
# --------------------------------------------------------------------------------------
def cone_forward(sd, src_txt: Tensor, src_txt_mask: Tensor, src_vid_motion: Tensor, src_vid_motion_mask: Tensor,
                 nheads: int = 8, enc_layers: int = 2, dec_layers: int = 2, n_input_proj: int = 2) -> Dict[str, Tensor]:
    src_vid = input_proj(sd, "input_vid_proj", src_vid_motion, n_input_proj)
    src_t = input_proj(sd, "input_txt_proj", src_txt, n_input_proj)
    src = torch.cat([src_vid, src_t], dim=1)
    mask = torch.cat([src_vid_motion_mask, src_txt_mask], dim=1).bool()
    d = src.shape[-1]
    pos = torch.cat([position_embedding_sine(src_vid_motion_mask, d), torch.zeros_like(src_t)], dim=1)
    hs, memory = transformer(sd, src, ~mask, sd["query_embed.weight"], pos, nheads, enc_layers, dec_layers)
    logits = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])
    spans = mlp(sd, "span_embed", hs, 3).sigmoid()
    vid_mem = memory[:, : src_vid.shape[1]]
    sal = F.linear(vid_mem, sd["saliency_proj.weight"], sd["saliency_proj.bias"]).squeeze(-1)
    return {"pred_logits": logits[-1], "pred_spans": spans[-1], "saliency_scores": sal,
            "aux_outputs": [{"pred_logits": a, "pred_spans": b} for a, b in zip(logits[:-1], spans[:-1])]}


# --------------------------------------------------------------------------------------
# span_cxw_to_xx                                                  cone/span_utils.py:25-41
# --------------------------------------------------------------------------------------
def span_cxw_to_xx(cxw: Tensor) -> Tensor:
    x1 = cxw[..., 0] - 0.5 * cxw[..., 1]
    x2 = cxw[..., 0] + 0.5 * cxw[..., 1]
    return torch.stack([x1, x2], dim=-1)


# --------------------------------------------------------------------------------------
# A9  fine-grained matching                              cone/model.py:130-152, 178-210
# --------------------------------------------------------------------------------------
def proposal_bounds(pred_spans: Tensor, vid_mask: Tensor) -> Tuple[Tensor, Tensor]:
    """int32 [start, end) frame bounds of every proposal (model.py:186-192)."""
    duration = torch.sum(vid_mask, dim=-1)
    prop = torch.einsum("bld,b->bld", span_cxw_to_xx(pred_spans), duration)
    start = F.relu(torch.floor(prop[:, :, 0]).to(torch.int32))
    end = torch.ceil(prop[:, :, 1]).to(torch.int32)
    return start, end


def clip_matching(sd, src_cls_txt: Tensor, src_vid_appear: Tensor, src_vid_appear_mask: Tensor,
                  pred_spans: Tensor) -> Tensor:
    """`forward_clip_matching(..., is_groundtruth=False)`: mean-pool the (zero-padded) appearance rows
    of each proposal — a Python slice, so `end` clips to the PADDED length and pad rows inside the
    slice are averaged in; an empty slice gives NaN — adapter + residual, L2-norm, dot with the
    normalised CLS."""
    t = src_cls_txt / src_cls_txt.norm(dim=1, keepdim=True)
    B, nq = pred_spans.shape[:2]
    start, end = proposal_bounds(pred_spans, src_vid_appear_mask)
    pooled = []
    for feat, ss, ee in zip(src_vid_appear, start, end):
        for s, e in zip(ss, ee):
            pooled.append(feat[s:e].mean(axis=0))
    p = torch.vstack(pooled)
    p = adapter(sd, p).reshape(B, nq, -1)
    p = p / p.norm(dim=2, keepdim=True)
    return torch.einsum("bld,bd->bl", p, t)


# --------------------------------------------------------------------------------------
# A10  compose per-window rows                                   cone/inference.py:47-53, 70-91
# --------------------------------------------------------------------------------------
def round4(x: float) -> float:
    return float(f"{x:.4f}")


def compose_window_rows(pred_spans: Tensor, prob_fg: Tensor, match: Tensor, duration: int, video_start: int,
                        clip_length: float) -> List[List[float]]:
    """One window: (nq,2),(nq,),(nq,) -> nq rows [st, ed, score, match], sorted by score (stable,
    descending) and rounded to 4 decimals (inference.py:75-83)."""
    spans = (span_cxw_to_xx(pred_spans) * duration + video_start) * clip_length
    rows = torch.cat([spans, prob_fg[:, None], match[:, None]], dim=1).tolist()
    rows = sorted(rows, key=lambda r: r[2], reverse=True)
    return [[round4(e) for e in r] for r in rows]


# --------------------------------------------------------------------------------------
# A11-A13  fusion, NMS                 cone/inference.py:103-127, 205-217; utils/temporal_nms.py
# --------------------------------------------------------------------------------------
def normalize_score(vals: List[float]) -> List[float]:
    """min-max; the list itself when constant (basic_utils.py:10-20)."""
    lo, hi = min(vals), max(vals)
    if lo == hi:
        return vals
    return [(v - lo) / (hi - lo) for v in vals]


def score_fusion(rows: List[List[float]]) -> "Dict[Tuple[float, float], List[float]]":
    """dict keyed by (st, ed): duplicates collapse, last value wins, first position kept
    (inference.py:205-217)."""
    a = normalize_score([r[2] for r in rows])
    b = normalize_score([r[3] for r in rows])
    out: Dict[Tuple[float, float], List[float]] = {}
    for r, f in zip(rows, [x + y for x, y in zip(a, b)]):
        out[(r[0], r[1])] = [r[2], r[3], f]
    return out


def temporal_iou_hull(a: Sequence[float], b: Sequence[float]) -> float:
    """`compute_temporal_iou` (temporal_nms.py:6-22): intersection over the HULL, 0 if hull is 0."""
    inter = max(0, min(a[1], b[1]) - max(a[0], b[0]))
    hull = max(a[1], b[1]) - min(a[0], b[0])
    return 0 if hull == 0 else 1.0 * inter / hull


def temporal_nms(predictions: List[List[float]], nms_thd: float, max_after_nms: int = 100) -> List[List[float]]:
    """Greedy NMS (temporal_nms.py:25-74): stable sort by score descending; keep the head, drop every
    later entry whose hull-IoU with it is strictly above `nms_thd`; stop at `max_after_nms`."""
    if len(predictions) == 1:
        return predictions
    order = sorted(predictions, key=lambda x: x[2], reverse=True)
    alive = list(order)
    kept: List[List[float]] = []
    while len(alive) > 1 and len(kept) < max_after_nms:
        head = alive[0]
        alive = [head] + [c for c in alive[1:] if not temporal_iou_hull(head, c) > nms_thd]
        kept.append(alive.pop(0))
    if len(kept) < max_after_nms and len(alive) >= 1:
        kept.append(alive.pop(0))
    return [[k[0], k[1], k[2]] for k in kept]


def post_processing_nms(fused: Dict[Tuple[float, float], List[float]], idx: int, nms_thd: float,
                        max_before_nms: int, max_after_nms: int) -> List[List[float]]:
    """`post_processing_mr_nms` (inference.py:103-127): rows [st, ed, score, match, fusion]."""
    moments = sorted([[k[0], k[1], v[idx]] for k, v in fused.items()], key=lambda x: x[2], reverse=True)
    if nms_thd != -1:
        kept = temporal_nms(moments[:max_before_nms], nms_thd=nms_thd, max_after_nms=max_after_nms)
        return [[m[0], m[1]] + fused[(m[0], m[1])] for m in kept]
    return [[m[0], m[1]] + fused[(m[0], m[1])] for m in moments][:max_after_nms]


def postprocess_query(cfg, windows: Sequence[Tuple[int, int]], pred_spans, prob_fg, match) -> Dict[str, list]:
    """Stages A10-A13 for ONE query from raw model outputs: `windows` = [(video_start, length)] in
    window-rank order; pred_spans (k,nq,2), prob_fg (k,nq), match (k,nq).  Returns rows + the three
    NMS'd lists (inference.py:70-91, 169-217, 103-127)."""
    rows: List[List[float]] = []
    for (s, n), sp, pr, mt in zip(windows, pred_spans, prob_fg, match):
        rows.extend(compose_window_rows(torch.as_tensor(sp), torch.as_tensor(pr), torch.as_tensor(mt),
                                        int(n), int(s), cfg.clip_length))
    fused = score_fusion(rows)
    out = {"rows": rows}
    for name, idx in (("fusion", 2), ("proposal", 0), ("matching", 1)):
        out[name] = post_processing_nms(fused, idx, cfg.nms_thd, cfg.max_before_nms, cfg.max_after_nms)
    return out


# --------------------------------------------------------------------------------------
# A14  R@K / IoU (parity judge)                          standalone_eval/evaluate_mad.py:32-104
# --------------------------------------------------------------------------------------
def recall_at_k_iou(pred: Dict[str, List[List[float]]], gt: Dict[str, Sequence[float]],
                    thresholds=(0.3, 0.5), topk=(1, 5)) -> np.ndarray:
    """recall[k, thr]: fraction of queries with any of the first k predictions whose hull-IoU with the
    ground truth is strictly above thr.  IoU in float32 as `_iou` does (evaluate_mad.py:32-37)."""
    rec = np.zeros((len(topk), len(thresholds)), dtype=np.float64)
    thr = torch.tensor(thresholds)
    for qid, rows in pred.items():
        g = torch.tensor(gt[qid])
        m = torch.tensor(rows)[: max(topk)]
        if m.numel() == 0:
            continue
        st, ed = m[:, 0].float(), m[:, 1].float()
        inter = ed.min(g[1].float()) - st.max(g[0].float())
        hull = ed.max(g[1].float()) - st.min(g[0].float())
        iou = inter.clamp(min=0) / hull
        hit = iou[:, None] > thr
        for i, k in enumerate(topk):
            rec[i] += hit[:k].any(dim=0).numpy()
    return rec / max(len(pred), 1)


def recall_ego4d(pred: Dict[str, List[List[float]]], gt: Dict[str, Sequence[float]],
                 thresholds=(0.3, 0.5), topk=(1, 5)):
    """standalone_eval/evaluate_ego4d_nlq.py:41-62, 65-117: float64 IoU (intersection and hull both clamped at 0),
    results[thr][k] = mean over queries of any(IoU[:k] > thr); mIoU = mean IoU of the FIRST prediction.
    Returns (recall [len(thresholds), len(topk)] float64, mIoU)."""
    hits = [[[] for _ in topk] for _ in thresholds]
    top1 = []
    for qid, rows in pred.items():
        p = np.array([r[:2] for r in rows], dtype=np.float64)
        g = np.array([list(gt[qid])], dtype=np.float64)
        inter = np.maximum(0.0, np.minimum(p[:, 1, None], g[None, :, 1]) - np.maximum(p[:, 0, None], g[None, :, 0]))
        hull = np.maximum(0.0, np.maximum(p[:, 1, None], g[None, :, 1]) - np.minimum(p[:, 0, None], g[None, :, 0]))
        with np.errstate(invalid="ignore", divide="ignore"):
            overlap = 1.0 * inter / hull
        top1.append(overlap[0])
        for t, thr in enumerate(thresholds):
            for r, k in enumerate(topk):
                hits[t][r].append((overlap > thr)[:k].any())
    return np.array(hits).mean(axis=-1), float(np.mean(top1))


def window_recall(ranklists: Dict[str, List[int]], gt: Dict[str, Sequence[float]], clip_length: float, max_v_l: int,
                  topk=(1, 5, 10, 30, 50)) -> np.ndarray:
    """standalone_eval/evaluate_pre_filtered_window.py:30-72: a query is recalled at K when one of its first K ranked
    window ids lies in range(floor(start / sws), ceil(end / sws) + 1), start / end = timestamps / clip_length in
    frames, sws = int(max_v_l / 2).  float32 result, as the reference's torch.zeros accumulator."""
    sws = int(max_v_l / 2)
    rec = torch.zeros(len(topk))
    for qid, wl in ranklists.items():
        start, end = gt[qid][0] / clip_length, gt[qid][1] / clip_length
        true = set(range(math.floor(start / sws), math.ceil(end / sws) + 1))
        bools = [w in true for w in wl[: max(topk)]]
        for i, k in enumerate(topk):
            rec[i] += float(any(bools[:k]))
    rec /= max(len(ranklists), 1)
    return rec.numpy()


# --------------------------------------------------------------------------------------
# single-video front end      run_on_video/cone_localizator.py:84-221
# --------------------------------------------------------------------------------------
def predict_moment(sd: Dict[str, Tensor], cfg, video_feats: Tensor, text_token_feats: Tensor, text_cls_feat: Tensor,
                   max_before_nms: int = 100, nms_thd: float = 0.5, max_after_nms: int = 5) -> Dict[str, object]:
    """`CONELocalizator.predict_moment` for one (video, query).  Differences from eval_epoch that are kept:
    F.normalize (x / max(||x||, 1e-5)) of frames and tokens (:127-131); the NORMALISED frames are what is
    sliced, projected and pooled (:156); the adapter output is NOT re-normalised before ranking and the CLS
    vector is used raw (:133-138); every window is zero-padded to max_v_l / max_q_l (:141-165, so more than
    max_q_l tokens is an error, as in the reference); spans are scaled by max_v_l, not by the window's length
    (:188); slots are not sorted within a window (:186-197); only the fusion ranking is produced (:199-219).
    The rank-list sort is stable (module docstring).  A video with fewer than topk_window windows makes the
    reference run the model on all-masked rows (NaN rows enter the fusion); here only real windows are used.
    Returns {"ranklist", "windows", "pred_spans", "prob_fg", "match", "rows", "moments"}."""
    k, Lv, Lt = cfg.topk_window, cfg.max_v_l, cfg.max_q_l
    with torch.no_grad():
        v = F.normalize(video_feats.float(), dim=-1, eps=1e-5)
        tok = F.normalize(text_token_feats.float(), dim=-1, eps=1e-5)
        if tok.shape[0] > Lt:
            raise ValueError(f"{tok.shape[0]} tokens exceed max_q_l={Lt} (pad_feature would fail in the reference)")
        a = adapter(sd, v)
        ranklist = window_ranklist(torch.einsum("db,b->d", a, text_cls_feat.float()), Lv)
        stride = int(Lv / 2)
        wins = [(max((i - 1) * stride, 0), min((i - 1) * stride + Lv, len(v))) for i in ranklist[:k]]
        B = len(wins)
        vid = torch.zeros((B, Lv, v.shape[1]))
        vmask = torch.zeros((B, Lv))
        txt = torch.zeros((B, Lt, tok.shape[1]))
        tmask = torch.zeros((B, Lt))
        for i, (s0, e0) in enumerate(wins):
            vid[i, : e0 - s0] = v[s0:e0]
            vmask[i, : e0 - s0] = 1
            txt[i, : len(tok)] = tok
            tmask[i, : len(tok)] = 1
        cls = text_cls_feat.float()[None].repeat(B, 1)
        out = cone_forward(sd, txt, tmask, vid, vmask, cfg.nheads, cfg.enc_layers, cfg.dec_layers, cfg.n_input_proj)
        prob = F.softmax(out["pred_logits"], -1)[..., 0]
        match = clip_matching(sd, cls, vid, vmask, out["pred_spans"])
        rows: List[List[float]] = []
        for i, (s0, _) in enumerate(wins):
            spans = (span_cxw_to_xx(out["pred_spans"][i]) * Lv + torch.tensor(s0, dtype=torch.int)) * cfg.clip_length
            cur = torch.cat([spans, prob[i][:, None], match[i][:, None]], dim=1).tolist()
            rows.extend([[round4(e) for e in r] for r in cur])
    fused = score_fusion(rows)
    moments = sorted([[kk[0], kk[1], vv[2]] for kk, vv in fused.items()], key=lambda x: x[2], reverse=True)
    moments = temporal_nms(moments[:max_before_nms], nms_thd, max_after_nms)
    return {"ranklist": ranklist, "windows": [(s0, e0 - s0) for s0, e0 in wins], "pred_spans": out["pred_spans"].numpy(),
            "prob_fg": prob.numpy(), "match": match.numpy(), "rows": rows, "moments": moments}


# --------------------------------------------------------------------------------------
# the whole path: eval_epoch stages 0-3                        cone/inference.py:227-322
# --------------------------------------------------------------------------------------
def eval_pipeline(sd: Dict[str, Tensor], cfg, videos: Sequence[np.ndarray], queries, *, collect_raw: bool = True,
                  progress=None) -> Dict[str, dict]:
    """Run stages 0-3 for in-memory data.  `queries` is a sequence with fields (query_id, video_idx,
    tokens, cls) in dataset order; `cfg` a `cone_b200.config.ConeConfig`-like object.

    Returns {query_id: {ranklist, windows: [(start, length)], pred_spans, prob_fg, match (raw model
    outputs, window-rank order), rows, fusion, proposal, matching}}.
    """
    nheads, el, dl, nip = cfg.nheads, cfg.enc_layers, cfg.dec_layers, cfg.n_input_proj
    with torch.no_grad():
        # stage 0 (inference.py:250-260): per video
        ctx = [stage0_video_context(sd, torch.from_numpy(l2_normalize_np(v))) for v in videos]
        raw = [torch.from_numpy(v) for v in videos]
        res: Dict[str, dict] = {}
        # stage 1 (inference.py:276-299): per query
        for q in queries:
            cls_n = torch.from_numpy(l2_normalize_np(q.cls))
            rl, _ = stage1_ranklist(ctx[q.video_idx], cls_n, cfg.max_v_l)
            res[q.query_id] = {"ranklist": rl}
        # stage 2 (inference.py:306-315, 29-100): batches of eval_bsz queries
        for b0 in range(0, len(queries), cfg.eval_bsz):
            batch = queries[b0: b0 + cfg.eval_bsz]
            vids, toks, clss, meta = [], [], [], []
            for q in batch:
                tok, cls_n = prepare_query_text(q.tokens, q.cls, cfg.max_q_l)
                for (s, n, rows) in slice_query_windows(raw[q.video_idx], res[q.query_id]["ranklist"],
                                                        cfg.topk_window, cfg.max_v_l):
                    vids.append(rows)
                    toks.append(tok)
                    clss.append(cls_n)
                    meta.append((q.query_id, s, n))
            src_vid, vid_mask = pad_sequences(vids)
            src_txt, txt_mask = pad_sequences(toks)
            src_cls = torch.FloatTensor(np.stack(clss))
            out = cone_forward(sd, src_txt, txt_mask, src_vid, vid_mask, nheads, el, dl, nip)
            prob = F.softmax(out["pred_logits"], -1)[..., 0]
            match = clip_matching(sd, src_cls, src_vid, vid_mask, out["pred_spans"])
            for i, (qid, s, n) in enumerate(meta):
                r = res[qid]
                r.setdefault("windows", []).append((s, n))
                r.setdefault("rows", []).extend(
                    compose_window_rows(out["pred_spans"][i], prob[i], match[i], n, s, cfg.clip_length))
                if collect_raw:
                    r.setdefault("pred_spans", []).append(out["pred_spans"][i].numpy().copy())
                    r.setdefault("prob_fg", []).append(prob[i].numpy().copy())
                    r.setdefault("match", []).append(match[i].numpy().copy())
            if progress is not None:
                progress(b0 + len(batch))
        # stage 3 (inference.py:169-217, 103-127): per query
        for q in queries:
            r = res[q.query_id]
            fused = score_fusion(r["rows"])
            for name, idx in (("fusion", 2), ("proposal", 0), ("matching", 1)):
                r[name] = post_processing_nms(fused, idx, cfg.nms_thd, cfg.max_before_nms, cfg.max_after_nms)
    return res
