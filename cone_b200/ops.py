"""Function-level mirrors of the reference's host ops, same names / arguments / return types, running on the GPU
through the C ABI: `span_cxw_to_xx` (cone/span_utils.py:25-41), `temporal_nms` (utils/temporal_nms.py:25-74),
`normalize_score` (utils/basic_utils.py:10-20), `compute_window_ranklist`
(run_on_video/cone_localizator.py:83-100 / cone/inference.py:284-299)."""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def span_cxw_to_xx(cxw_spans: torch.Tensor) -> torch.Tensor:
    """(..., 2) (center, width) -> (..., 2) (start, end).  Element-wise; the reference calls it on CPU tensors
    (inference.py:75), so it stays a torch expression with the reference's operation order; inside the fused
    path the same arithmetic is part of the `cone_fuse_nms` / `span_mean_pool` kernels."""
    x1 = cxw_spans[..., 0] - 0.5 * cxw_spans[..., 1]
    x2 = cxw_spans[..., 0] + 0.5 * cxw_spans[..., 1]
    return torch.stack([x1, x2], dim=-1)


def normalize_score(pre_list: List[float]) -> List[float]:
    """min-max normalisation, the list itself when constant (host scalar arithmetic in fp64, as the reference;
    in the fused path this is done inside `cone_fuse_nms`)."""
    amin, amax = min(pre_list), max(pre_list)
    if amin == amax:
        return pre_list
    return [(v - amin) / (amax - amin) for v in pre_list]


def temporal_nms(predictions: List[List[float]], nms_thd: float, max_after_nms: int = 100, device="cuda:0"):
    """Greedy temporal NMS on a list of [st, ed, score]; returns a new list, input not mutated."""
    lib = _lib.load()
    n = len(predictions)
    if n == 0:
        return []
    if n == 1:  # `if len(predictions) == 1: return predictions` (temporal_nms.py:40-41)
        return predictions
    dev = torch.device(device)
    t = torch.tensor([[float(p[0]), float(p[1]), float(p[2])] for p in predictions], dtype=torch.float64).t().contiguous()
    with torch.cuda.device(dev):
        t = t.to(dev)
        keep = torch.empty((max(min(max_after_nms, n), 1),), dtype=torch.int32, device=dev)
        cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
        _lib.check(lib.cone_temporal_nms(C.c_void_p(t[0].data_ptr()), C.c_void_p(t[1].data_ptr()),
                                         C.c_void_p(t[2].data_ptr()), n, float(nms_thd), int(min(max_after_nms, n)),
                                         C.c_void_p(keep.data_ptr()), C.c_void_p(cnt.data_ptr()), _stream()),
                   "cone_temporal_nms")
        k = int(cnt.item())
        idx = keep[:k].tolist()
    return [[predictions[i][0], predictions[i][1], predictions[i][2]] for i in idx]


def compute_window_ranklist(frame_matching_score: torch.Tensor, max_v_l: int) -> List[int]:
    """Window ids ranked by max frame score inside each window (score descending, index ascending on ties)."""
    lib = _lib.load()
    fs = frame_matching_score
    if not fs.is_cuda:
        raise _lib.ConeError("compute_window_ranklist needs a CUDA tensor (no CPU fallback)")
    fs = fs.float().contiguous()
    L = fs.numel()
    nw = math.ceil(L / int(max_v_l / 2)) + 1
    with torch.cuda.device(fs.device):
        off = torch.zeros((1,), dtype=torch.int64, device=fs.device)
        cnt = torch.full((1,), L, dtype=torch.int32, device=fs.device)
        rl = torch.empty((1, nw), dtype=torch.int32, device=fs.device)
        _lib.check(lib.cone_window_ranklist(C.c_void_p(fs.data_ptr()), C.c_void_p(off.data_ptr()),
                                            C.c_void_p(cnt.data_ptr()), 1, int(max_v_l), C.c_void_p(rl.data_ptr()),
                                            C.c_void_p(0), nw, _stream()), "cone_window_ranklist")
        return rl[0].tolist()
