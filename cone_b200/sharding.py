"""Multi-GPU layer: whole videos (with all their queries) are independent units through every
stage, so they are sharded across ranks with no data-path collective; the only exchange is one all-gather of
the fixed-size per-query prediction blocks at the end (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def video_cost(n_frames: int, n_queries: int, topk: int, alpha: float = 1.0, beta: float = 150.0) -> float:
    """cost ~ alpha * L (stage 0/1 streams the video once) + beta * Nq * k (windows through Moment-DETR)."""
    return alpha * n_frames + beta * n_queries * topk


def lpt_assign(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of videos to ranks; deterministic (ties -> lower id)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def shard_queries(queries: Sequence, world_size: int, rank: int, eval_bsz: int = 1) -> list:
    """One giant video (BASELINE.json configs[4]): the video is replicated on every rank and its QUERIES
    are sharded.  Whole eval batches of `eval_bsz` consecutive queries stay together (the reference pools proposals over
    windows padded to the longest window of an eval batch), batches are dealt round-robin."""
    n_batches = (len(queries) + eval_bsz - 1) // eval_bsz
    out = []
    for b in range(rank, n_batches, world_size):
        out.extend(queries[b * eval_bsz: (b + 1) * eval_bsz])
    return out


def gather_predictions(nms: torch.Tensor, count: torch.Tensor, qid: torch.Tensor | None = None, group=None,
                       equal_shards: bool = False) -> Tuple[torch.Tensor, ...]:
    """All-gather per-query prediction blocks [Nq_local, 3, max_after, 5] (+ counts [Nq_local, 3], + optional
    int64 query ordinals) from every rank.  Ranks may hold different numbers of queries: blocks are padded to the
    largest shard, gathered with one collective per tensor and trimmed.  Returns tensors concatenated in rank
    order.  `equal_shards=True` skips the size exchange when all ranks are known to hold equally many queries."""
    if not dist.is_available() or not dist.is_initialized():
        return (nms, count) if qid is None else (nms, count, qid)
    world = dist.get_world_size(group)
    if equal_shards:  # every rank holds the same number of queries: no size exchange, no host sync
        sizes = [nms.shape[0]] * world
    else:
        n_local = torch.tensor([nms.shape[0]], dtype=torch.int64, device=nms.device)
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local, group=group)
        sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes)

    def pad(t):
        if t.shape[0] == n_max:
            return t.contiguous()
        p = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        p[: t.shape[0]] = t
        return p

    outs = []
    for t in (nms, count) + ((qid,) if qid is not None else ()):
        buf = torch.empty((world, n_max) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf.view(-1), pad(t).view(-1), group=group) if t.is_cuda else \
            dist.all_gather(list(buf.unbind(0)), pad(t), group=group)
        outs.append(torch.cat([buf[r, : sizes[r]] for r in range(world)], dim=0))
    return tuple(outs)


def reduce_counters(counters, group=None):
    """Sum the device-side metric counters of `inference.MetricCounters` over ranks (one all-reduce per tensor:
    integer hit counts simply add up) and all-gather the per-query top-1 IoUs (Ego4D mIoU)."""
    if not dist.is_available() or not dist.is_initialized():
        return counters
    for t in (counters.hits, counters.window_hits, counters.n_queries):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    # Whether the top-1 IoUs are gathered is decided COLLECTIVELY: a rank that received no videos or queries
    # (world > n_videos, an empty query shard) has an empty list but must still enter the collectives.
    dev = counters.hits.device
    flag = torch.tensor([1 if counters.top1_iou else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if int(flag.item()):
        local = torch.cat(counters.top1_iou) if counters.top1_iou else torch.zeros((0, 3), dtype=torch.float64, device=dev)
        gathered = gather_predictions(local, local.new_zeros((local.shape[0], 1)), group=group)[0]
        counters.top1_iou = [gathered]
    return counters
