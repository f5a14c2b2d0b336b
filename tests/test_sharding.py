"""Multi-GPU host logic on CPU: LPT sharding of videos and the prediction all-gather (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cone_b200.sharding import gather_predictions, lpt_assign, shard_queries, video_cost


def test_lpt_assign_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    costs = [video_cost(int(rng.integers(36000, 54000)), int(rng.integers(300, 900)), 30) for _ in range(112)]
    a = lpt_assign(costs, 8)
    assert a == lpt_assign(costs, 8)
    assert sorted(i for r in a for i in r) == list(range(112))
    loads = [sum(costs[i] for i in r) for r in a]
    assert max(loads) / (sum(loads) / 8) < 1.03  # within 3 % of perfect balance
    assert lpt_assign([5.0, 1.0], 4) == [[0], [1], [], []]


def test_shard_queries_keeps_eval_batches_together():
    qs = list(range(23))
    parts = [shard_queries(qs, 4, r, eval_bsz=4) for r in range(4)]
    assert sorted(x for p in parts for x in p) == qs
    assert parts[0] == [0, 1, 2, 3, 16, 17, 18, 19] and parts[1] == [4, 5, 6, 7, 20, 21, 22] and parts[3] == [12, 13, 14, 15]
    assert shard_queries(qs, 1, 0, eval_bsz=16) == qs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 3 + 2 * rank  # ranks hold different numbers of queries
    g = torch.Generator().manual_seed(rank)
    nms = torch.rand((n, 3, 5, 5), dtype=torch.float64, generator=g)
    cnt = torch.randint(0, 6, (n, 3), dtype=torch.int32, generator=g)
    qid = torch.arange(n, dtype=torch.int64) + 1000 * rank
    a, b, c = gather_predictions(nms, cnt, qid)
    torch.save({"nms": a, "cnt": b, "qid": c, "local": (nms, cnt, qid)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_predictions_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    want_nms = torch.cat([res[r]["local"][0] for r in range(world)])
    want_cnt = torch.cat([res[r]["local"][1] for r in range(world)])
    want_qid = torch.cat([res[r]["local"][2] for r in range(world)])
    for r in range(world):  # every rank ends up with every rank's predictions, in rank order, padding trimmed
        assert torch.equal(res[r]["nms"], want_nms)
        assert torch.equal(res[r]["cnt"], want_cnt)
        assert torch.equal(res[r]["qid"], want_qid)


def test_gather_predictions_without_process_group_is_identity():
    nms, cnt = torch.rand(4, 3, 5, 5, dtype=torch.float64), torch.ones(4, 3, dtype=torch.int32)
    a, b = gather_predictions(nms, cnt)
    assert a is nms and b is cnt


def _counter_worker(rank, world, port, out_dir):
    """rank 1 received no videos (world > n_videos): it must still enter the collectives of reduce_counters."""
    from cone_b200.inference import new_counters
    from cone_b200.sharding import reduce_counters
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = new_counters("cpu")
    if rank == 0:
        c.hits += 2
        c.n_queries += 4
        c.top1_iou.append(torch.arange(12, dtype=torch.float64).reshape(4, 3))
    reduce_counters(c)
    torch.save({"hits": c.hits, "n": c.n_queries, "iou": torch.cat(c.top1_iou)}, os.path.join(out_dir, f"c{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reduce_counters_with_an_empty_rank_does_not_hang(tmp_path):
    world = 2
    mp.spawn(_counter_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(tmp_path / f"c{r}.pt")
        assert int(res["n"].item()) == 4 and int(res["hits"].sum().item()) == 2 * res["hits"].numel()
        assert torch.equal(res["iou"], torch.arange(12, dtype=torch.float64).reshape(4, 3))
