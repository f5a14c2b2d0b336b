"""Metric counters: the oracle's restatement of the reference's metric scripts against the golden
outputs of those scripts (CPU), the device kernels against both (GPU, through the C ABI), and the cross-rank
reduction of the counters (gloo, world_size 2)."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cone_b200.config import EGO4D, MAD512
from cone_b200.inference import MODES, MetricCounters, evaluate_dataset, new_counters
from cone_b200.sharding import reduce_counters
from oracle import cone_oracle as O
from helpers import GOLDEN

DEV = "cuda:0"


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "metrics.json")) as f:
        return json.load(f)


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_recall_matches_reference_scripts(golden):
    for case in golden["cases"]:
        for m in range(3):
            pred = {it["query_id"]: it["lists"][m] for it in case["items"]}
            gt = {it["query_id"]: it["gt"] for it in case["items"]}
            mad = O.recall_at_k_iou(pred, gt, thresholds=case["thresholds"], topk=case["topk"])
            assert np.array_equal(mad.astype(np.float32), np.asarray(case["mad"][m], dtype=np.float32)), case["name"]
            rec, miou = O.recall_ego4d(pred, gt, thresholds=case["thresholds"], topk=case["topk"])
            assert np.array_equal(rec, np.asarray(case["ego4d"][m]["recall"])), case["name"]
            assert miou == case["ego4d"][m]["mIoU"], case["name"]


def test_oracle_window_recall_matches_reference_script(golden):
    for case in golden["window_cases"]:
        rec = O.window_recall(case["ranklists"], case["gt"], case["clip_length"], case["max_v_l"], case["topk"])
        assert np.array_equal(rec, np.asarray(case["recall"], dtype=np.float32)), case["name"]


def test_fixture_holds_threshold_edge_cases(golden):
    """IoU exactly at a threshold must not count (strict >): the fixture was built with such rows."""
    case = golden["cases"][1]
    n_edge = 0
    for it in case["items"]:
        g0, g1 = it["gt"]
        for rows in it["lists"]:
            n_edge += sum(1 for r in rows if r[0] == g0 and abs((r[1] - r[0]) - 2 * (g1 - g0)) < 1e-9)
    assert n_edge >= 10


# ------------------------------------------------------------------------------------------------ CPU: reduction
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = new_counters("cpu")
    c.hits += rank + 1
    c.window_hits += 10 * (rank + 1)
    c.n_queries += 3 + 2 * rank
    c.top1_iou.append(torch.full((3 + 2 * rank, 3), float(rank), dtype=torch.float64))
    reduce_counters(c)
    torch.save({"hits": c.hits, "wh": c.window_hits, "n": c.n_queries, "iou": c.top1_iou[0]},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_counters_gloo_world2(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = torch.load(tmp_path / f"r{r}.pt")
        assert int(got["n"]) == 8 and torch.all(got["hits"] == 3) and torch.all(got["wh"] == 30)
        assert got["iou"].shape == (8, 3) and float(got["iou"][:3].sum()) == 0.0 and float(got["iou"][3:].sum()) == 15.0


def test_counters_to_tables():
    c = MetricCounters((1, 5), (0.3, 0.5), (1,), torch.tensor([[[1, 0], [3, 2]]] * 3), torch.tensor([2]),
                       torch.tensor([7]), [torch.tensor([[0.5, 0.0, 1.0]] * 7, dtype=torch.float64)])
    assert c.recall_mad("fusion").dtype == np.float32
    assert np.array_equal(c.recall_mad("matching"), np.float32([[1, 0], [3, 2]]) / np.float32(7))
    rec, miou = c.recall_ego4d("fusion")
    assert rec.shape == (2, 2) and rec[1, 0] == 0.0 and rec[0, 1] == 3 / 7 and miou == 0.5
    assert c.window_recall()[0] == np.float32(2) / np.float32(7)


# ------------------------------------------------------------------------------------------------ GPU: kernels
def _pack_case(case, dev):
    items = case["items"]
    ma = case["max_after"]
    nms = torch.zeros((len(items), 3, ma, 5), dtype=torch.float64)
    cnt = torch.zeros((len(items), 3), dtype=torch.int32)
    gt = torch.zeros((len(items), 2), dtype=torch.float64)
    for i, it in enumerate(items):
        gt[i, 0], gt[i, 1] = float(it["gt"][0]), float(it["gt"][1])
        for m in range(3):
            rows = torch.tensor(it["lists"][m], dtype=torch.float64)
            # the stage-3 kernel orders rankings (fusion, proposal, matching) = MODES; the fixture's lists are just 3 lists
            nms[i, m, : len(rows)] = rows
            cnt[i, m] = len(rows)
    return nms.to(dev), cnt.to(dev), gt.to(dev)


@pytest.mark.gpu
def test_eval_recall_kernel_matches_reference_scripts(golden):
    from cone_b200.engine import ConeEngine
    from cone_b200.weights import init_state_dict
    eng = ConeEngine(EGO4D, init_state_dict(EGO4D, 0), device=DEV, workspace_bytes=1 << 20)
    for case in golden["cases"]:
        nms, cnt, gt = _pack_case(case, DEV)
        n = nms.shape[0]
        hits, _ = eng.eval_recall(nms, cnt, gt, case["topk"], case["thresholds"], "mad")
        for m in range(3):
            got = hits[m].cpu().numpy().astype(np.float32) / np.float32(n)
            assert np.array_equal(got, np.asarray(case["mad"][m], dtype=np.float32)), (case["name"], m)
        hits, top1 = eng.eval_recall(nms, cnt, gt, case["topk"], case["thresholds"], "ego4d", want_top1=True)
        for m in range(3):
            rec = hits[m].cpu().numpy().astype(np.float64).T / n
            assert np.array_equal(rec, np.asarray(case["ego4d"][m]["recall"])), (case["name"], m)
            assert float(np.mean(top1[:, m].cpu().numpy())) == pytest.approx(case["ego4d"][m]["mIoU"], rel=0, abs=1e-15)
        # accumulation: calling twice into the same counters doubles them
        h2, _ = eng.eval_recall(nms, cnt, gt, case["topk"], case["thresholds"], "mad", hits=hits.clone())
        h1, _ = eng.eval_recall(nms, cnt, gt, case["topk"], case["thresholds"], "mad")
        assert torch.equal(h2, hits + h1)


@pytest.mark.gpu
def test_eval_window_recall_kernel_matches_reference_script(golden):
    from cone_b200.engine import ConeEngine
    from cone_b200.weights import init_state_dict
    eng = ConeEngine(EGO4D, init_state_dict(EGO4D, 0), device=DEV, workspace_bytes=1 << 20)
    for case in golden["window_cases"]:
        qids = list(case["ranklists"])
        stride = max(len(v) for v in case["ranklists"].values())
        rl = torch.full((len(qids), stride), -1, dtype=torch.int32)
        for i, q in enumerate(qids):
            rl[i, : len(case["ranklists"][q])] = torch.tensor(case["ranklists"][q], dtype=torch.int32)
        gt = torch.tensor([case["gt"][q] for q in qids], dtype=torch.float64)
        cfg = EGO4D.replace(max_v_l=case["max_v_l"], clip_length=case["clip_length"])
        hits = eng.eval_window_recall(rl.to(DEV), gt.to(DEV), case["topk"], cfg=cfg)
        got = hits.cpu().numpy().astype(np.float32) / np.float32(len(qids))
        assert np.array_equal(got, np.asarray(case["recall"], dtype=np.float32)), case["name"]


@pytest.mark.gpu
def test_evaluate_dataset_counters_equal_host_metric():
    """End to end: counters accumulated on the device over several steps equal the host-side metric of the same
    predictions and the oracle's metric of the oracle's predictions."""
    from cone_b200.engine import ConeEngine
    from cone_b200.inference import ground_dataset, recall_at_k
    from cone_b200.synth import make_dataset
    from cone_b200.weights import init_state_dict
    cfg = MAD512.replace(eval_bsz=4)
    sd = init_state_dict(cfg, 5)
    ds = make_dataset(cfg, 3, [1500, 700, 260], [4, 4, 4], seed=21)
    gt = {q.query_id: q.timestamps for q in ds.queries}
    eng = ConeEngine(cfg, sd, device=DEV, precision="fp32", workspace_bytes=2 << 30)
    c = evaluate_dataset(eng, ds.videos, ds.queries, gt, "mad", topk=(1, 5), thresholds=(0.1, 0.3, 0.5),
                         max_frames_per_step=1600)  # 3 steps: 1 + 1 + 1 videos
    res = ground_dataset(eng, ds.videos, ds.queries, max_frames_per_step=1600)
    for mode in MODES:
        want = recall_at_k(res, gt, mode=mode, thresholds=(0.1, 0.3, 0.5), topk=(1, 5))
        assert np.allclose(c.recall_mad(mode), want, atol=1e-7), mode
    wr = O.window_recall({q: r["ranklist"] for q, r in res.items()}, gt, cfg.clip_length, cfg.max_v_l, c.window_topk)
    assert np.array_equal(c.window_recall(), wr)
    assert int(c.n_queries.item()) == len(ds.queries)
    c2 = evaluate_dataset(eng, ds.videos, ds.queries, gt, "ego4d", max_frames_per_step=1 << 20)
    rec, miou = c2.recall_ego4d("fusion")
    want, want_miou = O.recall_ego4d({q: r["fusion"] for q, r in res.items()}, gt)
    assert np.array_equal(rec, want) and miou == pytest.approx(want_miou, abs=1e-12)
