#!/usr/bin/env python
"""Benchmark of the CONE coarse-to-fine grounding path on B200 (BASELINE.json metric: grounding queries/s,
MAD-shape, device-timed).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port) on host cores

A step = one pass of the whole hot path (stages 0-3: adapter + window pre-filter + Moment-DETR on the top-k
windows + proposal matching + fusion/NMS) over one synthetic MAD-shaped movie with all its queries.  Per-GPU
work is fixed as N grows (weak scaling): every rank owns `--movies` movies (its shard of the movie set); ranks
exchange nothing on the data path and all-gather the fixed-size per-query prediction blocks at the end.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=8)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="mad768", choices=["mad768", "mad512", "ego4d"])
    p.add_argument("--precision", default=os.environ.get("CONE_BENCH_PRECISION", "tc"), choices=["fp32", "tc"],
                   help="tc: tcgen05 fp16-operand / fp32-accumulate projections (1e-3 class); fp32: CUDA-core parity mode (1e-5)")
    p.add_argument("--movies", type=int, default=8, help="movies resident per GPU (cycled through by the steps)")
    p.add_argument("--queries-per-movie", type=int, default=640)
    p.add_argument("--videos-per-step", type=int, default=1,
                   help="videos batched into one step (short-clip configs: Ego4D clips are 900 frames with ~4.5 queries each)")
    p.add_argument("--frames", type=int, nargs=2, default=None, help="movie length range in frames")
    p.add_argument("--cpu-sample-queries", type=int, default=256)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true", help="skip the end-to-end loop (profiler passes only)")
    p.add_argument("--workspace-gb", type=float, default=24.0)
    p.add_argument("--seed", type=int, default=0)
    return p.parse_args()


def workload_name(cfg, args, frames):
    return (f"{cfg.name}: synthetic movies of {frames[0]}-{frames[1]} frames x {cfg.v_feat_dim}-d, "
            f"{args.queries_per_movie} queries/movie, window {cfg.max_v_l}, top-{cfg.topk_window} windows, "
            f"nms {cfg.nms_thd}")


def frames_range(cfg, args):
    if args.frames:
        return tuple(args.frames)
    return (36000, 54000) if cfg.name.startswith("mad") else (900, 900)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t_begin, self.t_end = 0.0, float("inf")

    def wait_ready(self, timeout=8.0):
        """Block until the first sample has arrived: nvidia-smi's start-up (process spawn, NVML initialisation) must not
        fall into the timed region — on a fresh box it cost several ms per step of an 8-step run."""
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout and self.proc is not None and self.proc.poll() is None:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        """Samples taken DURING the timed region only (between mark_begin and mark_end)."""
        rows = [r for t, r in self.rows if self.t_begin <= t <= self.t_end + 0.03]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops_per_query(cfg):
    """transformer + heads FLOPs per query, counted as the reference computes them (text and
    video projections once per window)."""
    d, ff, S = cfg.hidden_dim, cfg.dim_feedforward, cfg.max_v_l + cfg.max_q_l
    nq = cfg.num_queries
    vid = cfg.max_v_l * (cfg.v_feat_dim * d + d * d) * 2
    txt = cfg.max_q_l * (cfg.t_feat_dim * d + d * d) * 2
    enc = cfg.enc_layers * (S * (4 * d * d + 2 * d * ff) * 2 + 2 * S * S * d * 2)
    dec = cfg.dec_layers * (S * 2 * d * d * 2 + nq * (6 * d * d + 2 * d * ff) * 2 + 2 * nq * S * d * 2 + 2 * nq * nq * d * 2)
    heads = nq * (2 * d * d + 2 * d + 2 * d) * 2
    return cfg.topk_window * (vid + txt + enc + dec + heads)


def cpu_oracle_sample(cfg, sd, ds, n_queries, threads):
    """The reference's CPU path (oracle port) on a bounded sample: the first movie with its first n queries."""
    import torch
    from oracle import cone_oracle as O
    torch.set_num_threads(threads)
    qs = [q for q in ds.queries if q.video_idx == 0][:n_queries]
    t0 = time.perf_counter()
    O.eval_pipeline(sd, cfg, ds.videos[:1], qs, collect_raw=False)
    dt = time.perf_counter() - t0
    return len(qs) / dt, dt, len(qs)


def run_reference(args):
    import torch
    from cone_b200.config import PRESETS
    from cone_b200.synth import make_dataset
    from cone_b200.weights import init_state_dict
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = PRESETS[args.config]
    fr = frames_range(cfg, args)
    threads = os.cpu_count() or 1
    sd = init_state_dict(cfg, args.seed)
    ds = make_dataset(cfg, 1, None, args.cpu_sample_queries, seed=args.seed, frames_range=fr)
    times = []
    for i in range(args.warmup + args.steps):
        qps, dt, n = cpu_oracle_sample(cfg, sd, ds, args.cpu_sample_queries, threads)
        if i >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = args.steps * args.cpu_sample_queries / total
    sample = (f"each step = stages 0-3 on 1 movie of {len(ds.videos[0])} frames with {args.cpu_sample_queries} of its "
              f"{args.queries_per_movie} queries, torch-CPU fp32, {threads} threads")
    line = {"impl": "reference", "metric": "grounding_queries_per_sec", "value": value, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, args, fr), "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cone_b200 import _lib
    from cone_b200.config import PRESETS
    from cone_b200.engine import ConeEngine
    from cone_b200.inference import stage_step
    from cone_b200.sharding import gather_predictions
    from cone_b200.synth import make_dataset
    from cone_b200.weights import init_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cone_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = PRESETS[args.config]
    fr = frames_range(cfg, args)
    sd = init_state_dict(cfg, args.seed)
    # every rank owns its own shard of the movie set (weak scaling): movie ids rank*M .. rank*M+M-1
    ds = make_dataset(cfg, args.movies, None, args.queries_per_movie, seed=args.seed + 1000 * rank, frames_range=fr,
                      id_offset=rank * args.movies)
    eng = ConeEngine(cfg, sd, device=dev, precision=args.precision, workspace_bytes=int(args.workspace_gb * (1 << 30)))
    vps = max(1, args.videos_per_step)
    groups = [list(range(v, min(v + vps, args.movies))) for v in range(0, args.movies, vps)]
    host_steps = [stage_step(cfg, ds.videos, ds.queries, g) for g in groups]
    dev_steps = [(s.frames.to(dev), s.qb.to(dev)) for s in host_steps]
    # Size the caching allocator once: the movies differ in length, so without this the first timed steps on a longer
    # movie than the warm-up saw call cudaMalloc (a device-synchronising call) between kernels of the timed region
    # (seen as 4-5 ms of idle GPU per step in some runs while the kernel times were unchanged).
    presize = torch.empty(int(4e9), dtype=torch.uint8, device=dev)
    del presize
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region: inputs resident in HBM (1.1 GB of movies cycled: larger than the 126 MB L2) ----
    clocks = ClockSampler(local)
    clocks.__enter__()  # started before the warm-up so that its start-up stays outside the timed region
    clocks.wait_ready()
    for i in range(args.warmup):
        eng.ground(*dev_steps[i % len(dev_steps)])
    barrier()
    lib = _lib.load()
    _lib.reset_launch_count()
    has_prof = hasattr(lib, "cone_profile_enable")
    if has_prof:
        lib.cone_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_queries = 0
    try:
        clocks.mark_begin()
        ev0.record()
        for i in range(args.steps):
            out = eng.ground(*dev_steps[i % len(dev_steps)])
            n_queries += out.nms_count.shape[0]
        ev1.record()
        barrier()
        clocks.mark_end()
    finally:
        clocks.__exit__(None, None, None)
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count()
    prof = None
    if has_prof:
        from cone_b200.engine import read_profile
        prof = read_profile()
        lib.cone_profile_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    nq_t = torch.tensor([n_queries], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nq_t, op=dist.ReduceOp.SUM)
    ms_max, nq_all = float(t.item()), float(nq_t.item())
    value = nq_all / (ms_max / 1e3)

    # ---- end to end: pinned host inputs -> H2D -> path -> D2H of the predictions (+ all-gather across ranks) ----
    # Every step's inputs are copied from pinned host memory inside the timed region and every step's result is
    # read back to the host; the copy of step i+1 is issued on a side stream so that it overlaps step i's kernels
    # (what a serving loop does), and the host reads are asynchronous into pinned buffers, fenced at the end.
    if args.no_e2e:
        if rank == 0:
            emit({"metric": "grounding_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world,
                  "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "note": "profiler pass: no e2e"})
        if world > 1:
            dist.destroy_process_group()
        return 0
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    # two device landing buffers for the frame features (the step being computed and the one being copied), allocated
    # once: a serving loop does not call the allocator per request
    max_frames = max(s.frames.shape[0] for s in host_steps)
    fbuf = [torch.empty((max_frames, cfg.v_feat_dim), dtype=torch.float32, device=dev) for _ in range(2)]
    consumed = [None, None]  # event recorded on the main stream when the kernels reading fbuf[slot] have been queued
    # ... and for the packed queries (same shapes in every step of this workload): no allocator call inside the timed
    # loop (two runs on fresh boxes lost half of their end-to-end rate to cudaMalloc on the copy stream)
    import dataclasses
    same_shapes = all(all(getattr(s.qb, f.name).shape == getattr(host_steps[0].qb, f.name).shape
                          for f in dataclasses.fields(s.qb) if isinstance(getattr(s.qb, f.name), torch.Tensor))
                      for s in host_steps)
    qbuf = [host_steps[0].qb.to(dev, non_blocking=False) for _ in range(2)] if same_shapes else None

    def prefetch(i):
        s = host_steps[i % len(host_steps)]
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            if consumed[slot] is not None:
                copy_stream.wait_event(consumed[slot])
            frames_d = fbuf[slot][: s.frames.shape[0]]
            frames_d.copy_(s.frames, non_blocking=True)
            if qbuf is None:
                qb = s.qb.to(dev)
            else:
                dst, upd = qbuf[slot], {}
                for f in dataclasses.fields(s.qb):
                    t = getattr(s.qb, f.name)
                    if isinstance(t, torch.Tensor):
                        getattr(dst, f.name).copy_(t, non_blocking=True)
                        upd[f.name] = getattr(dst, f.name)
                qb = dataclasses.replace(s.qb, **upd)  # host metadata of this step, device tensors of the slot
        return s, frames_d, qb

    host_out = []
    # pinned landing buffers for every step's result, allocated outside the timed region (cudaHostAlloc is slow)
    probe = eng.ground(*dev_steps[0])
    n_rows = probe.nms.shape[0] * world
    pinned = [(torch.empty((n_rows,) + tuple(probe.nms.shape[1:]), dtype=probe.nms.dtype, pin_memory=True),
               torch.empty((n_rows,) + tuple(probe.nms_count.shape[1:]), dtype=probe.nms_count.dtype, pin_memory=True))
              for _ in range(args.steps)]
    del probe
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    nxt = prefetch(0)
    for i in range(args.steps):
        s, frames_d, qb = nxt
        main.wait_stream(copy_stream)
        if qbuf is None:
            qb.record_stream(main)
        if i + 1 < args.steps:
            nxt = prefetch(i + 1)
        out = eng.ground(frames_d, qb)
        consumed[i % 2] = torch.cuda.Event()
        consumed[i % 2].record(main)
        if world > 1:
            nms, cnt = gather_predictions(out.nms, out.nms_count, equal_shards=True)
        else:
            nms, cnt = out.nms, out.nms_count
        nms_h, cnt_h = pinned[i]
        nms_h.copy_(nms, non_blocking=True)  # device->host read of the step's result
        cnt_h.copy_(cnt, non_blocking=True)
        host_out.append((nms_h, cnt_h))
        h2d += s.h2d_bytes()
        d2h += nms_h.numel() * 8 + cnt_h.numel() * 4
    barrier()
    e2e_s = time.perf_counter() - t0
    assert all(int(c.sum()) > 0 for _, c in host_out)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nq_all / float(te.item())

    if rank == 0:
        peaks = load_peaks()
        line = {"metric": "grounding_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16",
                "data": "synthetic",
                "config": {"workload": workload_name(cfg, args, fr), "movies_per_gpu": args.movies,
                           "queries_per_step": args.queries_per_movie * vps, "videos_per_step": vps, "precision": args.precision,
                           "l2": "inputs larger than L2: steps cycle through %d movies (%.2f GB) resident in HBM" %
                                 (args.movies, sum(v.nbytes for v in ds.videos) / 1e9),
                           "parallelism": f"movie-sharded x{world}, all-gather of per-query predictions"},
                "clocks": clocks.summary(), "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": d2h // args.steps}}
        flops_q = algorithmic_flops_per_query(cfg)
        line["roofline"] = roofline_entry(prof, peaks, flops_q, args, cfg)
        if prof:
            # proposal ranking (A9): algorithmic bytes = pooled rows x Dv x 4, from the last step's own spans
            sp, wl = out.pred_spans.float(), out.win_len.float()[:, :, None]
            st = torch.clamp(torch.floor((sp[..., 0] - 0.5 * sp[..., 1]) * wl), min=0)
            en = torch.minimum(torch.ceil((sp[..., 0] + 0.5 * sp[..., 1]) * wl), wl)
            pool_rows = float(torch.clamp(en - st, min=0).sum().item())
            if "span_pool" in prof:
                # every pooled row is counted once per proposal; the 5 proposals of a window and the 50 %-overlapping
                # windows share rows, so most of these reads are served by L2 and the figure can exceed the HBM peak
                prof["span_pool"]["bytes"] = pool_rows * cfg.v_feat_dim * 4.0 * args.steps
            # window pre-filter (A3): one (frame, query) score each; the rank-list kernel reads every score once
            n_scores = float(sum(host_steps[i % len(host_steps)].qb.total_scores for i in range(args.steps)))
            n_frames_t = float(sum(host_steps[i % len(host_steps)].frames.shape[0] for i in range(args.steps)))
            if "frame_scores" in prof:
                prof["frame_scores"]["flops"] = 2.0 * cfg.v_feat_dim * n_scores
                prof["frame_scores"]["bytes"] = 4.0 * (n_scores + cfg.v_feat_dim * (n_frames_t + nq_all / world))
            if "window_ranklist" in prof:
                prof["window_ranklist"]["bytes"] = 4.0 * n_scores
            line["stages"] = stage_table(prof, peaks, args.steps)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            qps, dt, n = cpu_oracle_sample(cfg, sd, ds, args.cpu_sample_queries, threads)
            line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": f"stages 0-3 on movie 0 ({len(ds.videos[0])} frames) with {n} of its "
                                              f"{args.queries_per_movie} queries, torch-CPU fp32 oracle port, {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def load_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) or the fallback B200_PROFILING.md states."""
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pk, src = {}, "fallback (B200_PROFILING.md: 6.65 TB/s copy, 1.59 PFLOP/s burst / ~1.4 sustained)"
    return {"hbm_gbs": float(pk.get("hbm_gbs", 6650.0)),
            "tflops": float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0))), "source": src}


def load_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same
    command (profiles/r01_traffic.json, written by profiles/summarize_ncu.py); None when not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return t.get(key)
    except Exception:
        return None


def roofline_entry(prof, peaks, flops_per_query, args, cfg):
    """Dominant kernel of the step: the dense-projection GEMM (largest share of the step).  It sits at the
    ridge: K = 256 / 1024 with N <= 1024 gives ~200 FLOP per activation byte against a machine balance of
    ~210, so both rooflines are reported; `bound` names the one that is closer to its peak.
    achieved = algorithmic FLOPs (bytes) of the launches / their summed duration (CUDA events around every
    launch, on the launching stream, inside the timed region)."""
    key = "gemm_tc" if args.precision == "tc" else "gemm_fp32"
    if not prof or key not in prof or not prof[key]["launches"]:
        return {"bound": "tensor", "achieved": None, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": None,
                "traffic": None, "note": "per-kernel profile unavailable"}
    p = prof[key]
    sec = p["ms"] * 1e-3
    tf = p["flops"] / sec / 1e12
    gbs = p["bytes"] / sec / 1e9
    total = max(sum(v["ms"] for v in prof.values()), 1e-9)
    ent = {"kernel": key, "launches": p["launches"], "avg_launch_ms": p["ms"] / p["launches"],
           "share_of_step": p["ms"] / total, "peak_source": peaks["source"],
           "tensor": {"achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tflops"]},
           "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                   "algorithmic_bytes_per_launch": p["bytes"] / p["launches"]},
           "traffic": load_traffic(key)}
    if key == "gemm_fp32":
        # parity mode: GEMMs on the fp32 FMA pipe (148 SMs x 128 lanes x 2 FLOP x 1.965 GHz nominal), not the tensor pipe
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        ent.pop("tensor")
        ent.update(bound="fp32", achieved=tf, peak=fp32_peak, unit="TFLOP/s", frac=tf / fp32_peak,
                   peak_source="nominal fp32 FMA rate (no measured fp32 peak in MEASURED_PEAKS.json)")
    elif gbs / peaks["hbm_gbs"] >= tf / peaks["tflops"]:
        ent.update(bound="hbm", achieved=gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=gbs / peaks["hbm_gbs"])
    else:
        ent.update(bound="tensor", achieved=tf, peak=peaks["tflops"], unit="TFLOP/s", frac=tf / peaks["tflops"])
    return ent


def stage_table(prof, peaks, steps):
    """Per kernel category: ms/step and achieved GB/s / TFLOP/s from the algorithmic work the launches declared."""
    out = {}
    for k, v in prof.items():
        if not v["launches"]:
            continue
        sec = v["ms"] * 1e-3
        e = {"ms": round(v["ms"] / steps, 3), "launches": v["launches"] // steps}
        if v["bytes"]:
            e["GBps"] = round(v["bytes"] / sec / 1e9, 1)
            e["hbm_frac"] = round(v["bytes"] / sec / 1e9 / peaks["hbm_gbs"], 3)
        if v["flops"]:
            e["TFLOPs"] = round(v["flops"] / sec / 1e12, 2)
        out[k] = e
    return out


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


if __name__ == "__main__":
    a = parse_args()
    # native libraries (NCCL's version banner, driver notices) write to fd 1: keep stdout for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
