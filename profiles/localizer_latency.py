#!/usr/bin/env python
"""Per-query latency of the single-video front end (cone_b200.localizer.CONELocalizator.predict_moment), eager launches
against CUDA-graph replay, host call to host result (includes the H2D of the query and the D2H of the moments).
    python profiles/localizer_latency.py [--precision fp32|tc] [--frames 900] [--queries 200]"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cone_b200 import _lib
from cone_b200.localizer import CONELocalizator, EGO4D_DEMO
from cone_b200.weights import init_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--frames", type=int, default=900)
ap.add_argument("--queries", type=int, default=200)
a = ap.parse_args()
cfg = EGO4D_DEMO
sd = init_state_dict(cfg, 0)
rng = np.random.default_rng(0)
v = rng.standard_normal((a.frames, cfg.v_feat_dim), dtype=np.float32)
qs = [(rng.standard_normal((int(rng.integers(4, cfg.max_q_l + 1)), cfg.t_feat_dim), dtype=np.float32),
       rng.standard_normal(cfg.v_feat_dim, dtype=np.float32)) for _ in range(a.queries)]
res = {"frames": a.frames, "queries": a.queries, "precision": a.precision}
for mode in ("eager", "graph"):
    loc = CONELocalizator(sd, device="cuda:0", cfg=cfg, precision=a.precision, use_cuda_graph=(mode == "graph"))
    t0 = time.perf_counter(); loc.set_video(v); torch.cuda.synchronize(); res[f"{mode}_set_video_ms"] = 1e3 * (time.perf_counter() - t0)
    for q in qs[:5]:
        loc.predict_moment(v, q)
    _lib.reset_launch_count()
    lat = []
    for q in qs:
        t0 = time.perf_counter(); m = loc.predict_moment(v, q); lat.append(1e3 * (time.perf_counter() - t0))
    res[f"{mode}_ms_median"] = float(np.median(lat)); res[f"{mode}_ms_p95"] = float(np.percentile(lat, 95))
    res[f"{mode}_launches_per_query"] = _lib.launch_count() / len(qs)
print(json.dumps(res))
