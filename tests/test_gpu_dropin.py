"""INTEGRATION.md §1 executed verbatim: the UNMODIFIED reference driver (`cone/inference.py::eval_epoch`) runs with
cone_b200's operators substituted by assignment, on the GPU, and its outputs are compared with the driver's stock CPU
run on the same inputs.  Needs the reference's files (`/root/reference`, or `oracle/_ref` placed by
`python -m oracle.vendor_ref`, which `__graft_entry__.build()` runs in the build container)."""
import numpy as np
import pytest
import torch

from cone_b200.config import EGO4D
from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict
from helpers import ROUND_TOL, rows_close
from oracle import ref_harness as RH

DEV = "cuda:0"
needs_ref = pytest.mark.skipif(not RH.reference_available(), reason="reference files not placed (python -m oracle.vendor_ref)")


def _case():
    cfg = EGO4D.replace(eval_bsz=8, topk_window=6)
    return cfg, init_state_dict(cfg, 4), make_dataset(cfg, 3, [700, 455, 120], 4, seed=19)


@needs_ref
def test_vendored_reference_is_the_unmodified_reference():
    """CPU: the placed files hash-match their manifest, and the reference run from them reproduces the committed golden
    outputs of the build container's /root/reference (tests/golden/e2e_ego4d.*) through the same harness."""
    from oracle import vendor_ref
    from helpers import load_e2e
    if RH.REFERENCE_ROOT == vendor_ref.DEST:
        assert vendor_ref.verify()
    cfg, sd, ds, arrays, lists = load_e2e("e2e_ego4d")
    model = RH.build_reference_model(cfg, sd)
    got = RH.run_eval_epoch_files(cfg, ds, model, device="cpu")
    for q in ds.queries:
        assert got["ranklists"][q.query_id] == lists[q.query_id]["ranklist"]
        for mode in ("fusion", "proposal", "matching"):
            assert got[mode][q.query_id] == lists[q.query_id][mode], (q.query_id, mode)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_reference_driver_with_cone_b200_operators(precision):
    import cone_b200
    ref = RH.import_reference()
    cfg, sd, ds = _case()
    stock = RH.run_eval_epoch_files(cfg, ds, RH.build_reference_model(cfg, sd), device="cpu")
    saved = {k: getattr(ref, k) for k in ("build_model", "temporal_nms", "span_cxw_to_xx", "normalize_score")}
    try:
        # ---- INTEGRATION.md §1, verbatim
        ref.build_model = cone_b200.build_model
        ref.temporal_nms = cone_b200.temporal_nms
        ref.span_cxw_to_xx = cone_b200.span_cxw_to_xx
        ref.normalize_score = cone_b200.normalize_score
        # ---- what setup_model does with the factory (cone/inference.py:505-528)
        opt = RH.make_opt(cfg, "/tmp", "/tmp/val.jsonl", device=DEV)
        opt.precision = precision
        model, criterion = ref.build_model(opt)
        model.to(opt.device)
        model.load_state_dict(sd)
        model.eval()
        got = RH.run_eval_epoch_files(cfg, ds, model, device=DEV)
    finally:
        for k, v in saved.items():
            setattr(ref, k, v)
    span_tol = ROUND_TOL + (2e-5 if precision == "fp32" else 2e-3) * max(len(v) for v in ds.videos) * cfg.clip_length
    score_tol = ROUND_TOL if precision == "fp32" else 2.5e-3
    n_rank = sum(got["ranklists"][q.query_id] == stock["ranklists"][q.query_id] for q in ds.queries)
    assert n_rank == len(ds.queries), f"rank-lists equal for {n_rank} of {len(ds.queries)} queries"
    assert np.array_equal(np.asarray(got["recall"]), np.asarray(stock["recall"])), (got["recall"], stock["recall"])
    n_same = 0
    for q in ds.queries:
        a, b = got["fusion"][q.query_id], stock["fusion"][q.query_id]
        n_same += rows_close(a[:1], b[:1], span_tol, score_tol)
    print(f"[drop-in {precision}] top-1 fused prediction equal within rounding for {n_same} of {len(ds.queries)} queries")
    assert n_same >= len(ds.queries) - (0 if precision == "fp32" else 1)
