"""Live weights: `cone_weights_update` rewrites a handle in place — results must equal those of a
freshly created handle bit for bit, in both precisions, and a CUDA graph captured before the update must see the new
weights (device pointers are stable)."""
import numpy as np
import pytest
import torch

from cone_b200.config import EGO4D
from cone_b200.engine import ConeEngine
from cone_b200.inference import run_step, stage_step
from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _outputs(eng, step):
    o = run_step(eng, step)
    torch.cuda.synchronize()
    return [t.clone() for t in (o.ranklist, o.pred_spans, o.prob_fg, o.match, o.nms, o.nms_count)]


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_update_equals_fresh_handle(precision):
    cfg = EGO4D.replace(eval_bsz=4)
    sd_a, sd_b = init_state_dict(cfg, 11), init_state_dict(cfg, 12)
    ds = make_dataset(cfg, 2, [400, 250], [3, 3], seed=9)
    step = stage_step(cfg, ds.videos, ds.queries, [0, 1])
    eng = ConeEngine(cfg, sd_a, device=DEV, precision=precision, workspace_bytes=1 << 30)
    a1 = _outputs(eng, step)
    eng.load_state_dict(sd_b)  # in place
    b_live = _outputs(eng, step)
    b_fresh = _outputs(ConeEngine(cfg, sd_b, device=DEV, precision=precision, workspace_bytes=1 << 30), step)
    for x, y in zip(b_live, b_fresh):
        assert torch.equal(x, y) or (torch.isnan(x) == torch.isnan(y)).all() and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))
    assert not torch.equal(a1[1], b_live[1])  # the weights really changed
    eng.load_state_dict(sd_a)
    for x, y in zip(_outputs(eng, step), a1):
        assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))
    with pytest.raises(Exception):
        eng.load_state_dict({k: v for k, v in list(sd_a.items())[:-1]})


def test_graph_captured_before_update_sees_new_weights():
    from cone_b200.localizer import CONELocalizator, EGO4D_DEMO
    cfg = EGO4D_DEMO
    sd_a, sd_b = init_state_dict(cfg, 13), init_state_dict(cfg, 14)
    rng = np.random.default_rng(5)
    v = rng.standard_normal((700, cfg.v_feat_dim), dtype=np.float32)
    tok = rng.standard_normal((9, cfg.t_feat_dim), dtype=np.float32)
    cls = v[321] + 0.2 * rng.standard_normal(cfg.v_feat_dim).astype(np.float32)
    loc = CONELocalizator(sd_a, device=DEV, cfg=cfg, precision="tc", use_cuda_graph=True)
    m_a = loc.predict_moment(v, (tok, cls))  # captures the graph
    graph = loc._graph
    loc.localizator.load_state_dict(sd_b)
    loc.set_video(v)  # per-video tensors depend on the weights: recomputed into the buffers the graph reads
    assert loc._graph is graph and graph is not None  # not re-captured: pointers are stable across the update
    m_b = loc.predict_moment(v, (tok, cls))
    fresh = CONELocalizator(sd_b, device=DEV, cfg=cfg, precision="tc", use_cuda_graph=False)
    assert m_b == fresh.predict_moment(v, (tok, cls)) and m_b != m_a
