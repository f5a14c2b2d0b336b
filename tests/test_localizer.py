"""Single-video front end: the oracle's restatement of `CONELocalizator.predict_moment` against the
golden outputs of the reference class (CPU), and the CUDA path (eager and CUDA-graph replay) against both (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from cone_b200.weights import init_state_dict
from oracle import cone_oracle as O
from oracle.make_golden_localizer import CASES, localizer_case
from helpers import GOLDEN

DEV = "cuda:0"


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "localizer.json")) as f:
        return json.load(f)


def _moments_close(got, want, span_tol=2e-4, score_tol=5e-4):
    assert len(got) == len(want), (got, want)
    for a, b in zip(got, want):
        assert abs(a[0] - b[0]) <= span_tol and abs(a[1] - b[1]) <= span_tol, (a, b)
        assert abs(a[2] - b[2]) <= score_tol, (a, b)  # min-max fusion of 4-decimal scores amplifies one rounding step


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_predict_moment_matches_reference(golden, name):
    cfg, wseed, L, nt, seed = CASES[name]
    sd = init_state_dict(cfg, wseed)
    v, tok, cls = localizer_case(cfg, L, nt, seed)
    res = O.predict_moment(sd, cfg, torch.from_numpy(v), torch.from_numpy(tok), torch.from_numpy(cls))
    assert res["ranklist"] == golden[name]["ranklist"]
    _moments_close(res["moments"], golden[name]["moments"])


def test_oracle_predict_moment_refuses_too_many_tokens():
    cfg, wseed, L, nt, seed = CASES["ego4d_900"]
    v, tok, cls = localizer_case(cfg, 200, cfg.max_q_l + 1, 1)
    with pytest.raises(ValueError):
        O.predict_moment(init_state_dict(cfg, wseed), cfg, torch.from_numpy(v), torch.from_numpy(tok), torch.from_numpy(cls))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_localizer_matches_reference_and_oracle(golden, name):
    from cone_b200.localizer import CONELocalizator
    cfg, wseed, L, nt, seed = CASES[name]
    cfg = cfg.replace(max_before_nms=100, nms_thd=0.5, max_after_nms=5)
    sd = init_state_dict(cfg, wseed)
    v, tok, cls = localizer_case(cfg, L, nt, seed)
    loc = CONELocalizator(sd, device=DEV, cfg=cfg, use_cuda_graph=False)
    got = loc.predict_moment(v, (tok, cls))
    _moments_close(got, golden[name]["moments"])
    # the intermediate results against the oracle: rank-list bit-exact, raw scores within 1e-5
    ora = O.predict_moment(sd, cfg, torch.from_numpy(v), torch.from_numpy(tok), torch.from_numpy(cls))
    qb = loc._query_batch(L, torch.from_numpy(tok), torch.from_numpy(cls)).to(DEV)
    out = loc._run(qb)
    assert [int(x) for x in out.ranklist[0].cpu().tolist() if x >= 0] == ora["ranklist"] == golden[name]["ranklist"]
    k = len(ora["windows"])
    assert [(int(s), int(n)) for s, n in zip(out.win_start[0, :k].cpu(), out.win_len[0, :k].cpu())] == ora["windows"]
    for key, ref in (("pred_spans", ora["pred_spans"]), ("prob_fg", ora["prob_fg"]), ("match", ora["match"])):
        g = getattr(out, key)[0, :k].cpu().numpy()
        assert np.abs(g - ref).max() <= 1e-5, key
    # operator-level entry: compute_window_ranklist(adapter_video_feats, text_cls_feat)
    vn = torch.nn.functional.normalize(torch.from_numpy(v), dim=-1, eps=1e-5)
    adapted = O.adapter(sd, vn)
    assert loc.compute_window_ranklist(adapted, cls) == golden[name]["ranklist"]


@pytest.mark.gpu
def test_localizer_cuda_graph_replay_equals_eager():
    from cone_b200.localizer import CONELocalizator
    cfg, wseed, L, nt, seed = CASES["ego4d_900"]
    sd = init_state_dict(cfg, wseed)
    v, _, _ = localizer_case(cfg, L, nt, seed)
    eager = CONELocalizator(sd, device=DEV, cfg=cfg.replace(max_before_nms=100), use_cuda_graph=False)
    graph = CONELocalizator(sd, device=DEV, cfg=cfg.replace(max_before_nms=100), use_cuda_graph=True)
    rng = np.random.default_rng(3)
    for i, nt_i in enumerate((11, 20, 4, 7)):  # different token counts through the same captured graph
        tok = rng.standard_normal((nt_i, cfg.t_feat_dim), dtype=np.float32)
        cls = v[100 * i + 50] + 0.3 * rng.standard_normal(cfg.v_feat_dim).astype(np.float32)
        a = eager.predict_moment(v, (tok, cls))
        b = graph.predict_moment(v, (tok, cls))
        assert a == b and len(a) >= 1
    with pytest.raises(ValueError):
        graph.predict_moment(v, (rng.standard_normal((cfg.max_q_l + 1, cfg.t_feat_dim), dtype=np.float32), v[0]))


def test_video_cache_key_is_not_the_object_id():
    """CPython reuses ids of freed temporaries; the cache key must tell two such videos apart and must see
    in-place edits (numpy fingerprint, torch version counter)."""
    from cone_b200.localizer import _video_key, _as_f32
    a = np.zeros((5, 4), dtype=np.float32)
    ka = _video_key(a, _as_f32(a))
    b = np.ones((5, 4), dtype=np.float32)
    assert _video_key(b, _as_f32(b)) != ka
    a[4] = 7.0  # in-place edit of a sampled row
    assert _video_key(a, _as_f32(a)) != ka
    t = torch.zeros((6, 4))
    kt = _video_key(t, _as_f32(t))
    t.add_(1.0)
    assert _video_key(t, _as_f32(t)) != kt


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_localizer_second_temporary_video_is_not_served_from_the_cache(graph):
    """Two different videos passed as temporaries of the same shape (the reference driver's usage,
    run_on_video/run.py:54-57): each call must be computed on its own video."""
    from cone_b200.localizer import CONELocalizator
    cfg, wseed, L, nt, seed = CASES["ego4d_900"]
    sd = init_state_dict(cfg, wseed)
    v1, tok, cls = localizer_case(cfg, L, nt, seed)
    v2, _, _ = localizer_case(cfg, L, nt, seed + 17)
    ref = CONELocalizator(sd, device=DEV, cfg=cfg, use_cuda_graph=False)
    want1 = ref.predict_moment(v1, (tok, cls))
    ref2 = CONELocalizator(sd, device=DEV, cfg=cfg, use_cuda_graph=False)
    want2 = ref2.predict_moment(v2, (tok, cls))
    assert want1 != want2
    loc = CONELocalizator(sd, device=DEV, cfg=cfg, use_cuda_graph=graph)
    got1 = loc.predict_moment(v1.copy(), (tok, cls))  # temporaries: freed right after each call
    got2 = loc.predict_moment(v2.copy(), (tok, cls))
    assert got1 == want1 and got2 == want2
    buf = v1.copy()
    assert loc.predict_moment(buf, (tok, cls)) == want1
    buf[:] = v2  # in-place edit of the same array
    assert loc.predict_moment(buf, (tok, cls)) == want2
