"""Shared test helpers: golden loading and tolerant comparison of prediction lists."""
import json
import os

import numpy as np

from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict
from oracle.make_golden import E2E_CASES, GOLDEN, dense_case  # noqa: F401  (no reference import at module load)

# How north_star's tolerances are read, stated once: "scores and spans must agree within 1e-3 relative in bf16 or
# 1e-5 in fp32".  The compared quantities are the model's raw outputs — normalised spans (cx, w) in (0, 1), class
# probabilities in (0, 1), cosine matching scores in [-1, 1] — so "relative" is taken relative to their O(1) scale:
#     |a - b| <= tol * max(1, |b|)
# i.e. ABSOLUTE for |b| <= 1 (never looser than tol) and relative above 1 (spans in seconds after scaling).  A purely
# relative reading |a-b| <= tol*|b| is not meetable by ANY finite-precision implementation for values near 0 (a width
# of 1e-4 would need 1e-9 absolute) and is not what the reference's own 4-decimal rounding (inference.py:83) resolves.
FP32_TOL = 1e-5
TC_TOL = 1e-3  # reduced-precision (tensor-core, fp16 operands) mode: max over ALL compared values, no percentile form
# After the reference's 4-decimal rounding (inference.py:83) a 1e-7 difference can move a value by one
# unit in the 4th decimal; seconds are scaled by duration*clip_length (<= ~1e4 for MAD movies).
ROUND_TOL = 1.01e-4


def load_e2e(name):
    cfg, dkw, wseed, perturb = E2E_CASES[name]
    sd = init_state_dict(cfg, wseed, perturb=perturb)
    ds = make_dataset(cfg, **dkw)
    arrays = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        lists = json.load(f)
    return cfg, sd, ds, arrays, lists


def close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))))


def assert_close(a, b, tol, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e} > {tol:.1e}"


def rows_close(a, b, span_tol, score_tol=ROUND_TOL):
    """Two lists of [st, ed, score, ...] rows: same length, same order, values within tolerance."""
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if len(x) != len(y):
            return False
        if abs(x[0] - y[0]) > span_tol or abs(x[1] - y[1]) > span_tol:
            return False
        for u, v in zip(x[2:], y[2:]):
            if abs(u - v) > score_tol * max(1.0, abs(v)) * (2.5 if len(x) == 5 else 1.0):
                return False
    return True


def boundary_exempt(ref_spans, durations, eps=1e-4):
    """`start = floor(x1*dur)`, `end = ceil(x2*dur)` (model.py:187-192) turn a 1-ulp span
    difference into a +-1-frame pooling difference.  Returns a bool mask (k, nq) of proposals whose
    x1*dur or x2*dur lies within `eps` of an integer in the REFERENCE — only those may differ in the
    end-to-end matching score; everything else must meet the tolerance."""
    sp = np.asarray(ref_spans, dtype=np.float64)
    dur = np.asarray(durations, dtype=np.float64)[:, None]
    x1 = (sp[..., 0] - 0.5 * sp[..., 1]) * dur
    x2 = (sp[..., 0] + 0.5 * sp[..., 1]) * dur
    near = lambda x: np.abs(x - np.round(x)) < eps
    return near(x1) | near(x2)


def assert_match_close(got, ref, ref_spans, durations, tol, what="match"):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    ex = boundary_exempt(ref_spans, durations)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    bad = (err > tol) & ~ex
    assert not bad.any(), f"{what}: max err {err[~ex].max():.3e} > {tol:.1e} away from floor/ceil boundaries"
    assert ex.mean() <= 0.05, f"{what}: {ex.mean():.1%} of proposals sit on a floor/ceil boundary"
    return int(((err > tol) & ex).sum())


class Hatch:
    """A COUNTED exception to an otherwise exact comparison.  Every use is recorded with its reason, the total is printed
    (and appended to gpurun_out/hatches.log when that directory exists) and bounded by `limit` — so a regression cannot
    hide behind a silent `continue`."""

    def __init__(self, test: str, what: str, limit: int):
        self.test, self.what, self.limit, self.uses = test, what, limit, []

    def use(self, detail: str) -> None:
        self.uses.append(detail)

    def close(self, total: int) -> int:
        line = f"[hatch] {self.test}: {self.what}: used {len(self.uses)} of {total} (limit {self.limit})"
        if self.uses:
            line += " -> " + "; ".join(self.uses[:4])
        print(line)
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        if os.path.isdir(d):
            with open(os.path.join(d, "hatches.log"), "a") as f:
                f.write(line + "\n")
        assert len(self.uses) <= self.limit, line
        return len(self.uses)


def ranklist_near_tie(win_scores, got, want, k, rel=1e-5):
    """CPU-fp32 and GPU-fp32 frame scores differ in summation order (~1e-7), so two DIFFERENT windows
    whose reference scores lie within `rel` of each other may swap.  True iff every position < k where `got` and `want`
    disagree is such a swap: the reference's scores of got[p] and want[p] are within rel * max(1, |s|).  Exact ties
    (shared frame, equal stored score) must come out in index order and never pass as a "near-tie" with themselves
    swapped, because both sides break exact ties by ascending window id."""
    s = np.asarray(win_scores, dtype=np.float64)
    for p in range(min(k, len(want))):
        if p >= len(got):
            return False
        if got[p] != want[p]:
            a, b = s[got[p]], s[want[p]]
            if not (0 < abs(a - b) <= rel * max(1.0, abs(b))):
                return False
    return True


def oracle_window_scores(sd, cfg, ds, q):
    """The reference's window scores of one query (stage 0 + 1 of the oracle), for the near-tie audit."""
    import torch
    from oracle import cone_oracle as O
    with torch.no_grad():
        ctx = O.stage0_video_context(sd, torch.from_numpy(O.l2_normalize_np(ds.videos[q.video_idx])))
        _, fs = O.stage1_ranklist(ctx, torch.from_numpy(O.l2_normalize_np(q.cls)), cfg.max_v_l)
        return O.window_scores(fs, cfg.max_v_l).numpy()
