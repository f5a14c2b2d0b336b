"""Shared test helpers: golden loading and tolerant comparison of prediction lists."""
import json
import os

import numpy as np

from cone_b200.synth import make_dataset
from cone_b200.weights import init_state_dict
from oracle.make_golden import E2E_CASES, GOLDEN, dense_case  # noqa: F401  (no reference import at module load)

# north_star: "scores and spans must agree within ... 1e-5 in fp32" (relative to the O(1) scale of
# probabilities / normalised spans / cosines: |a-b| <= FP32_TOL * max(1, |b|)).
FP32_TOL = 1e-5
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
