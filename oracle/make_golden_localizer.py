"""Freeze outputs of the reference's single-video front end (`run_on_video/cone_localizator.py`) as
tests/golden/localizer.json.   Run in the build container:   python -m oracle.make_golden_localizer
Inputs are regenerated from seeds (`localizer_case`); the fixture holds the reference's OUTPUTS only.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cone_b200.config import EGO4D, MAD512  # noqa: E402
from cone_b200.weights import init_state_dict  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "localizer.json")

# name -> (config, weight seed, n_frames, n_tokens, data seed)
CASES = {
    "ego4d_900": (EGO4D.replace(clip_length=0.5333), 6, 1200, 11, 31),  # the demo's own constants (cone_localizator.py:12-37)
    "ego4d_ragged_end": (EGO4D.replace(clip_length=0.5333), 7, 1013, 20, 32),
    "mad512_4000": (MAD512, 8, 4000, 17, 33),
}


def localizer_case(cfg, n_frames, n_tokens, seed):
    """Raw (un-normalised) CLIP-like features with a planted moment so that scores are not degenerate."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n_frames, cfg.v_feat_dim), dtype=np.float32) * 3.0
    tok = rng.standard_normal((n_tokens, cfg.t_feat_dim), dtype=np.float32) * 2.0
    cls = rng.standard_normal((cfg.v_feat_dim,), dtype=np.float32)
    f0 = int(rng.integers(0, n_frames - 40))
    v[f0:f0 + 30] += 2.0 * cls[None, :]
    return v, tok, cls


def main():
    assert R.reference_available(), "needs /root/reference"
    out = {}
    for name, (cfg, wseed, L, nt, seed) in CASES.items():
        sd = init_state_dict(cfg, wseed)
        v, tok, cls = localizer_case(cfg, L, nt, seed)
        moments, ranklist = R.run_reference_localizer(cfg, sd, v, tok, cls)
        out[name] = dict(moments=moments, ranklist=ranklist)
        print(name, moments)
    with open(GOLDEN, "w") as f:
        json.dump(out, f)
    print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
