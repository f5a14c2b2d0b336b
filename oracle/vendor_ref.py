"""Recipe that makes the UNMODIFIED reference runnable where `/root/reference` does not exist (the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (houzhijian/CONE) is pure Python: there is nothing to compile, so
"building" it for the GPU box means placing the handful of modules the inference path imports — byte for byte, no edits
— under `oracle/_ref/` (git-ignored: reference sources never enter this repository's history; NOT gpurun-ignored: the
directory travels with the snapshot like the built `.so`).  `__graft_entry__.build()` runs this recipe whenever
`/root/reference` is present; on the GPU box only the already-placed files are used.

What is placed (the files on the hot path and its metric scripts):
    cone/{__init__, inference, model, transformer, position_encoding, span_utils, matcher, misc, ego4d_mad_dataloader,
          config}.py, utils/{basic_utils, temporal_nms, tensor_utils, model_utils}.py, standalone_eval/*.py,
    run_on_video/{__init__, cone_localizator, temporal_nms}.py
and a MANIFEST.json with the sha256 of every file (checked by `verify()`, so a modified copy is detected).

Used by: `oracle/ref_harness.py` (falls back to `oracle/_ref` when `/root/reference` is absent), `bench.py --impl
reference` (times the reference's own `eval_epoch` on the box's host cores) and `tests/test_gpu_dropin.py` (INTEGRATION.md
§1 executed verbatim).  Nothing under `cone_b200/` imports or reads it.

    python -m oracle.vendor_ref            # (re)place the files from /root/reference
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("CONE_REFERENCE_SOURCE", "/root/reference")
FILES = [
    "cone/__init__.py", "cone/inference.py", "cone/model.py", "cone/transformer.py", "cone/position_encoding.py",
    "cone/span_utils.py", "cone/matcher.py", "cone/misc.py", "cone/ego4d_mad_dataloader.py", "cone/config.py",
    "utils/basic_utils.py", "utils/temporal_nms.py", "utils/tensor_utils.py", "utils/model_utils.py",
    "standalone_eval/evaluate_ego4d_nlq.py", "standalone_eval/evaluate_mad.py",
    "standalone_eval/evaluate_pre_filtered_window.py",
    "run_on_video/__init__.py", "run_on_video/cone_localizator.py", "run_on_video/temporal_nms.py",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available() -> bool:
    return os.path.isfile(os.path.join(SOURCE, "cone", "inference.py"))


def vendor(verbose: bool = False) -> str:
    """Place the files (idempotent).  Raises if the reference tree is not there."""
    if not source_available():
        raise RuntimeError(f"reference tree not found at {SOURCE}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(DEST, rel)
        if not os.path.isfile(src):
            if rel.endswith("__init__.py"):  # namespace-style package in the reference: nothing to place
                continue
            raise RuntimeError(f"reference file missing: {src}")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "houzhijian/CONE (unmodified files)", "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"placed {len(manifest)} reference files under {DEST}")
    return DEST


def verify() -> bool:
    """True iff `oracle/_ref` holds every file of its manifest with the recorded hash."""
    mf = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isfile(mf):
        return False
    with open(mf) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    vendor(verbose=True)
    sys.exit(0 if verify() else 1)
