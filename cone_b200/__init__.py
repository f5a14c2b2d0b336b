"""cone_b200 — CONE's coarse-to-fine long-video grounding inference path on B200 (sm_100a).

The arithmetic lives in `libcone_b200.so` (hand-written CUDA behind the C ABI of `include/cone_b200.h`); this
package is the host-side mirror of the reference's operator surface.  There is no CPU fallback."""
from . import config, inference, ingest, sharding
from ._lib import ConeError
from .config import EGO4D, MAD512, MAD768, ConeConfig
from .engine import ConeEngine, GroundingOutput, QueryBatch, pack_queries
from .localizer import CONELocalizator
from .model import CONE, build_model
from .ops import compute_window_ranklist, normalize_score, span_cxw_to_xx, temporal_nms

__all__ = ["config", "inference", "ingest", "sharding", "CONELocalizator", "ConeError", "ConeConfig", "EGO4D", "MAD512", "MAD768", "ConeEngine",
           "GroundingOutput", "QueryBatch", "pack_queries", "CONE", "build_model", "compute_window_ranklist",
           "normalize_score", "span_cxw_to_xx", "temporal_nms"]
