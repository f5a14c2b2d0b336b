"""Stand-alone bring-up check of the fused encoder tail: python profiles/enc_tail_check.py M cta_group
(run under `timeout`: a protocol bug in a warp-specialised kernel shows up as a hang, not as an error)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from cone_b200.config import EGO4D  # noqa: E402
from cone_b200.engine import ConeEngine  # noqa: E402
from cone_b200.weights import init_state_dict  # noqa: E402
from test_gpu_enc_tail import run_case  # noqa: E402

M, cg = int(sys.argv[1]), int(sys.argv[2])
sd = init_state_dict(EGO4D, 7)
eng = ConeEngine(EGO4D, sd, device="cuda:0", precision="tc", workspace_bytes=1 << 30)
emu, exact, rms = run_case(eng, {k: v.cpu() for k, v in sd.items()}, M, cg, layer=1)
torch.cuda.synchronize()
print(f"enc_tail M={M} cg={cg}: max err vs emulated {emu:.3e}, vs exact {exact:.3e} (rms {rms:.3e})", flush=True)
