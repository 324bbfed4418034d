"""GroupNorm alone at the SVD-XT level-0 shape (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

F_, HW, C = 25, 9216, 320
M = F_ * HW
ws = torch.zeros(native.groupnorm_workspace_bytes(F_, HW) // 4 + 1, dtype=torch.float32, device="cuda")
xs = [torch.randn(M, C, device="cuda", dtype=torch.float16) for _ in range(3)]
out = torch.empty(M, C, device="cuda", dtype=torch.float16)
g, b = torch.randn(C, device="cuda").half(), torch.randn(C, device="cuda").half()
for i in range(6):
    native.groupnorm_silu(out, xs[i % 3], g, b, n_img=F_, HW=HW, eps=1e-5, silu=True, frames_per_stat=1, workspace=ws)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out).all()))
