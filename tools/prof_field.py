"""Short driver for ncu captures of the fused field kernels: python tools/prof_field.py [fwd|bwd|both] [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(n, 3, device=dev) * 2 - 1
img = torch.randint(0, 19, (n, 1), device=dev)
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
e.prepared()
gs, g3 = torch.randn(n, device=dev), torch.randn(n, 3, device=dev)
flat, views, gstruct = e.new_grads()
for it in range(3):
    f = e.fwd(n, False, x=x, img_idx=img, keep=True)
    if what in ("bwd", "both"):
        e.bwd(n, False, f, g_sigma=gs, g_rgb=g3, g_ts=gs, g_tb=gs, grads_struct=gstruct)
torch.cuda.synchronize()
print("ok")
