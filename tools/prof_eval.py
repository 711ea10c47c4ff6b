"""Short driver for ncu captures of the INFERENCE forms of the fused field kernel (what the render arm runs):
python tools/prof_eval.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(n, 3, device=dev) * 2 - 1
img = ((torch.arange(n, device=dev) // 127) % 19)[:, None]
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
e.prepared()
for it in range(3):
    e.fwd(n, False, x=x, img_idx=img, keep=False)
    e.fwd(n, True, x=x, keep=False)
torch.cuda.synchronize()
print("ok")
