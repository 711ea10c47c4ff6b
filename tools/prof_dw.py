"""The grouped dW GEMM alone (torch profiler): backward of the fused field on n samples, kernel times per call.
    EONERF_TN_DBG bits: 1 no red.global tail, 2 no bias-gradient side job, 4 no MMAs;  EONERF_SIDE_STREAM=0: head kernels in line."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(n, 3, device=dev) * 2 - 1
img = ((torch.arange(n, device=dev) // 127) % 19)[:, None]
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
f = e.fwd(n, False, x=x, img_idx=img, keep=True)
gs, g3 = torch.randn(n, device=dev), torch.randn(n, 3, device=dev)
flat, views, gstruct, direct = e.grads_for_backward()
for _ in range(2):
    e.bwd(n, False, f, g_sigma=gs, g_rgb=g3, g_ts=gs, g_tb=gs, grads_struct=gstruct)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        e.bwd(n, False, f, g_sigma=gs, g_rgb=g3, g_ts=gs, g_tb=gs, grads_struct=gstruct)
        if os.environ.get("COOL"):
            torch.cuda.synchronize()
            time.sleep(float(os.environ["COOL"]))
    torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda ev: -ev.device_time_total)[:6]:
    print(f"{ev.device_time_total / 4:9.1f} us/call {ev.count // 4:3d}x  {ev.key[:90]}")
