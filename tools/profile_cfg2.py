"""Where the cfg2 (vanilla NeRF, 4096 rays) step goes: torch.profiler kernel table of 5 steps + wall/device time per step."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.datasets.synthetic import make_pinhole_rays  # noqa: E402
from eonerf_code_b200.nerfacc_compat import OccGridEstimator  # noqa: E402
from eonerf_code_b200.radiance_fields import VanillaNeRFRadianceField  # noqa: E402
from eonerf_code_b200.vanilla_rendering import Rays, render_image_with_occgrid  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)
vm = VanillaNeRFRadianceField(precision="bf16_fused").to(dev).train()
est = OccGridEstimator(roi_aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], resolution=64, levels=1).to(dev)
opt = torch.optim.Adam(vm.parameters(), lr=5e-4)
vb = [tuple(t.to(dev) for t in make_pinhole_rays(4096, seed=77 + i)) for i in range(2)]
bk = torch.ones(3, device=dev)


def step(i):
    o, d, px = vb[i % 2]
    rgb, acc, depth, n = render_image_with_occgrid(vm, est, Rays(o, d), near_plane=0.0, render_step_size=5e-3, render_bkgd=bk)
    loss = torch.nn.functional.smooth_l1_loss(rgb, px)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss, n


for i in range(3):
    step(i)
torch.cuda.synchronize()
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    loss, n = step(i)
e1.record()
torch.cuda.synchronize()
print(f"device {e0.elapsed_time(e1) / 5:.2f} ms/step, wall {(time.time() - t0) * 200:.2f} ms/step, samples {int(n)}")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
