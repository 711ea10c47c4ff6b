import os, sys, torch
sys.path.insert(0, "/root/repo")
from eonerf_code_b200.datasets.synthetic import make_pinhole_rays
from eonerf_code_b200.nerfacc_compat import OccGridEstimator
from eonerf_code_b200.radiance_fields import VanillaNeRFRadianceField
from eonerf_code_b200.vanilla_rendering import Rays, render_image_with_occgrid
dev = torch.device("cuda:0")
torch.manual_seed(42)
vm = VanillaNeRFRadianceField(precision="bf16_fused").to(dev).train()
est = OccGridEstimator(roi_aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], resolution=64, levels=1).to(dev)
opt = torch.optim.Adam(vm.parameters(), lr=5e-4)
vb = [tuple(t.to(dev) for t in make_pinhole_rays(4096, seed=77 + i)) for i in range(2)]
bk = torch.ones(3, device=dev)
for i in range(12):
    o, d, px = vb[i % 2]
    rgb, acc, depth, n = render_image_with_occgrid(vm, est, Rays(o, d), near_plane=0.0, render_step_size=5e-3, render_bkgd=bk)
    a1 = torch.cuda.memory_allocated() / 2**30
    loss = torch.nn.functional.smooth_l1_loss(rgb, px)
    opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    print(i, int(n), f"alloc after fwd {a1:.1f} GiB, after step {torch.cuda.memory_allocated()/2**30:.1f}, peak {torch.cuda.max_memory_allocated()/2**30:.1f}, reserved {torch.cuda.memory_reserved()/2**30:.1f}")
