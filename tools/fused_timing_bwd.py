"""Wait-time breakdown of the fused backward kernel (library built with -DEONERF_TIMING); see tools/fused_timing.py."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200 import _capi as K  # noqa: E402
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

dev = torch.device("cuda:0")
n, n_img = 1_000_000, 19
x = torch.rand(n, 3, device=dev) * 2 - 1
img = torch.randint(0, n_img, (n, 1), device=dev)
m = EONerfMLP(n_img, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
lib = K.lib()
names = ["mma<-act_ready", "mma<-weights", "producer<-slot", "epi<-acc_full", "epi start barrier", "epi end barrier", "mma thread total"]
for dens in (False, True):
    f = e.fwd(n, dens, x=x, img_idx=None if dens else img)
    gs = torch.randn(n, device=dev)
    g3 = torch.randn(n, 3, device=dev)
    out = (C.c_ulonglong * 16)()
    for rep in range(2):
        lib.eonerf_debug_timing_bwd(out, 1)
        e.bwd(n, dens, f, g_sigma=gs, g_rgb=None if dens else g3, g_ts=None if dens else gs, g_tb=None if dens else gs,
              grads_struct=None, want_gx=dens)
        lib.eonerf_debug_timing_bwd(out, 1)
    tot = out[6]
    print(f"bwd density_only={int(dens)}: total {tot} cycles; " + ", ".join(f"{nm} {100 * out[i] / tot:.1f}%" for i, nm in enumerate(names[:6])))
