"""Where do the role threads of the fused forward kernel wait?  Needs a library built with
EONERF_EXTRA_NVCC_FLAGS=-DEONERF_TIMING (python -m eonerf_code_b200.build --force)."""
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
names = ["mma<-act_ready", "mma<-weights", "producer<-slot", "epi<-acc_full", "epi start barrier", "epi end barrier", "mma thread total",
         "epi chunks", "epi fences", "epi store-read wait", "epi barrier(t0)", "epi post-barrier", "epi total(t0)"]
for dens in (False, True):
    for keep in (True, False):
        e.fwd(n, dens, x=x, img_idx=None if dens else img, keep=keep)
        out = (C.c_ulonglong * 16)()
        lib.eonerf_debug_timing(out, 1)
        e.fwd(n, dens, x=x, img_idx=None if dens else img, keep=keep)
        lib.eonerf_debug_timing(out, 1)
        tot = out[6]
        print(f"density_only={int(dens)} keep={int(keep)}: total {tot} cycles; " + ", ".join(f"{nm} {100 * out[i] / tot:.1f}%" for i, nm in enumerate(names) if i != 6))
