"""Per-CTA wall time of the fused forward / backward-chain kernels (library built with -DEONERF_TIMING): is the static item
assignment balanced, do some SMs run slower?   python tools/fused_cta_times.py [n]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200 import _capi as K  # noqa: E402
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 980_794
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(n, 3, device=dev) * 2 - 1
img = ((torch.arange(n, device=dev) // 127) % 19)[:, None]
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
lib = C.CDLL(K.LIB_PATH)
gs, g3 = torch.randn(n, device=dev), torch.randn(n, 3, device=dev)
flat, views, gstruct, direct = e.grads_for_backward()
for _ in range(3):
    f = e.fwd(n, False, x=x, img_idx=img, keep=True)
    e.bwd(n, False, f, g_sigma=gs, g_rgb=g3, g_ts=gs, g_tb=gs, grads_struct=gstruct)
torch.cuda.synchronize()
for name in ("fwd", "bwd"):
    out = (C.c_ulonglong * 512)()
    getattr(lib, f"eonerf_debug_cta_time_{name}")(out)
    t = list(out)
    start, end = t[:148], t[256:256 + 148]
    t0 = min(start)
    dur = sorted((b - a) / 1e3 for a, b in zip(start, end))
    items = (n + 511) // 512
    print(f"{name}: {items} items over 74 pairs ({items / 74:.2f} per pair); kernel span {(max(end) - t0) / 1e3:.1f} us; CTA time min {dur[0]:.1f} / "
          f"median {dur[74]:.1f} / max {dur[-1]:.1f} us; exit times of the pairs (us):")
    ex = [(end[2 * p] - t0) / 1e3 for p in range(74)]
    print("  " + " ".join(f"{v:.0f}" for v in ex))
