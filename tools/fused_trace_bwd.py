"""Item-boundary timeline of CTA 0 of the fused backward chain (library built with -DEONERF_TIMING): when does the MMA issuer get each
slot / finish issuing a stage, when do the epilogue warps start an item and finish its prologue (head gradients -> first G)?"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200 import _capi as K  # noqa: E402
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

dens = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0")
n, n_img = 1_000_000, 19
x = torch.rand(n, 3, device=dev) * 2 - 1
img = ((torch.arange(n, device=dev) // 127) % n_img)[:, None]
m = EONerfMLP(n_img, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
lib = C.CDLL(K.LIB_PATH)
f = e.fwd(n, bool(dens), x=x, img_idx=None if dens else img)
gs, g3 = torch.randn(n, device=dev), torch.randn(n, 3, device=dev)
for _ in range(2):
    e.bwd(n, bool(dens), f, g_sigma=gs, g_rgb=None if dens else g3, g_ts=None if dens else gs, g_tb=None if dens else gs, grads_struct=None, want_gx=bool(dens))
out = (C.c_longlong * 2048)()
lib.eonerf_debug_trace_bwd(out)
mma, epi = list(out[:1024]), list(out[1024:])
n_st = 10 if dens else 12          # chain stages per item (no position gradients for the camera pass: 12; density-only with g_x: 10)
t0 = mma[0]
print(f"density_only={dens}; cycles relative to the first hand-over; per item: epilogue [item start, prologue done], MMA issuer [first slot got, last stage issued]")
for it in range(1, 9):
    es, ep = epi[2 * it] - t0, epi[2 * it + 1] - t0
    k0, k1 = 2 * (it * n_st * 2), 2 * ((it + 1) * n_st * 2) - 1
    print(f"item {it}: epi start {es:9d}  prologue done {ep:9d} (+{ep - es:5d}) | mma first-slot {mma[k0] - t0:9d}  last-issued {mma[k1] - t0:9d}  (item span {mma[k1] - mma[k0]:7d})")
