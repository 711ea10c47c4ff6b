"""Per-CTA timeline of the grouped dW GEMM (library built with -DEONERF_TIMING): start, last load issued, exit (globaltimer, ns) for
each of the 148 persistent CTAs, next to the GEMM(s) its cost interval covers."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200 import _capi as K  # noqa: E402
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
for _ in range(3):
    e.bwd(n, False, f, g_sigma=gs, g_rgb=g3, g_ts=gs, g_tb=gs, grads_struct=gstruct)
torch.cuda.synchronize()
lib = C.CDLL(K.LIB_PATH)
out = (C.c_ulonglong * 768)()
lib.eonerf_debug_tnb_time(out)
t = list(out)
start, last, end = t[:148], t[256:256 + 148], t[512:512 + 148]
t0 = min(start)
# boxes per chunk of the GEMMs in launch order (field_fused_bwd.cu): T3 T2 T1 HD0 BOTT trunk7..1 (+ layer-5 encoding part) layer 0
per = [4, 4, 4, 8, 8, 8, 8, 8, 5, 8, 8, 8, 8, 5]
names = ["T3", "T2", "T1", "HD0", "BOTT", "L7", "L6", "L5", "L5enc", "L4", "L3", "L2", "L1", "L0"]
tot = sum(per)
print(f"kernel span {(max(end) - t0) / 1e3:.1f} us; CTA run time min {min(e_ - s for s, e_ in zip(start, end)) / 1e3:.1f} / median "
      f"{sorted(e_ - s for s, e_ in zip(start, end))[74] / 1e3:.1f} / max {max(e_ - s for s, e_ in zip(start, end)) / 1e3:.1f} us")
print("cta  start  last_load  exit (us)   GEMMs covered")
for b in range(148):
    lo, hi = tot * b / 148, tot * (b + 1) / 148
    cov, acc = [], 0
    for nm, p in zip(names, per):
        if acc < hi and acc + p > lo:
            cov.append(nm)
        acc += p
    if b % 4 == 0 or (end[b] - t0) > 0.97 * (max(end) - t0):
        print(f"{b:3d} {(start[b] - t0) / 1e3:7.1f} {(last[b] - t0) / 1e3:9.1f} {(end[b] - t0) / 1e3:8.1f}   {'+'.join(cov)}")
