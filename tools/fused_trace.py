"""Raw pipeline timeline of CTA 0 of the fused forward kernel (library built with -DEONERF_TIMING): when does the MMA issuer get
each slot, when has it issued the stage, when do the epilogue warps see the accumulator, finish the chunks, pass the barrier?"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200 import _capi as K  # noqa: E402
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402

keep = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
n, n_img = 1_000_000, 19
x = torch.rand(n, 3, device=dev) * 2 - 1
img = ((torch.arange(n, device=dev) // 127) % n_img)[:, None] if os.environ.get("REAL_IMG", "1") == "1" else torch.randint(0, n_img, (n, 1), device=dev)
m = EONerfMLP(n_img, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
lib = C.CDLL(K.LIB_PATH)
e.fwd(n, False, x=x, img_idx=img, keep=bool(keep))
e.fwd(n, False, x=x, img_idx=img, keep=bool(keep))
out = (C.c_longlong * 2048)()
lib.eonerf_debug_trace(out)
mma, epi = list(out[:1024]), list(out[1024:])
t0 = mma[0]
print(f"keep={keep}; times in cycles relative to the first hand-over; slot-stage k = (item, stage, slot) in issue order")
print(" k  stage slot | mma:got_slot  issued | epi:acc_seen chunks_done barrier | mma_busy  drain  bar")
for k in range(26, 26 + 60):
    a, b = mma[2 * k] - t0, mma[2 * k + 1] - t0
    c, d, f = epi[3 * k] - t0, epi[3 * k + 1] - t0, epi[3 * k + 2] - t0
    print(f"{k:3d}  {(k % 26) // 2:3d}  {k % 2:3d} | {a:9d} {b:9d} | {c:9d} {d:9d} {f:9d} | {b - a:6d} {d - c:6d} {f - d:5d}")
