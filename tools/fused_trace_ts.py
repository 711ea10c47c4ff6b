"""Raw pipeline timeline of CTA 0 of the TMEM-operand inference kernel (library built with -DEONERF_TIMING): per MMA group when the
issuer reaches it, when its operands are ready, when it is issued; per accumulator half when the epilogue sees it, finishes, passes
the hand-over barrier.  Density-only program: 32 groups and 16 halves per item."""
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
m = EONerfMLP(n_img, radiometric_normalization=True, precision="bf16_fused").to(dev)
e = m._engine()
lib = C.CDLL(K.LIB_PATH)
e.fwd(n, True, x=x, keep=False)
e.fwd(n, True, x=x, keep=False)
out = (C.c_longlong * 2048)()
lib.eonerf_debug_trace_ts(out)
mma, epi = list(out[:1024]), list(out[1024:])
steps = [(0, 2)] + [(s, 4) for s in (1, 2, 3, 4)] + [(5, 6)] + [(6, 4), (7, 4)]
names = []
for s, k in steps:
    names += [f"s{s}.{i}" for i in range(k)]
t0 = mma[0]
print("issuer: group | reached  operands_ready  issued | wait  issue")
for k in range(64, 64 + 64):
    if 3 * k + 2 >= 1024:
        break
    a, b, c = (mma[3 * k + i] - t0 for i in range(3))
    print(f"{k:3d} {names[k % 32]:>5s} | {a:9d} {b:9d} {c:9d} | {b - a:6d} {c - b:5d}")
print("epilogue: half | acc_seen  work_done  barrier | work  sync")
for k in range(32, 32 + 32):
    a, b, c = (epi[3 * k + i] - t0 for i in range(3))
    print(f"{k:3d} s{(k % 16) // 2}.h{k % 2} | {a:9d} {b:9d} {c:9d} | {b - a:6d} {c - b:5d}")
