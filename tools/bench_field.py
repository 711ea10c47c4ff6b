"""Micro-benchmark of the radiance-field MLP kernels alone (CUDA events, inputs resident in HBM).
    python tools/bench_field.py [--n 1000000] [--modes bf16,bf16_fused]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--modes", default="bf16,bf16_fused")
    ap.add_argument("--bwd", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    n, n_img = a.n, 19
    x = (torch.rand(n, 3, device=dev) * 2 - 1)
    img = ((torch.arange(n, device=dev) // 127) % n_img)[:, None] if os.environ.get("REAL_IMG", "1") == "1" else torch.randint(0, n_img, (n, 1), device=dev)
    for mode in a.modes.split(","):
        m = EONerfMLP(n_img, radiometric_normalization=True, precision=mode).to(dev)
        e = m._engine()
        e.prepared()
        for dens, flop in ((False, 1345280.0), (True, 982528.0)):
            for keep in (True, False):
                if not keep and mode != "bf16_fused":
                    continue
                ms = timeit(lambda: e.fwd(n, dens, x=x, img_idx=None if dens else img, keep=keep))
                print(f"{mode:11s} fwd density_only={int(dens)} keep={int(keep)}: {ms:8.3f} ms  {n * flop / ms / 1e9:8.1f} TFLOP/s", flush=True)
            if a.bwd:
                f = e.fwd(n, dens, x=x, img_idx=None if dens else img)
                gs = torch.randn(n, device=dev)
                g3 = torch.randn(n, 3, device=dev)
                flat, views, gstruct = e.new_grads()
                ms = timeit(lambda: e.bwd(n, dens, f, g_sigma=gs, g_rgb=None if dens else g3, g_ts=None if dens else gs,
                                          g_tb=None if dens else gs, grads_struct=gstruct, want_gx=dens))
                print(f"{mode:11s} bwd density_only={int(dens)}: {ms:8.3f} ms  {2 * n * flop / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
