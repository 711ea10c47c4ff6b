"""HBM roofline of the stand-alone sampling / compositing / shadow kernels (csrc/sampling.cu, csrc/render.cu) at the
65 536-ray batch of BASELINE configs[4] (n_samples=128: ~8 M kept samples, every array several times the 126 MB L2).

Algorithmic bytes per kept sample / per ray are SURVEY.md §8d's; achieved = bytes / CUDA-event time; peak from
MEASURED_PEAKS.json.    python tools/bench_render.py [--rays 65536] [--n 128]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eonerf_code_b200 import _capi as K  # noqa: E402
from eonerf_code_b200 import ops  # noqa: E402
from eonerf_code_b200.datasets.synthetic import make_rays  # noqa: E402


def timeit(fn, iters=10, warm=3):
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
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--n", type=int, default=128)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    K.require_device()
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    B, n = a.rays, a.n
    rays, ts_img, _ = make_rays(B, 19, seed=1, variant="inside")
    rays = rays.to(dev)
    u = torch.rand(B, n, device=dev)
    s = lambda: torch.cuda.current_stream().cuda_stream
    p = lambda t: None if t is None else t.data_ptr()

    ri, ts, te, ppr, offs, stats = ops.sample_compact(rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], u)
    P = int(stats[0])
    rows = []

    def report(name, ms, bytes_):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        rows.append((name, ms * 1e3, bytes_ / 1e6, gbs, gbs / peak))
        print(f"{name:28s} {ms * 1e3:9.1f} us  {bytes_ / 1e6:9.1f} MB algorithmic  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f} % of {peak:.0f}", flush=True)

    # the C entry point alone, outputs pre-allocated (the Python wrapper's allocations are not the kernel's time)
    sa_ = K.SampleArgs()
    sa_.origins, sa_.origins_stride, sa_.viewdirs, sa_.viewdirs_stride = rays.data_ptr(), 11, rays.data_ptr() + 12, 11
    sa_.near, sa_.near_stride = rays.data_ptr() + 24, 11
    zs = ops.z_steps_for(n, dev)
    sa_.u, sa_.z_steps, sa_.n_rays, sa_.n_samples = p(u), p(zs), B, n
    sa_.ray_indices, sa_.t_starts, sa_.t_ends, sa_.pts_per_ray, sa_.ray_offsets, sa_.stats = p(ri), p(ts), p(te), p(ppr), p(offs), p(stats)
    for one_pass in (True, False):
        scratch = torch.empty(K.lib().eonerf_sample_scratch_bytes(B) // 8, dtype=torch.int64, device=dev)
        sa_.scratch = p(scratch) if one_pass else None
        report("sample_compact (" + ("one pass" if one_pass else "3 kernels") + ")", timeit(lambda: K.call("sample_compact", sa_, s())),
               P * 16 + B * 28 + B * n * 4)
    sa_.scratch = p(scratch)
    K.call("sample_compact", sa_, s())

    f32 = lambda *sh: torch.rand(*sh, device=dev, dtype=torch.float32)
    z, sigma, alb, tsc, tb, amb = f32(P), f32(P) * 3, f32(P, 3), f32(P), f32(P) + 0.1, f32(B, 3)
    ops.set_last_t_end(te, offs)
    comp = torch.empty(B, K.COMP_COLS, device=dev)
    fa = K.CompositeFwdArgs(p(ts), p(te), p(z), p(sigma), p(alb), p(tsc), p(tb), p(amb), p(offs), B, P, ops.BETA_MIN, p(comp))
    report("composite_fwd", timeit(lambda: K.call("composite_fwd", fa, s())), P * 32 + B * 36)
    g_comp = f32(B, K.COMP_COLS)
    g_sigma, g_alb, g_ts, g_tb, g_amb = torch.empty(P, device=dev), torch.empty(P, 3, device=dev), torch.empty(P, device=dev), torch.empty(P, device=dev), torch.empty(B, 3, device=dev)
    ba = K.CompositeBwdArgs(p(ts), p(te), p(z), p(sigma), p(alb), p(tsc), p(tb), p(amb), p(offs), B, P, p(g_comp), p(g_sigma), p(g_alb),
                            p(g_ts), p(g_tb), p(g_amb))
    report("composite_bwd", timeit(lambda: K.call("composite_bwd", ba, s())), P * 56 + B * 36)

    geo = torch.empty(B, 1, device=dev)
    sa = K.ShadowFwdArgs(p(ts), p(te), p(sigma), p(offs), B, P, p(geo))
    report("shadow_fwd", timeit(lambda: K.call("shadow_fwd", sa, s())), P * 12 + B * 4)
    g_geo = f32(B, 1)
    sb = K.ShadowBwdArgs(p(ts), p(te), p(offs), B, P, p(geo), p(g_geo), p(g_sigma))
    report("shadow_bwd", timeit(lambda: K.call("shadow_bwd", sb, s())), P * 12 + B * 8)

    w, T, al = torch.empty(P, device=dev), torch.empty(P, device=dev), torch.empty(P, device=dev)
    wa = K.WeightsFwdArgs(p(ts), p(te), p(sigma), p(offs), B, P, p(w), p(T), p(al))
    report("weights_fwd (nerfacc op)", timeit(lambda: K.call("weights_fwd", wa, s())), P * 24)
    out3 = torch.empty(B, 3, device=dev)
    aa = K.AccumFwdArgs(p(w), p(alb), 3, p(offs), B, P, p(out3))
    report("accumulate_fwd C=3", timeit(lambda: K.call("accumulate_fwd", aa, s())), P * 16 + B * 12)
    print(json.dumps({"rays": B, "n_samples": n, "kept_samples": P, "peak_gbs": peak,
                      "kernels": [{"name": r[0], "us": r[1], "algorithmic_mb": r[2], "gbs": r[3], "frac": r[4]} for r in rows]}))


if __name__ == "__main__":
    main()
