"""Where does the time of one eval chunk go?  (CUDA events around the phases of render_image, eval, static)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200 import ops, sat_rendering, _capi as K
from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
from eonerf_code_b200.datasets.synthetic import make_rays
from eonerf_code_b200.radiance_fields import EONerfMLP
dev = torch.device("cuda:0")
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev).eval()
B, n = 131072, 128
rays, ts, _ = make_rays(B, 19, seed=7, eval_mode=True)
rays, ts = rays.to(dev), ts.to(dev)
sat = define_satrays_from_tensors(rays, ts)
lib = K.lib()
def run(static):
    with torch.no_grad():
        return sat_rendering.render_image(m, None, sat, None, None, epoch_idx=2, chunk=B, render_step_size=2.0 / n, eval=True, static=static)
for static in (False, True):
    for _ in range(2): run(static)
    torch.cuda.synchronize()
    lib.eonerf_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); res, ns = run(static); e1.record(); torch.cuda.synchronize()
    lib.eonerf_profile_enable(0)
    prof = (K.Profile * 5)(); lib.eonerf_profile_read(prof, 5)
    print(f"static={static}: chunk {e0.elapsed_time(e1):.2f} ms, samples {int(ns)}, fused fwd kernels {prof[3].ms:.2f} ms in {prof[3].launches} launches "
          f"({prof[3].flops / max(prof[3].ms, 1e-9) / 1e9:.0f} TFLOP/s by the host-side count)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(True); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
