"""Per-kernel GPU time of one eager training step (torch profiler / CUPTI), sorted: where the non-dominant time goes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.datasets.synthetic import make_rays
from eonerf_code_b200.radiance_fields import EONerfMLP
from eonerf_code_b200.training import TrainStep
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
step = TrainStep(m, n_samples=128, graph="--graph" in sys.argv)
batches = [tuple(t.to(dev) for t in make_rays(8192, 19, seed=42 + i)) for i in range(3)]
for i in range(4):
    step(*batches[i % 3], 2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        step(*batches[i % 3], 2)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 3e3:.3f} ms per step over {sum(e.count for e in rows) // 3} kernels/memops")
for e in rows[:int(os.environ.get("TOPN", "40"))]:
    print(f"{e.device_time_total / 3:9.1f} us/step {e.count // 3:4d}x  {e.key[:110]}")
