"""Timeline of one CUDA-graph training step (torch profiler / CUPTI): busy time (union of kernel intervals), idle gaps between
kernels and the largest of them, per step.   python tools/step_gaps.py"""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.datasets.synthetic import make_rays  # noqa: E402
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402
from eonerf_code_b200.training import TrainStep  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)
m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
step = TrainStep(m, n_samples=128, graph=True)
batches = [tuple(t.to(dev) for t in make_rays(8192, 19, seed=42 + i)) for i in range(3)]
for i in range(6):
    step(*batches[i % 3], 2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(5):
        step(*batches[i % 3], 2)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "step_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
ev = ev[len(ev) // 5:]          # the first profiled step carries the profiler's start-up gaps: drop it
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
# union of intervals
busy, cur_s, cur_e, gaps = 0.0, ev[0]["ts"], ev[0]["ts"] + ev[0]["dur"], []
prev_name = ev[0]["name"]
for e in ev[1:]:
    if e["ts"] > cur_e:
        gaps.append((e["ts"] - cur_e, prev_name[:60], e["name"][:60]))
        busy += cur_e - cur_s
        cur_s, cur_e = e["ts"], e["ts"] + e["dur"]
    else:
        cur_e = max(cur_e, e["ts"] + e["dur"])
    if e["ts"] + e["dur"] >= cur_e:
        prev_name = e["name"]
busy += cur_e - cur_s
n_steps = 4
print(f"{len(ev)} GPU activities over {n_steps} steps; span {(t1 - t0) / n_steps / 1e3:.3f} ms/step, busy {busy / n_steps / 1e3:.3f} ms/step, "
      f"idle {(t1 - t0 - busy) / n_steps / 1e3:.3f} ms/step in {len(gaps) // n_steps} gaps/step; sum of durations {sum(e['dur'] for e in ev) / n_steps / 1e3:.3f} ms/step")
agg = {}
for g, a, b in gaps:
    k = (a, b)
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += g
    agg[k][1] += 1
print("largest idle gaps (us per step, count per step): after -> before")
for (a, b), (g, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:14]:
    print(f"  {g / n_steps:8.1f} us {c / n_steps:5.1f}x  {a}  ->  {b}")
if os.environ.get("SEQ"):
    # the activities around every gap > 50 us
    run_end, shown = ev[0]["ts"] + ev[0]["dur"], 0
    for i in range(1, len(ev)):
        if ev[i]["ts"] - run_end > 50 and shown < 5:
            shown += 1
            print(f"---- idle {ev[i]['ts'] - run_end:.1f} us before activity {i}")
            for e in ev[max(0, i - 4):i + 4]:
                print(f"   t={(e['ts'] - t0) / 1e3:9.3f} ms dur={e['dur']:8.1f} us  stream={e.get('args', {}).get('stream')}  {e['name'][:90]}")
        run_end = max(run_end, ev[i]["ts"] + ev[i]["dur"])
