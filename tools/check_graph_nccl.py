"""torchrun --nproc-per-node 2 tools/check_graph_nccl.py : the data-parallel CUDA-graph step (all-reduce captured in the graph) against
the eager data-parallel step on the same batches and uniforms: parameters after three steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eonerf_code_b200.datasets.synthetic import make_rays  # noqa: E402
from eonerf_code_b200.radiance_fields import EONerfMLP  # noqa: E402
from eonerf_code_b200.training import TrainStep  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, n = 2048, 128
res = {}
for mode in ("graph", "eager"):
    torch.manual_seed(42)
    m = EONerfMLP(19, radiometric_normalization=True, precision="bf16_fused").to(dev)
    step = TrainStep(m, n_samples=n, world=world, graph=mode == "graph")
    for i in range(3):
        rays, ts, px = (t.to(dev) for t in make_rays(B, 19, seed=100 + 10 * i + rank))
        g = torch.Generator(device=dev).manual_seed(7 + i + 100 * rank)
        uni = {k: torch.rand(B, n, device=dev, generator=g) for k in ("u_cam", "u_sun", "u_cam2")}
        loss, _ = step(rays, ts, px, 2, uniforms=uni)
    torch.cuda.synchronize()
    res[mode] = torch.cat([p.detach().flatten() for p in m.parameters()]).clone()
    res[mode + "_loss"] = float(loss)
d = (res["graph"] - res["eager"]).abs().max().item()
ref = res["eager"].abs().max().item()
# replicas stay identical across ranks
mine = res["graph"].clone()
other = mine.clone()
dist.broadcast(other, src=0)
print(f"rank {rank}: loss graph {res['graph_loss']:.6f} eager {res['eager_loss']:.6f}; max |param graph - eager| = {d:.3e} (max |param| {ref:.3f}); "
      f"max |param - rank 0's| = {(mine - other).abs().max().item():.3e}", flush=True)
assert d <= 2e-3 and (mine - other).abs().max().item() == 0.0
dist.destroy_process_group()
