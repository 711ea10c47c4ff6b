"""Data parallelism over rays: one process per GPU, rays of a batch / rows of an image sharded across
ranks with no data-path collective, one NCCL all-reduce of the flat fp32 parameter gradient (2.7 MB) per training step,
one gather of the rendered rows for evaluation.  The reference is single-GPU (no torch.distributed anywhere)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """(rank, world, local_rank).  Initialises the default process group when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(n, rank, world):
    """Contiguous [begin, end) of `n` items owned by `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def flat_offsets(params, align=4):
    """Element offsets of the parameters inside a flat buffer; every tensor starts on a 16-byte boundary (align=4 floats),
    as separately allocated tensors do (the kernels use 128-bit accesses on rows of some of them).  -> (offsets, total)."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


class FlatGrads:
    """Makes every parameter's .grad a view of ONE flat fp32 buffer, so the gradient all-reduce is a single collective
    and optimizer.zero_grad(set_to_none=False) is a single memset."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, total = flat_offsets(self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self):
        self.flat.zero_()
        for p, off in zip(self.params, self.offsets):   # autograd may have replaced a .grad: re-point it at the flat buffer
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def all_reduce_mean(self, world, group=None):
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)


def gather_rows(t, world, dst=0):
    """Evaluation: every rank renders a contiguous block of rows; rank `dst` receives the concatenation."""
    if world == 1:
        return t
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    if dist.get_rank() != dst:
        return None
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], 0)
