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


def gather_rows(t, world, dst=0, n_total=None, out=None):
    """Evaluation: every rank renders the contiguous block of rows `shard_bounds(n_total, rank, world)` gives it; rank `dst`
    receives the concatenation [n_total, ...] (everybody else gets None).  The row counts follow from shard_bounds, so there
    is no size exchange, no host synchronisation and no padding: rank `dst` posts one receive per peer straight into the
    slice of the (pre-allocated, reusable via `out`) result, the peers post one send each, all in one batched P2P group.
    n_total: total number of rows (default: world * t.shape[0], equal shards)."""
    if world == 1:
        return t
    rank = dist.get_rank()
    if n_total is None:
        n_total = world * t.shape[0]
    b0, b1 = shard_bounds(n_total, rank, world)
    if b1 - b0 != t.shape[0]:
        raise ValueError(f"rank {rank} holds {t.shape[0]} rows, shard_bounds gives {b1 - b0}")
    t = t.contiguous()
    if rank != dst:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, dst)]):
            w.wait()
        return None
    if out is None:
        out = torch.empty((n_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    ops = []
    for r in range(world):
        r0, r1 = shard_bounds(n_total, r, world)
        if r == dst:
            out[r0:r1].copy_(t)
        elif r1 > r0:
            ops.append(dist.P2POp(dist.irecv, out[r0:r1], r))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return out


class ChunkQueue:
    """Dynamic assignment of work chunks to ranks for the full-image evaluation render (strong scaling of a FIXED image): with
    contiguous row blocks the slowest GPU of the box sets the time (measured at 8 B200: per-rank render times 24.6-30.8 ms for
    equal shares, mean 27.5 ms); here every rank pulls the next chunk index from an atomic counter in the process group's
    key-value store, keeping at most `in_flight` chunks queued on its GPU, so faster GPUs render more chunks.  All ranks must
    construct their queues in the same order (the n-th queue of every rank shares one counter)."""
    _serial = 0

    def __init__(self, n_chunks, world, in_flight=2):
        self.n, self.world, self.in_flight = n_chunks, world, in_flight
        self.events = []
        self.next_static = 0
        if world > 1:
            self.store = dist.distributed_c10d._get_default_store()
            self.key = f"eonerf_chunk_queue_{ChunkQueue._serial}"
        ChunkQueue._serial += 1

    def __iter__(self):
        while True:
            if len(self.events) >= self.in_flight:          # back-pressure: the host must not run ahead of its GPU and drain the queue
                self.events.pop(0).synchronize()
            if self.world > 1:
                i = self.store.add(self.key, 1) - 1
            else:
                i, self.next_static = self.next_static, self.next_static + 1
            if i >= self.n:
                return
            yield i
            if torch.cuda.is_available():
                ev = torch.cuda.Event()
                ev.record()
                self.events.append(ev)


def reduce_disjoint(t, world, dst=0):
    """Every rank holds the full-size result with zeros where it rendered nothing: ONE sum-reduce to rank `dst` assembles the
    image (dynamic chunk assignment needs no metadata exchange this way).  -> the assembled tensor on `dst`, None elsewhere."""
    if world == 1:
        return t
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t if dist.get_rank() == dst else None


def guided_chunks(n_rows, world, min_rows=8):
    """Row ranges for ChunkQueue, large first: half of the rows in chunks of n_rows / (4 world), a quarter in chunks half that
    size, the rest in chunks a quarter of that size (never below min_rows).  Big chunks keep the kernels efficient, the small
    ones at the end bound the imbalance to one small chunk.  -> [(row_begin, row_end), ...]"""
    a = max(min_rows, n_rows // (4 * max(1, world)))
    out, r = [], 0
    for frac_end, size in ((0.5, a), (0.75, max(min_rows, a // 2)), (1.0, max(min_rows, a // 4))):
        end = n_rows if frac_end == 1.0 else int(n_rows * frac_end)
        while r < end:
            e = min(n_rows, r + size)
            out.append((r, e))
            r = e
    return out
