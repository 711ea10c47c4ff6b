"""CPU, gloo, world size 2: the data-parallel host logic of SURVEY.md §8e (rays sharded, one all-reduce of the flat
gradient, replicated optimiser step, row gather for evaluation).  The per-ray maths here is the ORACLE (test
infrastructure) because the product kernels need a GPU; what is under test is eonerf_code_b200/parallel.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from eonerf_code_b200 import parallel
    from eonerf_code_b200.datasets.synthetic import make_rays
    from oracle import eonerf_oracle as O
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    n_img, B, n = 3, 32, 16
    p = O.init_params(n_img, seed=5, bias_scale=0.05)
    params = [v.clone().requires_grad_(True) for v in p.values()]
    named = dict(zip(p.keys(), params))
    flat = parallel.FlatGrads(params)
    rays, ts, pixels = make_rays(B, n_img, seed=9)
    g = torch.Generator().manual_seed(11)
    u_cam, u_sun = torch.rand(B, n, generator=g), torch.rand(B, n, generator=g)
    b, e = parallel.shard_bounds(B, rank, world)
    sl = slice(b, e)
    out_l, _ = O.render_chunk(named, O.satrays_from_table(rays[sl], ts[sl]), n, 2, u_cam[sl], u_sun[sl], None)
    loss = O.loss_from_out(out_l, pixels[sl], 2)
    flat.zero()
    loss.backward()
    flat.all_reduce_mean(world)
    rows = parallel.gather_rows(out_l.detach()[:, :4].contiguous(), world)
    # uneven shards (7 rows over 2 ranks = 4 + 3) into a caller-provided buffer: no size exchange, only rank 0 receives
    u0, u1 = parallel.shard_bounds(7, rank, world)
    mine = torch.arange(u0, u1, dtype=torch.float32)[:, None].repeat(1, 3) + 0.5
    buf = torch.full((7, 3), -1.0) if rank == 0 else None
    uneven = parallel.gather_rows(mine, world, n_total=7, out=buf)
    assert (uneven is None) == (rank != 0)
    if rank == 0:
        assert uneven.data_ptr() == buf.data_ptr() and torch.equal(uneven, torch.arange(7.0)[:, None].repeat(1, 3) + 0.5)
    # dynamic chunk queue: every chunk index is handed out exactly once across the ranks; one sum-reduce assembles the result
    mine = list(parallel.ChunkQueue(11, world))
    img = torch.zeros(11, 2)
    for c in mine:
        img[c] = c + 1.0
    img = parallel.reduce_disjoint(img, world)
    got = [None] * world
    dist.all_gather_object(got, mine)
    assert sorted(sum(got, [])) == list(range(11))
    if rank == 0:
        assert torch.equal(img, (torch.arange(11.0) + 1)[:, None].repeat(1, 2))
    if rank == 0:       # per-parameter views (the flat buffer pads every tensor to a 16-byte boundary)
        torch.save({"grad": torch.cat([q.grad.reshape(-1) for q in params]), "rows": rows, "bounds": (b, e),
                    "flat_numel": flat.flat.numel(), "offsets": flat.offsets}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gradients_match_single_process(tmp_path):
    from eonerf_code_b200 import parallel
    from eonerf_code_b200.datasets.synthetic import make_rays
    from oracle import eonerf_oracle as O
    world, out = 2, str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)

    n_img, B, n = 3, 32, 16
    p = O.init_params(n_img, seed=5, bias_scale=0.05)
    params = [v.clone().requires_grad_(True) for v in p.values()]
    named = dict(zip(p.keys(), params))
    rays, ts, pixels = make_rays(B, n_img, seed=9)
    g = torch.Generator().manual_seed(11)
    u_cam, u_sun = torch.rand(B, n, generator=g), torch.rand(B, n, generator=g)
    out_full, _ = O.render_chunk(named, O.satrays_from_table(rays, ts), n, 2, u_cam, u_sun, None)
    loss = O.loss_from_out(out_full, pixels, 2)         # batch mean over equal shards == mean of the shard means
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    ref = torch.cat([(torch.zeros_like(q) if gq is None else gq).reshape(-1) for q, gq in zip(params, grads)])
    scale = float(ref.abs().max())
    assert float((got["grad"] - ref).abs().max()) <= 2e-5 * scale
    assert all(o % 4 == 0 for o in got["offsets"]) and got["flat_numel"] >= ref.numel()
    assert torch.allclose(got["rows"], out_full.detach()[:, :4], rtol=1e-6, atol=1e-7)


def test_shard_bounds_cover_everything():
    from eonerf_code_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1025):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_guided_chunks_cover_the_rows_once():
    from eonerf_code_b200.parallel import guided_chunks
    for n, w in ((1024, 1), (1024, 8), (1000, 3), (17, 4)):
        c = guided_chunks(n, w)
        assert c[0][0] == 0 and c[-1][1] == n and all(a[1] == b[0] for a, b in zip(c, c[1:]))
        assert c[0][1] - c[0][0] >= c[-1][1] - c[-1][0]
