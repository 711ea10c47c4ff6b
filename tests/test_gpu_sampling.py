"""GPU: stratified sampling + cube mask + compaction through the C ABI — bit-exact (ray_indices, t_starts, t_ends,
per-ray counts) against the golden vectors from the reference and against the oracle at larger sizes."""
import pytest
import torch

from helpers import t
from oracle import eonerf_oracle as O

pytestmark = pytest.mark.gpu


def _run(cuda, rays, u, n, z_steps, near=True):
    from eonerf_code_b200 import ops
    r = rays.to(cuda)
    ri, ts, te, ppr, offs, stats = ops.sample_compact(r[:, 0:3], r[:, 3:6], r[:, 6:7] if near else None, u.to(cuda),
                                                      None if z_steps is None else z_steps.to(cuda))
    P, n_empty = stats.tolist()
    return ri[:P].cpu(), ts[:P].cpu(), te[:P].cpu(), ppr.cpu(), offs.cpu(), P, n_empty


def test_golden_bit_exact(cuda, golden):
    g = golden["sampling"]
    for tag in ("n64", "n96", "n128_inside"):
        rays, u, n = t(g[f"{tag}_rays"]), t(g[f"{tag}_u"]), int(g[f"{tag}_n"])
        ri, ts, te, ppr, offs, P, n_empty = _run(cuda, rays, u, n, t(g[f"{tag}_z_steps"]))
        assert P == g[f"{tag}_ray_indices"].shape[0]
        assert torch.equal(ri, t(g[f"{tag}_ray_indices"]))
        assert torch.equal(ts, t(g[f"{tag}_t_starts"])), (ts - t(g[f"{tag}_t_starts"])).abs().max()
        assert torch.equal(te, t(g[f"{tag}_t_ends"]))
        assert torch.equal(ppr, t(g[f"{tag}_pts_per_ray"]))
        assert n_empty == int((g[f"{tag}_pts_per_ray"] == 0).sum())
        assert torch.equal(offs[1:] - offs[:-1], t(g[f"{tag}_pts_per_ray"]).long())


def test_device_linspace_equals_host_linspace(cuda):
    """The product takes z_steps from torch.linspace on the device, as the reference would (sat_rendering.py:67)."""
    for n in (32, 47, 64, 95, 100, 128, 191, 256):
        assert torch.equal(torch.linspace(0, 1, n, device=cuda).cpu(), torch.linspace(0, 1, n)), n


@pytest.mark.parametrize("B,n,variant", [(4096, 128, "spread"), (8192, 64, "spread"), (2048, 128, "inside"), (1000, 95, "spread")])
def test_oracle_bit_exact_large(cuda, B, n, variant):
    from eonerf_code_b200.datasets.synthetic import make_rays
    rays, _, _ = make_rays(B, 19, seed=B + n, variant=variant)
    rays[3, 0:3] = torch.tensor([5.0, 5.0, 1.0])                      # an empty ray
    rays[::7, 6] = 0.125                                              # some non-zero t_near
    u = torch.rand(B, n, generator=torch.Generator().manual_seed(n))
    ri, ts, te, ppr, offs, P, n_empty = _run(cuda, rays, u, n, None)
    ori, ots, ote, _ = O.satnerf_sampling(rays[:, 0:3], rays[:, 3:6], n, u, near=rays[:, 6:7])
    assert P == ori.numel() and torch.equal(ri, ori) and torch.equal(ts, ots) and torch.equal(te, ote)
    assert torch.equal(ppr, O.pts_per_ray(ori, B)) and n_empty == int((O.pts_per_ray(ori, B) == 0).sum()) and n_empty >= 1
    if variant == "inside":
        assert P == (B - n_empty) * (n - 1) or P <= B * (n - 1)
    # sortedness / order-preservation properties
    assert torch.all(ri[1:] >= ri[:-1]) and torch.all(te >= ts)
    same = ri[1:] == ri[:-1]
    assert torch.all(ts[1:][same] >= ts[:-1][same])


def test_edge_cases(cuda):
    from eonerf_code_b200 import ops, sat_rendering
    # zero rays
    z = torch.zeros(0, 3, device=cuda)
    ri, ts, te, ppr, offs, stats = ops.sample_compact(z, z, None, torch.zeros(0, 16, device=cuda))
    assert stats.tolist() == [0, 0] and offs.tolist() == [0]
    # every ray outside the cube -> P = 0
    o = torch.full((5, 3), 3.0, device=cuda)
    d = torch.tensor([[0.0, 0.0, -1.0]], device=cuda).repeat(5, 1)
    ri, ts, te = sat_rendering.satnerf_sampling(o, d, {"render_step_size": 2 / 16})
    assert ri.numel() == 0 and ts.numel() == 0
    # near=None equals near=0 (sat_rendering.py:60-61)
    o = torch.tensor([[0.1, -0.2, 1.0]], device=cuda).repeat(4, 1)
    u = torch.rand(4, 16, device=cuda)
    a = ops.sample_compact(o, d[:4], None, u)
    b = ops.sample_compact(o, d[:4], torch.zeros(4, 1, device=cuda), u)
    assert a[5].tolist() == b[5].tolist() and torch.equal(a[1][:a[5][0]], b[1][:b[5][0]])
    # count_number_of_pts_per_nerfacc_ray (sat_rendering.py:10-16)
    from eonerf_code_b200.datasets.satellite import SatRays
    rays = SatRays(o, d[:4], d[:4], None, None, None)
    ri = torch.tensor([0, 0, 2, 2, 2], device=cuda)
    assert sat_rendering.count_number_of_pts_per_nerfacc_ray(rays, ri).tolist() == [2.0, 0.0, 3.0, 0.0]
