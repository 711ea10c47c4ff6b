"""Evaluation epilogue: nadir virtual-view rays, UTM/altitude point cloud, DSM rasterisation."""
import numpy as np
import pytest
import torch

from helpers import t
from oracle import evalpost_ref as E


def test_nadir_rays_match_the_reference(golden):
    """create_rays_from_nadir (eval_eonerf.py:78-249) vs the reference function's own output (tests/golden/nadir.npz)."""
    from eonerf_code_b200.datasets.satellite import create_rays_from_nadir
    g = golden["nadir"]
    for tag in ("a", "b"):
        rays = create_rays_from_nadir(g["scene_scale"], int(g[f"{tag}_h"]), int(g[f"{tag}_w"]), float(g[f"{tag}_sun_el"]),
                                      float(g[f"{tag}_sun_az"]), img_downscale=float(g[f"{tag}_downscale"]))
        ref = t(g[f"{tag}_rays"])
        assert rays.shape == ref.shape and rays.dtype == torch.float32
        assert torch.equal(rays, ref), tag


def test_oracle_utm_points_match_the_reference(golden):
    g = golden["nadir"]
    e, n, a = E.utm_points(g["b_rays"], g["utm_depth"], g["scene_scale"], g["scene_offset"])
    assert np.array_equal(e, g["utm_easts"]) and np.array_equal(n, g["utm_norths"]) and np.array_equal(a, g["utm_alts"])


def test_oracle_plyflatten_properties():
    """Known answers of the restated plyflatten: one point splats its height over the plus-shaped 5-cell neighbourhood
    (radius 1), means combine, empty cells are NaN, points outside the grid are ignored."""
    cloud = np.array([[10.25, 99.75, 5.0], [10.25, 99.75, 7.0], [12.75, 98.25, 1.0], [-50.0, 0.0, 9.0]])
    dsm = E.plyflatten(cloud, 9.0, 101.0, 0.5, 10, 8, radius=1)
    i, j = int((10.25 - 9.0) / 0.5), int((101.0 - 99.75) / 0.5)
    assert dsm.shape == (8, 10)
    for di, dj in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)):
        assert dsm[j + dj, i + di] == 6.0
    assert np.isnan(dsm[j + 1, i + 1]) and int(np.isfinite(dsm).sum()) == 10
    g = E.plyflatten(cloud[:2], 9.0, 101.0, 0.5, 10, 8, radius=0, sigma=0.1)
    assert g[j, i] == 6.0 and int(np.isfinite(g).sum()) == 1


@pytest.mark.gpu
def test_utm_points_and_altitude_kernel(cuda, golden):
    """eonerf_utm_points vs the reference's own fp64 outputs: bit-exact (separately rounded fp64 multiply / add)."""
    from eonerf_code_b200.datasets.satellite import get_utmalt_from_nerf_prediction
    g = golden["nadir"]
    rays, depth = t(g["b_rays"], cuda), t(g["utm_depth"], cuda)
    e, n, a, alt32 = get_utmalt_from_nerf_prediction(rays, depth, g["scene_scale"], g["scene_offset"], want_alt_f32=True)
    assert e.dtype == torch.float64 and alt32.dtype == torch.float32
    assert torch.equal(e.cpu(), t(g["utm_easts"])) and torch.equal(n.cpu(), t(g["utm_norths"])) and torch.equal(a.cpu(), t(g["utm_alts"]))
    assert torch.equal(alt32.cpu(), t(g["utm_alts"]).float())
    with pytest.raises(RuntimeError):
        get_utmalt_from_nerf_prediction(rays.cpu(), depth.cpu(), g["scene_scale"], g["scene_offset"])       # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("n,south", [(20000, False), (1 << 20, False), (5000, True)])
def test_dsm_rasterize_vs_oracle(cuda, n, south):
    """get_dsm_from_nerf_prediction (GPU: eonerf_utm_points + eonerf_dsm_rasterize) vs the numpy restatement, incl. the
    negative-depth filter and the southern-hemisphere north shift; 1 Mi rays = BASELINE configs[3]'s image size."""
    from eonerf_code_b200.datasets.satellite import create_rays_from_nadir, get_dsm_from_nerf_prediction
    scale = np.array([143.5, 139.25, 51.0])
    offset = np.array([435500.5, -3354950.25 if south else 3354950.25, 12.5])
    side = int(round(n ** 0.5))
    rays = create_rays_from_nadir(scale, side, side, 40.0, 140.0)
    g = torch.Generator().manual_seed(n)
    depth = 1.0 + 0.8 * torch.rand(rays.shape[0], 1, generator=g)
    depth[::97] = -0.5                                                   # dropped (satellite.py:561)
    dsm, grid = get_dsm_from_nerf_prediction(rays.to(cuda), depth.to(cuda), scale, offset, resolution=0.5)
    ref, grid_o = E.dsm_from_prediction(rays.numpy(), depth.numpy(), scale, offset, resolution=0.5)
    assert tuple(grid) == tuple(grid_o) and dsm.shape == ref.shape
    got = dsm.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ok.sum() > 0.5 * side * side / 4 and np.allclose(got[ok], ref[ok], rtol=1e-6, atol=1e-5)
