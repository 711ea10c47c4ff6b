"""Boundary types of the ray table (/root/reference/datasets/satellite.py:21-30).

RPC ray generation stays on the host and is out of scope (BASELINE.json north_star); the renderer only
needs the `[B,11]` fp32 table `[o(3) d(3) near far sun(3)]` + int64 image index viewed as six fields."""
from collections import namedtuple

SatRays = namedtuple("SatRays", ("origins", "viewdirs", "sundirs", "img_idx", "t_near", "t_far"))


def define_satrays_from_tensors(rays, ts):
    """satellite.py:23-26 — column views, no copies."""
    return SatRays(origins=rays[:, 0:3], viewdirs=rays[:, 3:6], sundirs=rays[:, 8:11], img_idx=ts,
                   t_near=rays[:, 6:7], t_far=rays[:, 7:8])


def namedtuple_map(fn, tup):
    """datasets/utils.py:9-11"""
    return type(tup)(*(None if x is None else fn(x) for x in tup))


def get_utmalt_from_nerf_prediction(rays, depth, scene_scale, scene_offset, double=True):
    """satellite.py:502-531 (the `utm_sampling` branch): the evaluation epilogue that turns a rendered depth map into a
    point cloud, x = (o + d * depth) * scene_scale + scene_offset, in fp64 as the reference does to keep metre-level UTM
    coordinates exact.  rays [N,11], depth [N,1] or [N]; scene_scale / scene_offset [3].  -> (easts, norths, alts), each [N].
    O(N) element-wise work on the per-ray outputs, device-agnostic torch (the ECEF / lon-lat branch needs the reference's
    un-vendored geodesy helpers and stays out of scope)."""
    import torch
    if double:
        rays, depth = rays.double(), depth.double()
    scale = torch.as_tensor(scene_scale, dtype=rays.dtype, device=rays.device)
    offset = torch.as_tensor(scene_offset, dtype=rays.dtype, device=rays.device)
    xyz = (rays[:, 0:3] + rays[:, 3:6] * depth.reshape(-1, 1)) * scale + offset
    return xyz[:, 0], xyz[:, 1], xyz[:, 2]
