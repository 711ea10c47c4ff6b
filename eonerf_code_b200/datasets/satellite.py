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
