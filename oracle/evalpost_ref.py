"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's evaluation epilogue.

* utm_points: SatelliteDataset.get_utmalt_from_nerf_prediction, utm_sampling branch (/root/reference/datasets/satellite.py:
  502-531): rays/depth promoted to fp64, xyz = (o + d * depth) * scene_scale + scene_offset.
* plyflatten: the rasteriser behind get_dsm_from_nerf_prediction (satellite.py:548-587), `plyflatten(cloud, xoff, yoff,
  resolution, xsize, ysize, radius, sigma)` of the pip package `plyflatten` (an un-vendored dependency of the reference, not
  in this image).  Restated from its published C source (plyflatten.c: rescale_float_to_int = floor((x - min) / res) on x
  and on -y against -ymax; every cell (i+k1, j+k2) with k1^2 + k2^2 <= radius^2 accumulates the height with weight
  exp(-d^2 / (2 sigma^2)), 1 when sigma is inf; output = weighted mean, NaN for empty cells).  PARITY UNPINNED: no upstream
  source, tests or vectors are available here.
* dsm_from_prediction: satellite.py:556-577 (north shift of negative norths, negative depths dropped, raster bounds).
* nadir rays: compared against the reference's own function in tests/golden (oracle/make_golden.py gen_nadir)."""
import math

import numpy as np


def utm_points(rays, depth, scene_scale, scene_offset):
    rays, depth = np.asarray(rays, np.float64), np.asarray(depth, np.float64).reshape(-1, 1)
    xyz = (rays[:, 0:3] + rays[:, 3:6] * depth) * np.asarray(scene_scale, np.float64) + np.asarray(scene_offset, np.float64)
    return xyz[:, 0], xyz[:, 1], xyz[:, 2]


def plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius=1, sigma=float("inf")):
    acc = np.zeros((ysize, xsize), np.float64)
    wsum = np.zeros((ysize, xsize), np.float64)
    x, y, z = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    i = np.floor((x - xoff) / resolution).astype(np.int64)
    j = np.floor((-y + yoff) / resolution).astype(np.int64)
    for k1 in range(-radius, radius + 1):
        for k2 in range(-radius, radius + 1):
            if k1 * k1 + k2 * k2 > radius * radius:
                continue
            ii, jj = i + k1, j + k2
            ok = (ii >= 0) & (jj >= 0) & (ii < xsize) & (jj < ysize)
            if math.isinf(sigma):
                w = np.ones_like(x)
            else:
                dx = x - (xoff + resolution * (0.5 + ii))
                dy = y - (yoff - resolution * (0.5 + jj))
                w = np.exp(-(dx * dx + dy * dy) / (2 * sigma * sigma))
            np.add.at(acc, (jj[ok], ii[ok]), (w * z)[ok])
            np.add.at(wsum, (jj[ok], ii[ok]), w[ok])
    with np.errstate(invalid="ignore", divide="ignore"):
        out = np.where(wsum > 0, acc / wsum, np.nan)
    return out.astype(np.float32)


def dsm_from_prediction(rays, depth, scene_scale, scene_offset, resolution=0.5, radius=1, sigma=float("inf")):
    e, n, a = utm_points(rays, depth, scene_scale, scene_offset)
    cloud = np.vstack([e, n, a]).T
    cloud[cloud[:, 1] < 0, 1] += 10e6
    cloud = cloud[np.asarray(depth).reshape(-1) >= 0.0, :]
    xmin, xmax, ymin, ymax = cloud[:, 0].min(), cloud[:, 0].max(), cloud[:, 1].min(), cloud[:, 1].max()
    xoff = np.floor(xmin / resolution) * resolution
    xsize = int(1 + np.floor((xmax - xoff) / resolution))
    yoff = np.ceil(ymax / resolution) * resolution
    ysize = int(1 - np.floor((ymin - yoff) / resolution))
    return plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma), (xoff, yoff, xsize, ysize, resolution)
