"""Synthetic "JAX_068-shaped" satellite rays (no dataset, no network).

The real pipeline builds rays on the host from RPC camera models
(/root/reference/datasets/satellite.py:65-121 `get_rays`, :124-139 `normalize_rays`) and hands the
renderer a `[B, 11]` fp32 table `[o(3) d(3) near far sun(3)]` plus an int64 image index
(/root/reference/datasets/satellite.py:23-26).  RPC ray generation stays on the host and is out of
scope; this module produces tables with the same layout and statistics:

* scene half extents X=Y=140 m, Z=50 m  -> anisotropic normalisation of directions,
* per image: off-nadir 5..35 deg, azimuth 0..360 deg, sun elevation 35..70 deg, sun azimuth 100..180 deg,
* origins on the max-altitude plane z=+1 ("spread" variant) or back-projected from a ground point so
  that every stratified sample on t in [0,2] lies inside the cube ("inside" variant, what real scenes
  look like because `scene.loc` scaling is fitted that way, satellite.py:377-404).
"""
import math
import numpy as np
import torch

SCENE_SCALE = np.array([140.0, 140.0, 50.0])


def _dir_from_el_az(elevation_deg, azimuth_deg):
    # same convention as satellite.py:57-63: elevation 0 = nadir-looking, vector points downwards
    el = np.radians(90.0 - elevation_deg)
    az = np.radians(azimuth_deg)
    return -1.0 * np.stack([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)], -1)


def make_rays(n_rays, n_images=19, seed=42, variant="spread", eval_mode=False, device="cpu"):
    """Returns (rays[B,11] fp32, ts[B,1] int64, pixels[B,3] fp32)."""
    rng = np.random.default_rng(seed)
    theta = np.radians(rng.uniform(5.0, 35.0, n_images))
    phi = np.radians(rng.uniform(0.0, 360.0, n_images))
    sun_el = rng.uniform(35.0, 70.0, n_images)
    sun_az = rng.uniform(100.0, 180.0, n_images)

    ts = np.zeros(n_rays, np.int64) if eval_mode else rng.integers(0, n_images, n_rays)
    img = ts if not eval_mode else rng.integers(0, 1, n_rays)
    d_metric = np.stack([np.sin(theta[img]) * np.cos(phi[img]),
                         np.sin(theta[img]) * np.sin(phi[img]),
                         -np.cos(theta[img])], -1)
    d_metric = d_metric + rng.normal(0.0, 1e-3, d_metric.shape)
    d = d_metric / SCENE_SCALE
    d = d / np.linalg.norm(d, axis=1, keepdims=True)

    if variant == "spread":
        o = np.concatenate([rng.uniform(-0.95, 0.95, (n_rays, 2)), np.ones((n_rays, 1))], 1)
    elif variant == "inside":
        p = np.concatenate([rng.uniform(-0.7, 0.7, (n_rays, 2)), -np.ones((n_rays, 1))], 1)
        # the sampler spans t in [0, 2] whatever t_far says (sat_rendering.py:60-63): start the ray so
        # that the ground point sits at t = 1.999 and every interval midpoint stays strictly inside the cube
        o = p - d * 1.999
        o[:, 2] = np.minimum(o[:, 2], 0.9995)
    else:
        raise ValueError(variant)
    far = 2.0 / np.abs(d[:, 2:3])
    near = np.zeros_like(far)

    s = _dir_from_el_az(sun_el[img], sun_az[img]) / SCENE_SCALE
    s = s / np.linalg.norm(s, axis=1, keepdims=True)

    rays = torch.from_numpy(np.hstack([o, d, near, far, s]).astype(np.float32))
    pixels = torch.from_numpy(rng.uniform(0.0, 1.0, (n_rays, 3)).astype(np.float32))
    ts_t = torch.from_numpy(ts.astype(np.int64)).view(-1, 1)
    return rays.to(device), ts_t.to(device), pixels.to(device)


def make_pinhole_rays(n_rays, seed=42, width=800, height=800, camera_angle_x=0.6911112070083618,
                      radius=4.0311, device="cpu"):
    """Blender/"lego"-shaped pinhole rays for the vanilla-NeRF benchmark (BASELINE config 2):
    800x800, focal from camera_angle_x, camera on a sphere of radius ~4.03 looking at the origin
    (/root/reference/datasets/nerf_synthetic.py:69-70,165-233).  Returns (origins[B,3], viewdirs[B,3], pixels[B,3])."""
    rng = np.random.default_rng(seed)
    focal = 0.5 * width / math.tan(0.5 * camera_angle_x)
    n_cam = 100
    el = rng.uniform(0.0, 0.5 * math.pi * 0.9, n_cam)
    az = rng.uniform(0.0, 2 * math.pi, n_cam)
    cam = rng.integers(0, n_cam, n_rays)
    c = radius * np.stack([np.cos(el[cam]) * np.cos(az[cam]), np.cos(el[cam]) * np.sin(az[cam]), np.sin(el[cam])], -1)
    fwd = -c / np.linalg.norm(c, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right, axis=1, keepdims=True)
    true_up = np.cross(right, fwd)
    x = rng.integers(0, width, n_rays) + 0.5
    y = rng.integers(0, height, n_rays) + 0.5
    dx = (x - 0.5 * width) / focal
    dy = -(y - 0.5 * height) / focal
    d = fwd + dx[:, None] * right + dy[:, None] * true_up
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pixels = rng.uniform(0.0, 1.0, (n_rays, 3))
    f32 = lambda a: torch.from_numpy(a.astype(np.float32)).to(device)
    return f32(c), f32(d), f32(pixels)
