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


def _scale_offset(scene_scale, scene_offset):
    import torch
    sc = [float(v) for v in torch.as_tensor(scene_scale, dtype=torch.float64).flatten().tolist()]
    of = [float(v) for v in torch.as_tensor(scene_offset, dtype=torch.float64).flatten().tolist()]
    if len(sc) != 3 or len(of) != 3:
        raise ValueError("scene_scale / scene_offset must hold 3 values")
    return sc, of


def get_utmalt_from_nerf_prediction(rays, depth, scene_scale, scene_offset, double=True, want_alt_f32=False):
    """satellite.py:502-531 (the `utm_sampling` branch): the evaluation epilogue that turns a rendered depth map into a
    point cloud, x = (o + d * depth) * scene_scale + scene_offset, in fp64 as the reference does to keep metre-level UTM
    coordinates exact.  rays [N,>=6] fp32, depth [N,1] or [N] fp32 on the GPU; scene_scale / scene_offset [3] (host values).
    -> (easts, norths, alts), each fp64 [N] (+ the fp32 altitude when want_alt_f32).  One sm_100a kernel
    (`eonerf_utm_points`, csrc/evalpost.cu); the ECEF / lon-lat branch needs the reference's un-vendored geodesy helpers and
    stays out of scope."""
    import torch
    from .. import _capi as K
    from ..ops import _f32, _need_cuda, on_tensor_device
    if not double:
        raise NotImplementedError("double=False is only used by the reference's lon/lat branch (satellite.py:535), out of scope")

    @on_tensor_device
    def run(rays, depth):
        _need_cuda(rays, depth)
        r = _f32(rays)
        if r.stride(1) != 1:
            r = r.contiguous()
        d = _f32(depth).reshape(-1)
        N = r.shape[0]
        if d.numel() != N:
            raise ValueError("depth must hold one value per ray")
        out = torch.empty(3, N, dtype=torch.float64, device=r.device)
        alt32 = torch.empty(N, dtype=torch.float32, device=r.device) if want_alt_f32 else None
        a = K.UtmPointsArgs()
        a.rays, a.rays_stride, a.depth, a.depth_stride, a.n_rays = r.data_ptr(), r.stride(0), d.data_ptr(), d.stride(0), N
        sc, of = _scale_offset(scene_scale, scene_offset)
        for c in range(3):
            a.scene_scale[c], a.scene_offset[c] = sc[c], of[c]
        a.easts, a.norths, a.alts = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
        a.alt_f32 = None if alt32 is None else alt32.data_ptr()
        K.call("utm_points", a, torch.cuda.current_stream().cuda_stream)
        return (out[0], out[1], out[2]) + ((alt32,) if want_alt_f32 else ())
    return run(rays, depth)


def dsm_grid_from_cloud(easts, norths, resolution):
    """Bounds of the DSM raster as satellite.py:571-577 derives them from the cloud -> (xoff, yoff, xsize, ysize)."""
    import math
    import torch
    lo_hi = torch.stack([easts.min(), easts.max(), norths.min(), norths.max()]).tolist()      # the one host read
    xmin, xmax, ymin, ymax = lo_hi
    xoff = math.floor(xmin / resolution) * resolution
    xsize = int(1 + math.floor((xmax - xoff) / resolution))
    yoff = math.ceil(ymax / resolution) * resolution
    ysize = int(1 - math.floor((ymin - yoff) / resolution))
    return xoff, yoff, xsize, ysize


def get_dsm_from_nerf_prediction(rays, depth, scene_scale, scene_offset, resolution=0.5, roi=None, radius=1, sigma=float("inf")):
    """satellite.py:548-587 up to (not including) the GeoTIFF write: point cloud from the rendered depth, norths < 0 shifted by
    10e6, negative depths dropped, raster bounds from the cloud (or `roi` = (xoff, yoff, xsize, ysize, resolution) as the
    reference reads them from roi_txt, :564-569), then plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius=1,
    sigma=inf) on the GPU (`eonerf_dsm_rasterize`).  -> (dsm [ysize,xsize] fp32 device tensor, NaN = empty cell;
    (xoff, yoff, xsize, ysize, resolution)).  plyflatten is un-vendored: semantics restated, parity unpinned."""
    import torch
    from .. import _capi as K
    from ..ops import _f32
    easts, norths, alts = get_utmalt_from_nerf_prediction(rays, depth, scene_scale, scene_offset)
    d = _f32(depth).reshape(-1)
    with torch.cuda.device(easts.device):
        if roi is not None:
            xoff, yoff, xsize, ysize, resolution = roi
        else:
            keep = d >= 0
            n_fix = torch.where(norths < 0, norths + 10e6, norths)
            xoff, yoff, xsize, ysize = dsm_grid_from_cloud(easts[keep], n_fix[keep], resolution)
        acc = torch.empty(ysize * xsize * 2, dtype=torch.float64, device=easts.device)
        dsm = torch.empty(ysize, xsize, dtype=torch.float32, device=easts.device)
        a = K.DsmArgs(easts.data_ptr(), norths.data_ptr(), alts.data_ptr(), d.data_ptr(), d.stride(0), easts.numel(), float(xoff), float(yoff),
                      float(resolution), int(xsize), int(ysize), int(radius), float(sigma), 10e6, acc.data_ptr(), dsm.data_ptr())
        K.call("dsm_rasterize", a, torch.cuda.current_stream().cuda_stream)
    return dsm, (xoff, yoff, xsize, ysize, resolution)


def get_dir_vec_from_el_az(elevation_deg, azimuth_deg):
    """satellite.py:57-63: elevation is 0 degrees at nadir, 90 at frontal view."""
    import numpy as np
    el = np.radians(90 - elevation_deg)
    az = np.radians(azimuth_deg)
    return -1.0 * np.array([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)])


def create_rays_from_nadir(scene_scale, h, w, sun_el_deg, sun_az_deg, img_downscale=1.0, device=None):
    """eval_eonerf.py:78-96 + generate_rays_from_virtual_pinhole (:135-249, the live `pinhole = False` branch): the virtual
    nadir view `eval_eonerf.py --dsm` renders — parallel rays along the (scaled, normalised) nadir direction from a plane
    2 units above the centre of the cube's bottom face, x in [-1, 1), y in (-1, 1], near 0, far 2.5, constant sun direction.
    Host-side ray generation (numpy fp64 -> fp32, as the reference), returned as the [h*w, 11] table on `device`."""
    import numpy as np
    import torch
    scale = np.asarray(torch.as_tensor(scene_scale, dtype=torch.float64).cpu().numpy(), dtype=np.float64)
    radius = 2
    h, w = int(h // img_downscale), int(w // img_downscale)
    near = max(0, radius - 2)
    far = near + 2.5
    d = get_dir_vec_from_el_az(0, 0)
    d = d / scale
    d = d / np.linalg.norm(d)
    pt_a = np.array([0, 0, -1]) - radius * d
    x = (np.arange(w) - w * 0.5) / (1 * w / radius) + pt_a[0]
    y = -(np.arange(h) - h * 0.5) / (1 * h / radius) + pt_a[1]
    X, Y = np.meshgrid(x, y)
    Z = ((-d[0] * (X - pt_a[0]) - d[1] * (Y - pt_a[1])) / d[2]) + pt_a[2]
    origins = np.vstack([X.ravel(), Y.ravel(), Z.ravel()]).T
    directions = np.tile(d, (h, w, 1))
    viewdirs = (directions / np.linalg.norm(directions, axis=-1, keepdims=True)).reshape(-1, 3)
    ones = np.ones_like(origins[..., :1])
    rays = torch.from_numpy(np.hstack([origins, viewdirs, near * ones, far * ones])).type(torch.FloatTensor)
    sun_d = get_dir_vec_from_el_az(sun_el_deg, sun_az_deg)
    sun = torch.from_numpy(np.tile(sun_d, (rays.shape[0], 1)))
    sun /= torch.as_tensor(scale)
    sun /= np.linalg.norm(sun, axis=1)[:, np.newaxis]
    rays = torch.hstack([rays, sun.type(torch.FloatTensor)])
    return rays if device is None else rays.to(device)
