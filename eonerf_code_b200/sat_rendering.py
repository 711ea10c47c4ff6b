"""Drop-in for /root/reference/sat_rendering.py (the live functions: :10-22, :46-118, :176-335).

Same names, argument meaning and return values; the arithmetic runs in sm_100a kernels (csrc/sampling.cu,
csrc/render.cu, csrc/field.cu).  Random numbers are drawn with torch on the device in the reference's
order (camera, [camera re-draw], sun — SURVEY.md Appendix C), or passed in explicitly for parity tests.
"""
from typing import Optional

import torch

from . import ops
from .datasets.satellite import SatRays, namedtuple_map  # noqa: F401

OUT_SLICES = (("rgb", 0, 3), ("depth", 3, 4), ("albedo_rgb", 4, 7), ("ambient_rgb", 7, 10), ("geo_shadows", 10, 11),
              ("transient_s", 11, 12), ("beta", 12, 13), ("entropy", 13, 14), ("pts_per_ray", 14, 15),
              ("sc_pts_per_ray", 15, 16), ("opacity_after_surface", 16, 18), ("shadowless_rgb", 18, 21))


def n_samples_from_step(render_step_size):
    return int(2 / render_step_size)                       # sat_rendering.py:64


def count_number_of_pts_per_nerfacc_ray(rays, ray_indices):
    """sat_rendering.py:10-16: fp32 histogram of ray_indices."""
    n_rays = rays.origins.shape[0]
    offs = ops.pack_info(ray_indices, n_rays)
    return (offs[1:] - offs[:-1]).to(torch.float32)


def _sample(origins, viewdirs, n_samples, near, u, z_steps):
    if u is None:
        u = torch.rand(origins.shape[0], n_samples, dtype=torch.float32, device=origins.device)   # :52
    ri, ts, te, ppr, offs, stats = ops.sample_compact(origins, viewdirs, near, u, z_steps)
    P, n_empty = stats.tolist()                            # one host sync (the reference has >= 3 here)
    return ri[:P], ts[:P], te[:P], ppr, offs, n_empty


def satnerf_sampling(origins, viewdirs, sampling_args, near=None, far=None, perturb=True, u=None, z_steps=None):
    """sat_rendering.py:56-84 -> (ray_indices i64[P], t_starts[P], t_ends[P]).  `u` [B,n] overrides the uniforms."""
    if far is not None:
        raise NotImplementedError("far != near + 2 is never used by the reference (sat_rendering.py:60-63)")
    n = n_samples_from_step(sampling_args["render_step_size"])
    if not perturb:
        raise NotImplementedError("perturb=False is not reachable from the reference's call sites (every caller uses the default)")
    ri, ts, te, _, _, _ = _sample(origins, viewdirs, n, near, u, z_steps)
    return ri, ts, te


def compute_geometric_shadows(chunk_rays, depth, radiance_field, occupancy_grid, sampling_args, u=None, z_steps=None,
                              info=None):
    """sat_rendering.py:87-118 -> (geo_shadow[B,1], sc_pts_per_ray[B]); differentiable in depth and the parameters."""
    e = radiance_field._engine()
    n = n_samples_from_step(sampling_args["render_step_size"])
    info = {} if info is None else info
    geo = ops._SunPassFn.apply(torch.is_grad_enabled(), e, chunk_rays.origins, chunk_rays.viewdirs, chunk_rays.sundirs, depth, n, u, z_steps, info,
                               False, *e.tensors())
    return geo, info["sc_pts_per_ray"]


def render_packed(
    radiance_field: torch.nn.Module,
    occupancy_grid,
    rays: SatRays,
    scene_aabb: torch.Tensor,
    args,
    epoch_idx: Optional[int] = None,
    chunk: int = 5120,
    near_plane: Optional[float] = None,
    far_plane: Optional[float] = None,
    render_step_size: float = 1e-3,
    render_bkgd: Optional[torch.Tensor] = None,
    cone_angle: float = 0.0,
    alpha_thre: float = 0.0,
    early_stop_eps: float = 0.0,
    timestamps: Optional[torch.Tensor] = None,
    only_depth: bool = False,
    eval: bool = False,
    uniforms=None,
    z_steps=None,
    static=False,
    counters=None,
):
    """The body of render_image: -> (packed [B,21] per-ray outputs in the column order of OUT_SLICES (or depth [B,1] when
    only_depth), n_rendering_samples, rays_shape).  training.TrainStep evaluates its fused loss on the packed tensor."""
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        num_rays = rays_shape[0] * rays_shape[1]
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays = rays_shape[0]
    sampling_args = {"render_step_size": render_step_size}
    n = n_samples_from_step(render_step_size)
    e = radiance_field._engine()
    outs, n_rendering_samples = [], 0
    for ci, i in enumerate(range(0, num_rays, chunk)):
        chunk_rays = namedtuple_map(lambda r: r[i:i + chunk], rays)
        us = uniforms[ci] if uniforms is not None else {}
        n_dev = None
        if static:
            B_c, dev = chunk_rays.origins.shape[0], chunk_rays.origins.device
            draw = lambda key: us[key] if us.get(key) is not None else torch.rand(B_c, n, dtype=torch.float32, device=dev)
            first = ops.sample_compact(chunk_rays.origins, chunk_rays.viewdirs, chunk_rays.t_near, draw("u_cam"), z_steps)
            ri, ts, te, ppr, offs, stats = ops.sample_compact(chunk_rays.origins, chunk_rays.viewdirs, None, draw("u_cam2"),
                                                              z_steps, redraw_of=first)     # :260-262, decided on the device
            n_dev = stats[0:1]
            n_rendering_samples = n_rendering_samples + stats[0]
        else:
            ri, ts, te, ppr, offs, n_empty = _sample(chunk_rays.origins, chunk_rays.viewdirs, n, chunk_rays.t_near,
                                                     us.get("u_cam"), z_steps)
            if n_empty:                                     # :260-262 re-draw with near=None
                ri, ts, te, _, offs, _ = _sample(chunk_rays.origins, chunk_rays.viewdirs, n, None, us.get("u_cam2"), z_steps)
            n_rendering_samples += ts.numel()
        comp = ops._CameraPassFn.apply(torch.is_grad_enabled(), e, only_depth, chunk_rays.origins, chunk_rays.viewdirs, chunk_rays.sundirs,
                                       chunk_rays.img_idx, ri, ts, te, offs, n_dev, *e.tensors())
        if only_depth:                                      # :227-249
            outs.append(comp[:, 3:4])
            continue
        geo = sc_ppr = None
        if epoch_idx >= 2:                                  # :269-276
            info = {}
            geo = ops._SunPassFn.apply(torch.is_grad_enabled(), e, chunk_rays.origins, chunk_rays.viewdirs, chunk_rays.sundirs, comp[:, 3:4], n,
                                       us.get("u_sun"), z_steps, info, static, *e.tensors())
            sc_ppr = info["sc_pts_per_ray"]
            if counters is not None:                        # kept sun samples Q (int, or a 0-d device tensor when static)
                counters["n_sun_samples"] = counters.get("n_sun_samples", 0) + info["n_sun_samples"]
        rad = radiance_field.radiometricT_enc.weight if radiance_field.radiometric_normalization else None
        rad_sink = e.grad_sink.get("radiometricT_enc.weight") if (rad is not None and e.grad_sink is not None) else None
        outs.append(ops._EpilogueFn.apply(comp, geo, ppr, sc_ppr, chunk_rays.img_idx, eval, rad, e.n_images, rad_sink))
    out = torch.cat(outs, dim=0) if len(outs) != 1 else outs[0]
    return out, n_rendering_samples, rays_shape


def render_image(
    radiance_field: torch.nn.Module,
    occupancy_grid,
    rays: SatRays,
    scene_aabb: torch.Tensor,
    args,
    epoch_idx: Optional[int] = None,
    chunk: int = 5120,
    near_plane: Optional[float] = None,
    far_plane: Optional[float] = None,
    render_step_size: float = 1e-3,
    render_bkgd: Optional[torch.Tensor] = None,
    cone_angle: float = 0.0,
    alpha_thre: float = 0.0,
    early_stop_eps: float = 0.0,
    timestamps: Optional[torch.Tensor] = None,
    only_depth: bool = False,
    eval: bool = False,
    uniforms=None,
    z_steps=None,
    static=False,
    counters=None,
):
    """sat_rendering.py:176-335 -> (dict of 12 [..., C] tensors (or {"depth"}), n_rendering_samples).
    `uniforms`: optional list with one dict per chunk {u_cam, u_sun, u_cam2} replacing the device RNG.
    `static=True` (fused precision only): the sync-free form used under CUDA-graph capture.  Sample counts never reach
    the host: packed arrays keep their B*(n-1) capacity and the kernels read P / Q on the device, the reference's
    "some ray kept no sample -> draw again" branch becomes a device-side condition (its uniforms are always drawn), and
    n_rendering_samples comes back as a 0-d int64 device tensor.
    `counters`: optional dict; receives "n_sun_samples" = kept sun-ray samples (the FLOP count of the shadow pass follows it)."""
    out, n_rendering_samples, rays_shape = render_packed(
        radiance_field, occupancy_grid, rays, scene_aabb, args, epoch_idx=epoch_idx, chunk=chunk, near_plane=near_plane,
        far_plane=far_plane, render_step_size=render_step_size, render_bkgd=render_bkgd, cone_angle=cone_angle, alpha_thre=alpha_thre,
        early_stop_eps=early_stop_eps, timestamps=timestamps, only_depth=only_depth, eval=eval, uniforms=uniforms, z_steps=z_steps,
        static=static, counters=counters)
    if only_depth:
        return {"depth": out.view((*rays_shape[:-1], -1))}, n_rendering_samples
    return {k: out[:, a:b].view((*rays_shape[:-1], -1)) for k, a, b in OUT_SLICES}, n_rendering_samples
