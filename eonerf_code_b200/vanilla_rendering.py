"""BASELINE configs[1]: the vanilla-NeRF renderer `train_mlp_nerf.py` expects (/root/reference/train_mlp_nerf.py:155-170,
223-233).  Upstream this is `render_image_with_occgrid` of nerfacc v0.5.2's examples/utils.py, which the reference imports
from a module (`utils2`, train_mlp_nerf.py:17) that is MISSING from the repository: the entry point cannot run as shipped and
nothing here can be pinned against it.  Implemented: the helper's signature and the conventions of
`nerfacc.rendering` on top of the sm_100a kernels —

    samples   uniform marching of render_step_size inside the estimator's box (csrc/march.cu): the occupancy-free limit of
              estimator.sampling; one stratified offset per ray while training
    field     VanillaNeRFRadianceField (mlp.py:211-250): sigma = relu(.), rgb = sigmoid(.), view-direction conditioned
    weights   render_weight_from_density, NO 1e10 last interval (sum of weights = opacity < 1)
    outputs   colors = sum w rgb + bkgd (1 - opacity);  opacities = sum w;  depths = sum w (t_s + t_e)/2 / max(opacities, eps)
"""
from collections import namedtuple

import torch

from . import nerfacc_compat as nf
from . import ops

Rays = namedtuple("Rays", ("origins", "viewdirs"))            # /root/reference/datasets/utils.py:6


def _aabb_floats(estimator):
    """The estimator's first box as Python floats, read from the device once per version of the buffer (the marcher takes the box by
    value: a per-step .tolist() would be a host synchronisation, and is illegal inside a graph capture)."""
    if estimator is None:
        return [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
    key = (estimator.aabbs.data_ptr(), estimator.aabbs._version)
    cached = getattr(estimator, "_aabb_floats", None)
    if cached is None or cached[0] != key:
        cached = (key, [float(v) for v in estimator.aabbs[0].tolist()])
        estimator._aabb_floats = cached
    return cached[1]


def render_image_with_occgrid(radiance_field, estimator, rays, near_plane=0.0, far_plane=1e10, render_step_size=1e-3,
                              render_bkgd=None, cone_angle=0.0, alpha_thre=0.0, test_chunk_size=8192, jitter=None, static=False):
    """-> (rgb[...,3], acc[...,1], depth[...,1], n_rendering_samples).  `jitter` [B] in [0,1) replaces the device RNG.
    static=True (precision "bf16_fused"): no host synchronisation — the sample count stays on the device (buffers have the worst-case
    capacity) and n_rendering_samples comes back as a 0-d device tensor: the whole step can be captured in a CUDA graph."""
    if cone_angle != 0.0 or alpha_thre != 0.0:
        raise NotImplementedError("cone_angle / alpha_thre pruning needs the occupancy-grid traversal of nerfacc (not in the reference)")
    shape = rays.origins.shape
    o, d = rays.origins.reshape(-1, 3), rays.viewdirs.reshape(-1, 3)
    n = o.shape[0]
    aabb = _aabb_floats(estimator)
    e = radiance_field._engine()
    chunk = n if radiance_field.training else test_chunk_size
    outs, total = [], 0
    for i in range(0, n, chunk):
        oc, dc = o[i:i + chunk], d[i:i + chunk]
        B = oc.shape[0]
        jit = None
        if radiance_field.training:                                        # stratified=radiance_field.training upstream
            jit = jitter[i:i + chunk] if jitter is not None else torch.rand(B, device=oc.device)
        n_dev = None
        if static:
            ri, ts, te, offs, n_dev = ops.march_aabb(oc, dc, aabb, near_plane, far_plane, render_step_size, jit, static=True)
            total = total + n_dev[0]
        else:
            ri, ts, te, offs = ops.march_aabb(oc, dc, aabb, near_plane, far_plane, render_step_size, jit)
            total += ts.numel()
        sigma, rgb, z = ops._VanillaRaysFn.apply(torch.is_grad_enabled(), e, oc, dc, ri, ts, te, n_dev, *e.tensors())
        w, _, _ = ops._WeightsFn.apply(ts, te, sigma.squeeze(-1), offs)
        colors = ops._AccumFn.apply(w, rgb, offs)
        opac = ops._AccumFn.apply(w, None, offs)
        depth = ops._AccumFn.apply(w, z[:, None], offs) / opac.clamp_min(torch.finfo(torch.float32).eps)
        if render_bkgd is not None:
            colors = colors + render_bkgd * (1.0 - opac)
        outs.append((colors, opac, depth))
    rgb, acc, depth = (torch.cat(x, 0) if len(outs) > 1 else x[0] for x in zip(*outs))
    return rgb.view(*shape[:-1], -1), acc.view(*shape[:-1], -1), depth.view(*shape[:-1], -1), total
