"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the three nerfacc v0.5.2 volume-rendering
operators the EO-NeRF hot path calls.

nerfacc is a third-party dependency of the reference, pinned at v0.5.2
(/root/reference/setup_env.sh:10) and *not vendored* under /root/reference; it is not installed in
this image and there is no network.  This file restates its published algorithm
(nerfacc/volrend.py, nerfacc/scan.py, nerfacc/pack.py at tag v0.5.2):

    sigmas_dt = sigmas * (t_ends - t_starts)
    alphas    = 1 - exp(-sigmas_dt)
    trans     = exp(-exclusive_sum_per_ray(sigmas_dt))
    weights   = trans * alphas
    accumulate_along_rays(w, v)[r, c] = sum_{i in ray r} w_i * v_{i, c}       (index_add_)

Call sites in the reference that anchor the semantics:
    radiance_fields/eonerf.py:186-193 (render_depth), :229-242 (rendering),
    sat_rendering.py:106-110 (sun-ray transmittance).

PARITY STATUS: "parity unpinned" for this file in isolation — upstream nerfacc has its own tests
(tests/test_scan.py, tests/test_rendering.py) but they are not available here and the reference
repository holds no golden vectors.  What pins it instead: the dense dead-code twin
`weights_from_sigma` at radiance_fields/eonerf.py:37-54 (checked in tests/test_oracle.py) and the
analytic identity sum(w) == 1 when the last interval is 1e10 (radiance_fields/eonerf.py:214-220).

The same module is installed as the `nerfacc` stand-in when the *real* reference Python is imported
in this container to generate tests/golden/ (see oracle/ref_harness.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (eonerf_code_b200/) never does.
"""
import torch


def _segment_positions(ray_indices: torch.Tensor, n_rays: int):
    """ray_indices is sorted (samples are packed ray after ray). Returns (counts[n_rays], pos[n_pts])."""
    counts = torch.bincount(ray_indices, minlength=n_rays)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(ray_indices.numel(), device=ray_indices.device) - starts[ray_indices]
    return counts, pos


class _ExclusiveSum(torch.autograd.Function):
    """True per-ray exclusive prefix sum (nerfacc/scan.py::exclusive_sum); backward is the reverse
    exclusive sum of the incoming gradient.  NOT computed as inclusive - self: the last interval of
    every ray is ~1e10*sigma and that form cancels catastrophically."""

    @staticmethod
    def forward(ctx, x, ray_indices, n_rays):
        counts, pos = _segment_positions(ray_indices, n_rays)
        ctx.save_for_backward(ray_indices, pos)
        ctx.n_rays = n_rays
        ctx.width = int(counts.max()) if counts.numel() else 0
        return _ExclusiveSum._scan(x, ray_indices, pos, n_rays, ctx.width, reverse=False)

    @staticmethod
    def _scan(x, ray_indices, pos, n_rays, width, reverse):
        if x.numel() == 0:
            return x.clone()
        pad = x.new_zeros(n_rays, width + 1)
        pad[ray_indices, pos + (0 if reverse else 1)] = x
        if reverse:
            # suffix sums excluding self: out[i] = sum_{j>i} x[j]
            inc = torch.flip(torch.cumsum(torch.flip(pad, [1]), 1), [1])
            return inc[ray_indices, pos + 1]
        inc = torch.cumsum(pad, 1)
        return inc[ray_indices, pos]

    @staticmethod
    def backward(ctx, g):
        ray_indices, pos = ctx.saved_tensors
        return _ExclusiveSum._scan(g.contiguous(), ray_indices, pos, ctx.n_rays, ctx.width, reverse=True), None, None


def exclusive_sum(x, ray_indices, n_rays):
    return _ExclusiveSum.apply(x, ray_indices, n_rays)


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None,
                                      n_rays=None, prefix_trans=None):
    sigmas_dt = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sigmas_dt)
    trans = torch.exp(-exclusive_sum(sigmas_dt, ray_indices, n_rays))
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans, alphas


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None,
                               n_rays=None, prefix_trans=None):
    trans, alphas = render_transmittance_from_density(
        t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=n_rays, prefix_trans=prefix_trans)
    return trans * alphas, trans, alphas


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    src = weights[..., None] if values is None else weights[..., None] * values
    out = torch.zeros((n_rays, src.shape[-1]), device=src.device, dtype=src.dtype)
    return out.index_add_(0, ray_indices, src)


class OccGridEstimator(torch.nn.Module):
    """The reference only constructs / updates / state_dict()s the grid (train_eonerf.py:74,112-119,187);
    every `.sampling` call site is commented out (sat_rendering.py:92,94,234,257)."""

    def __init__(self, roi_aabb=None, resolution=128, levels=1, **kw):
        super().__init__()
        self.register_buffer("aabbs", torch.tensor([roi_aabb if roi_aabb is not None else [-1.] * 3 + [1.] * 3],
                                                   dtype=torch.float32))

    @property
    def device(self):
        return self.aabbs.device

    def update_every_n_steps(self, *a, **kw):
        return None


def install_as_nerfacc():
    """Register this module as `nerfacc` / `nerfacc.volrend` so the unmodified reference imports."""
    import sys
    import types
    me = sys.modules[__name__]
    top = types.ModuleType("nerfacc")
    for name in ("render_transmittance_from_density", "render_weight_from_density",
                 "accumulate_along_rays", "OccGridEstimator", "exclusive_sum"):
        setattr(top, name, getattr(me, name))
    top.rendering = None
    top.volrend = me
    sys.modules["nerfacc"] = top
    sys.modules["nerfacc.volrend"] = me
    return top
