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
    """estimators/occ_grid.py of nerfacc v0.5.2, the part the reference exercises (train_eonerf.py:74,112-119,187;
    eval_eonerf.py:66-71): buffers `resolution`, `aabbs`, `occs`, `binaries` (persistent) + `grid_coords`, `grid_indices`
    (non-persistent); `update_every_n_steps` = every n-th step, EMA-max update of the cell occupancies from
    `occ_eval_fn` at one jittered point per visited cell (all cells during warm-up, then cells_per_lvl/4 uniform + up to
    as many currently-occupied ones), binary grid = occs > min(mean(occs), occ_thre).  Every `.sampling` call site of the
    reference is commented out (sat_rendering.py:92,94,234,257).  Restated from memory of the published source: unpinned."""

    def __init__(self, roi_aabb=None, resolution=128, levels=1, **kw):
        super().__init__()
        roi_aabb = [-1.] * 3 + [1.] * 3 if roi_aabb is None else roi_aabb
        resolution = [resolution] * 3 if isinstance(resolution, int) else list(resolution)
        aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        c, h = (aabb[:3] + aabb[3:]) / 2, (aabb[3:] - aabb[:3]) / 2
        self.levels = levels
        self.cells_per_lvl = int(torch.tensor(resolution).prod())
        self.register_buffer("resolution", torch.tensor(resolution, dtype=torch.int32))
        self.register_buffer("aabbs", torch.stack([torch.cat([c - h * 2 ** i, c + h * 2 ** i]) for i in range(levels)]))
        self.register_buffer("occs", torch.zeros(levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + resolution, dtype=torch.bool))
        gc = torch.stack(torch.meshgrid([torch.arange(r) for r in resolution], indexing="ij"), -1).reshape(self.cells_per_lvl, 3)
        self.register_buffer("grid_coords", gc, persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)

    @property
    def device(self):
        return self.aabbs.device

    @torch.no_grad()
    def update_every_n_steps(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16):
        if not self.training:
            raise RuntimeError("Please call estimator.train() before calling update_every_n_steps().")
        if step % n != 0:
            return
        for lvl in range(self.levels):
            if step < warmup_steps:
                idx = self.grid_indices
            else:
                k = self.cells_per_lvl // 4
                uni = torch.randint(self.cells_per_lvl, (k,), device=self.device)
                occd = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
                if k < len(occd):
                    occd = occd[torch.randint(len(occd), (k,), device=self.device)]
                idx = torch.cat([uni, occd], 0)
            gc = self.grid_coords[idx]
            x = (gc + torch.rand_like(gc, dtype=torch.float32)) / self.resolution
            x = self.aabbs[lvl, :3] + x * (self.aabbs[lvl, 3:] - self.aabbs[lvl, :3])
            occ = occ_eval_fn(x).squeeze(-1)
            cid = lvl * self.cells_per_lvl + idx
            self.occs[cid] = torch.maximum(self.occs[cid] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)


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
