"""The nerfacc v0.5.2 surface the reference imports (`from nerfacc import ...`:
/root/reference/radiance_fields/eonerf.py:15, /root/reference/sat_rendering.py:2, /root/reference/train_eonerf.py:22),
served by the sm_100a kernels in csrc/render.cu.  Signatures follow nerfacc/volrend.py at tag v0.5.2."""
import torch

from . import ops


def _offsets(packed_info, ray_indices, n_rays, n_pts, device):
    if packed_info is not None:                       # [n_rays, 2] (start, count)
        offs = torch.zeros(packed_info.shape[0] + 1, dtype=torch.int64, device=device)
        offs[1:] = torch.cumsum(packed_info[:, 1].long(), 0)
        return offs
    if ray_indices is None or n_rays is None:
        raise ValueError("give packed_info, or ray_indices and n_rays")
    return ops.pack_info(ray_indices, n_rays)


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                      prefix_trans=None):
    offs = _offsets(packed_info, ray_indices, n_rays, t_starts.numel(), t_starts.device)
    _, trans, alphas = ops._WeightsFn.apply(t_starts, t_ends, sigmas, offs)
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans, alphas


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                               prefix_trans=None):
    offs = _offsets(packed_info, ray_indices, n_rays, t_starts.numel(), t_starts.device)
    weights, trans, alphas = ops._WeightsFn.apply(t_starts, t_ends, sigmas, offs)
    if prefix_trans is not None:
        trans = trans * prefix_trans
        weights = trans * alphas
    return weights, trans, alphas


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    if ray_indices is None:
        raise ValueError("the flattened-samples form needs ray_indices and n_rays")
    offs = ops.pack_info(ray_indices, n_rays)
    return ops._AccumFn.apply(weights, values, offs)


def pack_info(ray_indices, n_rays=None):
    """[n_rays, 2] (start, count) like nerfacc/pack.py."""
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    offs = ops.pack_info(ray_indices, n_rays)
    return torch.stack([offs[:-1], offs[1:] - offs[:-1]], 1)


class OccGridEstimator(torch.nn.Module):
    """Checkpoint-compatible shim.  The reference constructs, updates and state_dict()s the grid but every
    `.sampling` call site is commented out (sat_rendering.py:92,94,234,257), so its contents never reach the
    renderer.  Buffers follow nerfacc v0.5.2 so `occ_grid_state_dict` round-trips (train_eonerf.py:187,
    eval_eonerf.py:69-71).  `update_every_n_steps` is a no-op: the 2.1 M-point density pass it would run
    produces a result nobody reads."""

    def __init__(self, roi_aabb, resolution=128, levels=1, **kwargs):
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * 3
        res = torch.tensor(resolution, dtype=torch.int32)
        aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        aabbs = torch.stack([torch.cat([(aabb[:3] + aabb[3:]) / 2 - (aabb[3:] - aabb[:3]) / 2 * 2 ** i,
                                        (aabb[:3] + aabb[3:]) / 2 + (aabb[3:] - aabb[:3]) / 2 * 2 ** i]) for i in range(levels)])
        n_cells = int(res.prod())
        self.levels, self.cells_per_lvl = levels, n_cells
        self.register_buffer("resolution", res)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * n_cells))
        self.register_buffer("binaries", torch.zeros([levels] + list(resolution), dtype=torch.bool))

    @property
    def device(self):
        return self.aabbs.device

    def update_every_n_steps(self, step, occ_eval_fn=None, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16):
        return None
