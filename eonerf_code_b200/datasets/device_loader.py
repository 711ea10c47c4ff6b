"""Device-resident ray table + on-GPU batch sampler.

Drop-in for `DataLoader(train_dataset, shuffle=True, batch_size=B)` over the reference's `SatelliteDataset` in training
mode (/root/reference/datasets/satellite.py:799-807, /root/reference/train_eonerf.py:70,99-109): iterating yields dicts with
the reference's keys — "rays" [B,11] fp32, "rgbs" [B,3] fp32, "ts" [B,1] int64, "idx" [B] int64 — already on the GPU.
The whole table (all_rays, all_rgbs, all_ids_img) lives in HBM, an epoch is a device-side permutation (shuffle=True draws
one permutation per epoch, without replacement; the last batch may be short, drop_last=False as in the reference), and a
batch is ONE gather kernel (`eonerf_gather_batch`) instead of B `__getitem__` calls + collate + a pageable H2D copy
(measured at 44-54 k rays/s whatever the batch size).  No CPU fallback."""
import torch

from .. import _capi as K


class DeviceRayLoader:
    def __init__(self, all_rays, all_rgbs, all_ids_img, batch_size, shuffle=True, device="cuda", generator=None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceRayLoader keeps the ray table in HBM: it needs a CUDA device (there is no CPU fallback)")
        K.require_device()
        self.all_rays = all_rays.to(dev, torch.float32).contiguous()
        self.all_rgbs = all_rgbs.to(dev, torch.float32).contiguous()
        self.all_ts = all_ids_img.to(dev).reshape(-1).long().contiguous()
        n = self.all_rays.shape[0]
        if self.all_rays.shape != (n, 11) or self.all_rgbs.shape != (n, 3) or self.all_ts.shape != (n,):
            raise ValueError("expected all_rays [N,11], all_rgbs [N,3], all_ids_img [N] or [N,1]")
        self.n, self.batch_size, self.shuffle, self.generator = n, int(batch_size), shuffle, generator
        self._identity = None

    @classmethod
    def from_dataset(cls, dataset, batch_size, **kw):
        """dataset: the reference's SatelliteDataset in training mode (attributes all_rays, all_rgbs, all_ids_img)."""
        return cls(dataset.all_rays, dataset.all_rgbs, dataset.all_ids_img, batch_size, **kw)

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def _epoch_order(self):
        if self.shuffle:
            return torch.randperm(self.n, device=self.all_rays.device, generator=self.generator)
        if self._identity is None:
            self._identity = torch.arange(self.n, device=self.all_rays.device)
        return self._identity

    def gather(self, perm, first, batch, out=None):
        """Rows perm[first:first+batch] -> {"rays","rgbs","ts","idx"}; `out` = a dict from a previous call to reuse its buffers
        (what a CUDA-graph training step wants: fixed addresses)."""
        dev = self.all_rays.device
        if out is None or out["rays"].shape[0] != batch:
            out = {"rays": torch.empty(batch, 11, dtype=torch.float32, device=dev), "rgbs": torch.empty(batch, 3, dtype=torch.float32, device=dev),
                   "ts": torch.empty(batch, 1, dtype=torch.int64, device=dev), "idx": torch.empty(batch, dtype=torch.int64, device=dev)}
        a = K.GatherBatchArgs(self.all_rays.data_ptr(), self.all_rays.stride(0), self.all_rgbs.data_ptr(), self.all_rgbs.stride(0),
                              self.all_ts.data_ptr(), perm.data_ptr(), self.n, first, batch, out["rays"].data_ptr(), out["rgbs"].data_ptr(),
                              out["ts"].data_ptr(), out["idx"].data_ptr())
        with torch.cuda.device(dev):
            K.call("gather_batch", a, torch.cuda.current_stream().cuda_stream)
        return out

    def __iter__(self):
        perm = self._epoch_order()
        for first in range(0, self.n, self.batch_size):
            yield self.gather(perm, first, min(self.batch_size, self.n - first))
