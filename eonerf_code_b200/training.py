"""One EO-NeRF training step with the reference's semantics (/root/reference/train_eonerf.py:99-161): render_image on a
batch of rays -> MSE (epoch < 2) or uncertainty-aware loss -> backward -> Adam(lr=5e-4).  GradScaler(1) in the reference
is a no-op scale (train_eonerf.py:58,158-160) and is not reproduced.  Data parallel: gradients are averaged over ranks
(equal shards, batch-mean losses => identical to the single-GPU gradient of the global batch)."""
import torch

from . import metrics, sat_rendering
from .datasets.satellite import define_satrays_from_tensors
from .parallel import FlatGrads


class TrainStep:
    def __init__(self, radiance_field, n_samples=128, chunk=None, lr=5e-4, world=1):
        self.field = radiance_field
        self.n_samples = n_samples
        self.render_step_size = (torch.tensor(2.0) / n_samples).item()      # fp32 quotient (train_eonerf.py:50-53)
        self.chunk = chunk
        self.world = world
        self.optimizer = torch.optim.Adam(radiance_field.parameters(), lr=lr)
        self.grads = FlatGrads(radiance_field.parameters())

    def __call__(self, rays, ts, pixels, epoch_idx):
        """rays [B,11], ts [B,1] int64, pixels [B,3] on the device -> (loss tensor, n_rendering_samples)."""
        self.field.train()
        sat = define_satrays_from_tensors(rays, ts)
        res, n_rendered = sat_rendering.render_image(self.field, None, sat, None, None, epoch_idx=epoch_idx,
                                                     chunk=self.chunk or rays.shape[0], render_step_size=self.render_step_size)
        if n_rendered == 0:                                                  # train_eonerf.py:135-136
            return None, 0
        if epoch_idx < 2:
            loss = metrics.mse(pixels, res["rgb"])
        else:
            loss, _ = metrics.uncertainty_aware_loss(pixels, res["rgb"], res["beta"])
        self.grads.zero()
        loss.backward()
        self.grads.all_reduce_mean(self.world)
        self.optimizer.step()
        return loss.detach(), n_rendered
