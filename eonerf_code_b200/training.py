"""One EO-NeRF training step with the reference's semantics (/root/reference/train_eonerf.py:99-161): render_image on a
batch of rays -> MSE (epoch < 2) or uncertainty-aware loss -> backward -> Adam(lr=5e-4).  GradScaler(1) in the reference
is a no-op scale (train_eonerf.py:58,158-160) and is not reproduced.  Data parallel: gradients are averaged over ranks
(equal shards, batch-mean losses => identical to the single-GPU gradient of the global batch).

Two ways to run the same step:

* eager (`graph=False`): the reference's control flow — the host reads the sample counts P and Q (two syncs per step,
  the reference has more) and returns n_rendering_samples as an int.
* CUDA graph (`graph=True`, fused precision): the whole step is sync-free (`render_image(static=True)`: sample counts
  stay on the device), so it is captured once and replayed: ~200 kernel launches and all of Python leave the step.
  With world > 1 the NCCL all-reduce stays outside the graphs (render + backward | all-reduce | Adam); EONERF_GRAPH_NCCL=1 captures it
  inside one graph instead (measured: no gain)."""
import os

import torch

from . import _capi as K
from . import metrics, sat_rendering
from .datasets.satellite import define_satrays_from_tensors
from .optim import FlatAdam
from .parallel import FlatGrads


class TrainStep:
    def __init__(self, radiance_field, n_samples=128, chunk=None, lr=5e-4, world=1, graph=False, micro_batch=None):
        """micro_batch: rays per forward+backward pass.  A batch larger than that is processed in slices whose gradients
        accumulate in the flat buffer before ONE optimiser step (the losses are batch means, so the sum of the slice losses
        weighted by slice size is the loss of the whole batch): BASELINE configs[4]'s 65 536-ray batch keeps ~170 GB of
        activations alive if rendered in one piece; 16 384-ray slices need a quarter of that."""
        self.field = radiance_field
        self.micro_batch = micro_batch
        self.n_samples = n_samples
        self.render_step_size = (torch.tensor(2.0) / n_samples).item()      # fp32 quotient (train_eonerf.py:50-53)
        self.chunk = chunk
        self.world = world
        self.graph = graph
        if graph and radiance_field.precision != "bf16_fused":
            raise RuntimeError("graph=True needs precision='bf16_fused' (device-side sample counts)")
        params = list(radiance_field.parameters())
        self.grads = FlatGrads(params)
        # torch.optim.Adam's state layout, one kernel over flat buffers (device-side step counter: graph-capturable)
        self.optimizer = FlatAdam(params, self.grads.flat, lr=lr)
        # backward kernels accumulate straight into the flat gradient buffer (no per-parameter autograd adds)
        radiance_field._engine().use_grad_sink({k: p.grad for k, p in radiance_field.named_parameters()})
        self._graphs = {}
        self._first_done = False
        self.launches_per_step = None           # kernels of libeonerf_b200 inside one captured step
        self.n_rendered_total = None            # graph mode: running sum of n_rendering_samples, on the device
        self.n_sun_total = None                 # graph mode: running sum of the kept sun-ray samples, on the device

    # ------------------------------------------------------------------------------------------------------------------
    def _forward_backward(self, rays, ts, pixels, epoch_idx, static, uniforms=None):
        """uniforms: optional {u_cam, u_sun, u_cam2} [B,n] tensors replacing the device RNG (parity tests)."""
        self.field.train()
        if static or self.graph:               # graph replays update the parameters behind Python's back
            self.field._engine().refresh_prepared()
        self.grads.zero()
        B = rays.shape[0]
        mb = self.micro_batch or B
        total_loss, total_rendered = None, 0
        stats = {}
        for b0 in range(0, B, mb):
            sl = slice(b0, min(B, b0 + mb))
            sat = define_satrays_from_tensors(rays[sl], ts[sl])
            us = None if uniforms is None else [{k: v[sl] for k, v in uniforms.items()}]
            out, n_rendered, _ = sat_rendering.render_packed(self.field, None, sat, None, None, epoch_idx=epoch_idx,
                                                             chunk=self.chunk or (sl.stop - sl.start), render_step_size=self.render_step_size,
                                                             static=static, uniforms=us, counters=stats)
            if not static and n_rendered == 0:                               # train_eonerf.py:135-136
                continue
            loss, _ = metrics.packed_loss(out, pixels[sl], epoch_idx)       # MSE (epoch < 2) / uncertainty-aware loss, value + gradient
            if mb < B:
                loss = loss * ((sl.stop - sl.start) / B)
            loss.backward()
            total_loss = loss.detach() if total_loss is None else total_loss + loss.detach()
            total_rendered = total_rendered + n_rendered
        if total_loss is None:
            return None, 0
        self.last_sun_samples = stats.get("n_sun_samples", 0)     # int (eager) or 0-d device tensor (static)
        if static and self.n_sun_total is not None and torch.is_tensor(self.last_sun_samples):
            self.n_sun_total += self.last_sun_samples
        return total_loss, total_rendered

    def _update(self, averaged):
        self.optimizer.step(grad_scale=1.0 if (averaged or self.world == 1) else 1.0 / self.world)

    def eager(self, rays, ts, pixels, epoch_idx, uniforms=None):
        """rays [B,11], ts [B,1] int64, pixels [B,3] on the device -> (loss tensor, n_rendering_samples int)."""
        loss, n_rendered = self._forward_backward(rays, ts, pixels, epoch_idx, static=False, uniforms=uniforms)
        if loss is None:
            return None, 0
        self.grads.all_reduce_mean(self.world)
        self._update(averaged=True)
        return loss, n_rendered

    # ------------------------------------------------------------------------------------------------------------------
    def _capture(self, rays, ts, pixels, epoch_idx, uniforms=None):
        g = {"rays": rays.clone(), "ts": ts.clone(), "pixels": pixels.clone(),
             "uniforms": None if uniforms is None else {k: v.clone() for k, v in uniforms.items()}}
        lib = K.lib()
        before = int(lib.eonerf_launch_count(0))
        # world > 1, EONERF_GRAPH_NCCL=1: the flat-gradient all-reduce is captured INSIDE the graph (NCCL collectives are capturable): one
        # launch per step.  Measured at N = 2: 8.59 ms per step against 8.56 ms for the default form (graph | host-issued all-reduce |
        # graph) — no gain, so the default stays; tools/check_graph_nccl.py checks both against the eager data-parallel step.
        fuse_nccl = self.world > 1 and os.environ.get("EONERF_GRAPH_NCCL", "0") == "1"
        if fuse_nccl:
            torch.distributed.all_reduce(torch.zeros(1, device=rays.device))      # communicator set-up happens outside the capture
            torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, capture_error_mode="thread_local"):
            loss, n = self._forward_backward(g["rays"], g["ts"], g["pixels"], epoch_idx, static=True, uniforms=g["uniforms"])
            self.n_rendered_total += n
            if self.world == 1:
                self._update(averaged=True)
            elif fuse_nccl:
                torch.distributed.all_reduce(self.grads.flat)
                self._update(averaged=False)
        g["g1"], g["loss"], g["n"] = g1, loss, n
        if self.world > 1 and not fuse_nccl:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, capture_error_mode="thread_local"):
                self._update(averaged=False)
            g["g2"] = g2
        self.launches_per_step = int(lib.eonerf_launch_count(0)) - before
        return g

    def _params_changed(self):
        """A graph replay ran Adam behind Python's back: bump the parameters' version counters so that every consumer keyed
        on them (ops.FieldEngine.prepared(): the operand-layout weights an eval render_image uses) sees the new values."""
        for p in self.optimizer._params:
            torch.autograd.graph.increment_version(p)

    def __call__(self, rays, ts, pixels, epoch_idx, uniforms=None):
        """Eager: as `eager`.  Graph mode: -> (loss, n_rendering_samples) as 0-d device tensors that the next call
        overwrites; the very first call runs eagerly (it also initialises the optimiser state), the second captures.
        uniforms: optional {u_cam, u_sun, u_cam2} [B,n] replacing the device RNG (parity tests); in graph mode they are
        copied into the captured step's own buffers like the batch."""
        if not self.graph:
            return self.eager(rays, ts, pixels, epoch_idx, uniforms=uniforms)
        if self.n_rendered_total is None:
            self.n_rendered_total = torch.zeros((), dtype=torch.int64, device=rays.device)
            self.n_sun_total = torch.zeros((), dtype=torch.int64, device=rays.device)
        if not self._first_done:
            self._first_done = True
            loss, n = self._forward_backward(rays, ts, pixels, epoch_idx, static=True, uniforms=uniforms)
            self.n_rendered_total += n
            if self.world > 1:
                torch.distributed.all_reduce(self.grads.flat)
            self._update(averaged=False)
            return loss, n
        key = (epoch_idx >= 2, tuple(rays.shape), rays.device, uniforms is not None)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = self._capture(rays, ts, pixels, epoch_idx, uniforms)
        g["rays"].copy_(rays, non_blocking=True)
        g["ts"].copy_(ts, non_blocking=True)
        g["pixels"].copy_(pixels, non_blocking=True)
        if uniforms is not None:
            for k, v in g["uniforms"].items():
                v.copy_(uniforms[k], non_blocking=True)
        self.optimizer.sync_hyper()            # the captured Adam reads lr on the device: follow schedulers between replays
        g["g1"].replay()
        if "g2" in g:
            torch.distributed.all_reduce(self.grads.flat)
            g["g2"].replay()
        self._params_changed()
        return g["loss"], g["n"]


class VanillaTrainStep:
    """One training step of BASELINE configs[1] (/root/reference/train_mlp_nerf.py:155-190): render_image_with_occgrid on a batch of
    pinhole rays -> smooth-L1 against the pixels -> backward -> Adam(lr).  Eager (torch control flow, one host read of the sample
    count) or, with precision "bf16_fused", sync-free and replayed as ONE CUDA graph like the EO-NeRF step (the marcher leaves the
    sample count on the device, the flat Adam keeps its step counter there)."""

    def __init__(self, radiance_field, estimator, render_step_size=5e-3, near_plane=0.0, far_plane=1e10, render_bkgd=None, lr=5e-4,
                 graph=False):
        from .vanilla_rendering import Rays, render_image_with_occgrid
        self._Rays, self._render = Rays, render_image_with_occgrid
        self.field, self.estimator = radiance_field, estimator
        self.step_size, self.near, self.far, self.bkgd = render_step_size, near_plane, far_plane, render_bkgd
        self.graph = graph
        if graph and radiance_field.precision != "bf16_fused":
            raise RuntimeError("graph=True needs precision='bf16_fused' (device-side sample counts)")
        params = list(radiance_field.parameters())
        self.grads = FlatGrads(params)
        self.optimizer = FlatAdam(params, self.grads.flat, lr=lr)
        radiance_field._engine().use_grad_sink({k: p.grad for k, p in radiance_field.named_parameters()})
        self._graphs = {}

    def _forward_backward(self, origins, viewdirs, pixels, static, jitter=None):
        self.field.train()
        if static:
            self.field._engine().refresh_prepared()
        self.grads.zero()
        rgb, acc, depth, n = self._render(self.field, self.estimator, self._Rays(origins, viewdirs), near_plane=self.near, far_plane=self.far,
                                          render_step_size=self.step_size, render_bkgd=self.bkgd, jitter=jitter, static=static)
        loss = torch.nn.functional.smooth_l1_loss(rgb, pixels)                    # train_mlp_nerf.py:183
        loss.backward()
        return loss.detach(), n

    def eager(self, origins, viewdirs, pixels, jitter=None):
        loss, n = self._forward_backward(origins, viewdirs, pixels, static=False, jitter=jitter)
        self.optimizer.step()
        return loss, n

    def __call__(self, origins, viewdirs, pixels, jitter=None):
        """-> (loss, n_rendering_samples).  Graph mode: both are 0-d device tensors that the next call overwrites."""
        if not self.graph:
            return self.eager(origins, viewdirs, pixels, jitter)
        key = (tuple(origins.shape), jitter is not None)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = {"o": origins.clone(), "d": viewdirs.clone(), "px": pixels.clone(),
                                     "jit": None if jitter is None else jitter.clone()}
            self._forward_backward(g["o"], g["d"], g["px"], static=True, jitter=g["jit"])      # warm-up outside the capture (lazy init)
            self.grads.zero()
            torch.cuda.synchronize()
            torch.cuda.empty_cache()               # the capture allocates its own (worst-case sized) buffers from a private pool
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                loss, n = self._forward_backward(g["o"], g["d"], g["px"], static=True, jitter=g["jit"])
                self.optimizer.step()
            g["graph"], g["loss"], g["n"] = graph, loss, n
        g["o"].copy_(origins, non_blocking=True)
        g["d"].copy_(viewdirs, non_blocking=True)
        g["px"].copy_(pixels, non_blocking=True)
        if jitter is not None:
            g["jit"].copy_(jitter, non_blocking=True)
        self.optimizer.sync_hyper()
        g["graph"].replay()
        for p in self.optimizer._params:                 # Adam ran behind Python's back: see TrainStep._params_changed
            torch.autograd.graph.increment_version(p)
        return g["loss"], g["n"]
