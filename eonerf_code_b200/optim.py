"""Adam over flat buffers (csrc/optim.cu), with torch.optim.Adam's state layout.

The reference trains with `torch.optim.Adam(radiance_field.parameters(), lr=5e-4)` and checkpoints its
`optimizer_state_dict` (/root/reference/train_eonerf.py:57,158-160,185-191).  `FlatAdam` IS a torch.optim.Adam as far as
state_dict()/load_state_dict()/param_groups go (per-parameter `step`, `exp_avg`, `exp_avg_sq`), but the parameters, their
moments and (through parallel.FlatGrads) their gradients are views of four flat fp32 buffers and `step()` is one
sm_100a kernel over them instead of torch's multi-tensor launches.  No CPU fallback."""
import torch

from . import _capi as K
from .parallel import flat_offsets


class FlatAdam(torch.optim.Adam):
    def __init__(self, params, flat_grad, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, lr=lr, betas=betas, eps=eps, capturable=True)
        ref = params[0]
        if not ref.is_cuda:
            raise RuntimeError("FlatAdam runs on sm_100 GPUs only (there is no CPU fallback)")
        self._params = params
        self._offsets, total = flat_offsets(params)
        if flat_grad.numel() != total:
            raise RuntimeError("flat_grad must be parallel.FlatGrads(params).flat of the same parameter list")
        self.flat_grad = flat_grad
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=ref.device)
        self.flat_exp_avg = torch.zeros_like(self.flat_param)
        self.flat_exp_avg_sq = torch.zeros_like(self.flat_param)
        self._step_buf = torch.zeros(4, dtype=torch.float32, device=ref.device)    # [step, lr/(1-b1^t), sqrt(1-b2^t), -]
        self.step_t = self._step_buf[0]                                            # torch keeps `step` as an fp32 tensor
        # the learning rate lives on the device too: a captured step reads it there, so a scheduler / a manual
        # param_groups[0]["lr"] change reaches graph replays through sync_hyper() (train_eonerf.py:64,304: StepLR)
        self._lr_buf = torch.full((1,), float(lr), dtype=torch.float64, device=ref.device)
        self._lr_synced = float(lr)
        for p, off in zip(params, self._offsets):
            n = p.numel()
            if p.dtype != torch.float32:
                raise RuntimeError("FlatAdam: fp32 master parameters only")
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view(p.shape)                   # the module now trains inside the flat buffer
        self._bind_state()

    def _bind_state(self):
        for p, off in zip(self._params, self._offsets):
            n = p.numel()
            self.state[p] = {"step": self.step_t, "exp_avg": self.flat_exp_avg[off:off + n].view(p.shape),
                             "exp_avg_sq": self.flat_exp_avg_sq[off:off + n].view(p.shape)}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)               # torch re-creates the state tensors: copy them back into the flat views
        loaded = {p: dict(self.state[p]) for p in self._params if p in self.state and "exp_avg" in self.state[p]}
        for p, off in zip(self._params, self._offsets):
            st = loaded.get(p)
            if st is None:
                continue
            n = p.numel()
            self.flat_exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.flat_exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_t.copy_(torch.as_tensor(st["step"], dtype=torch.float32))
        self._bind_state()

    def state_dict(self):
        """torch.optim.Adam's layout with an INDEPENDENT fp32 scalar `step` per parameter: inside this optimiser all
        parameters share one device-side counter, and saving that aliasing would make a stock (non-capturable)
        torch.optim.Adam increment the shared tensor once per parameter per iteration after load_state_dict."""
        sd = super().state_dict()
        for st in sd["state"].values():
            if "step" in st:
                st["step"] = st["step"].detach().clone()
        return sd

    def sync_hyper(self):
        """Refresh the device-side learning rate from param_groups (no-op unless it changed).  `step()` calls it; a caller
        that replays a captured step calls it before the replay."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_synced:
            self._lr_buf.fill_(lr)
            self._lr_synced = lr

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        if closure is not None:
            raise RuntimeError("FlatAdam.step does not take a closure")
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        first, last = self._params[0], self._params[-1]
        if (first.data_ptr() != self.flat_param.data_ptr()
                or last.data_ptr() != self.flat_param.data_ptr() + 4 * self._offsets[-1]
                or first.grad is None or first.grad.data_ptr() != self.flat_grad.data_ptr()):
            raise RuntimeError("FlatAdam: a parameter or gradient no longer lives in the flat buffers (module moved / re-created?)")
        g = self.param_groups[0]
        a = K.AdamArgs(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.flat_exp_avg.data_ptr(),
                       self.flat_exp_avg_sq.data_ptr(), self.flat_param.numel(), self._step_buf.data_ptr(),
                       g["lr"], g["betas"][0], g["betas"][1], g["eps"], grad_scale, self._lr_buf.data_ptr())
        with torch.cuda.device(self.flat_param.device):
            K.call("adam_step", a, torch.cuda.current_stream().cuda_stream)
        for p in self._params:                            # the kernel wrote behind autograd's back: bump the version counters
            torch.autograd.graph.increment_version(p)      # (ops.FieldEngine.prepared() keys its weight cache on them)
