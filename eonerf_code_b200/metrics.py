"""Losses of the training step (/root/reference/metrics.py:9-22,68-69).

`uncertainty_aware_loss`, `mse`, `psnr` keep the reference's signatures (torch element-wise work on the per-ray outputs).
`packed_loss` is what `training.TrainStep` uses: the same two losses evaluated on the packed [B,21] output of the
renderer by ONE pair of sm_100a kernels that also write d loss / d out (`eonerf_loss_fwd_bwd`), instead of ~25
element-wise / reduction / slice-backward launches."""
import ctypes as C

import torch

from . import _capi as K


class _PackedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, gt_rgb, mode):
        if not out.is_cuda:
            raise RuntimeError("packed_loss runs on sm_100 GPUs only (there is no CPU fallback)")
        out = out.contiguous()
        gt = gt_rgb.to(torch.float32).contiguous()
        B = out.shape[0]
        if out.shape[1] != K.OUT_COLS or gt.shape != (B, 3):
            raise RuntimeError("packed_loss: expected out [B,21] and gt_rgb [B,3]")
        terms = torch.empty(3, dtype=torch.float32, device=out.device)
        g_out = torch.empty_like(out)
        partials = torch.empty(K.lib().eonerf_loss_partials(B), dtype=torch.float32, device=out.device)
        a = K.LossArgs(out.data_ptr(), gt.data_ptr(), B, mode, terms.data_ptr(), g_out.data_ptr(), partials.data_ptr())
        K.call("loss_fwd_bwd", a, torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(g_out)
        ctx.mark_non_differentiable(terms)
        return terms[0], terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        (g_out,) = ctx.saved_tensors
        return g_out * g_loss, None, None


def packed_loss(out, gt_rgb, epoch_idx):
    """out: packed [B,21] renderer output (sat_rendering.render_packed).  -> (loss, dict) as the reference's
    uncertainty_aware_loss (epoch_idx >= 2) or its epoch < 2 MSE (train_eonerf.py:139-143)."""
    loss, terms = _PackedLossFn.apply(out, gt_rgb, 0 if epoch_idx < 2 else 1)
    if epoch_idx < 2:
        return loss, {"loss": loss}
    return loss, {"loss": loss, "coarse_color": terms[1], "coarse_logbeta": terms[2]}


def uncertainty_aware_loss(gt_rgb, rgb, beta):
    """metrics.py:17-22 -> (loss, dict of terms)."""
    color = ((rgb - gt_rgb) ** 2 / (2 * beta ** 2)).mean()
    logbeta = (3 + torch.log(beta).mean()) / 2
    loss = color + logbeta
    return loss, {"loss": loss, "coarse_color": color, "coarse_logbeta": logbeta}


def mse(gt_rgb, rgb):
    return torch.mean((rgb - gt_rgb) ** 2)


def psnr(gt_rgb, rgb):
    return -10.0 * torch.log10(mse(gt_rgb, rgb))
