"""Losses of the training step (/root/reference/metrics.py:9-22,68-69).

`uncertainty_aware_loss`, `mse`, `psnr` keep the reference's signatures (torch element-wise work on the per-ray outputs).
`packed_loss` is what `training.TrainStep` uses: the same two losses evaluated on the packed [B,21] output of the
renderer by ONE pair of sm_100a kernels that also write d loss / d out (`eonerf_loss_fwd_bwd`), instead of ~25
element-wise / reduction / slice-backward launches."""
import ctypes as C

import torch

from . import _capi as K
from .ops import on_tensor_device


class _PackedLossFn(torch.autograd.Function):
    @staticmethod
    @on_tensor_device
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
    @on_tensor_device
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


# ---- optional prior terms of the training loss (train_eonerf.py:144-156); element-wise torch on per-ray tensors ----------
def update_loss_with_aux_term(loss, loss_dict, aux_loss, aux_dict, epoch, start_epoch=0, end_epoch=float("inf")):
    """metrics.py:9-15: the auxiliary term counts only while start_epoch <= epoch < end_epoch; its entries are always logged."""
    if start_epoch <= epoch < end_epoch:
        loss = loss + aux_loss
    loss_dict.update(aux_dict)
    return loss, loss_dict


def depth_loss_L2(gt_depth, pred_depth, gt_conf=None, w=100):
    """metrics.py:24-31: w * mean squared depth error over the rays that have a prior (gt_depth >= 0) and, when
    confidences are given, a confidence of at least 4."""
    keep = gt_depth >= 0
    if gt_conf is not None:
        keep = keep & (gt_conf >= 4)
    term = w * torch.mean(torch.square(pred_depth[keep] - gt_depth[keep]))
    return term, {"depth_l2": term, "depth_weight": w}


def shadow_loss_L2(smask, geo_shadows, epoch=None):
    """metrics.py:36-58: squared difference to the prior shadow mask on the pixels the prior marks as shadow (smask <= 0.5),
    averaged over them and weighted by their share of the batch; also reports the fraction of pixels rendered lit
    (> 0.2) that the prior calls shadow."""
    in_shadow = smask <= 0.5
    lit_but_shadow = (geo_shadows > 0.2) & (smask < 0.5)
    frac_to_penalize = lit_but_shadow.sum(dim=0) / torch.ones_like(smask).sum(dim=0)
    mean_sq = torch.sum(in_shadow * torch.square(geo_shadows - smask)) / (torch.sum(in_shadow) + 1e-6)
    term = torch.sum(in_shadow) / torch.sum(smask >= 0) * mean_sq
    return term, {"shadows_term1": term, "shadow_vals_to_penalize": frac_to_penalize}
