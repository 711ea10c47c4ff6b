"""Losses of the training step (/root/reference/metrics.py:9-22,68-69).  O(B) element-wise work on the per-ray
outputs: left to torch (SURVEY.md §8f N2 lists a fused loss + Adam as "next")."""
import torch


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
