"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np
import torch

from oracle import eonerf_oracle as O


def fingerprint(p):
    return np.array([float(v.double().sum()) for v in p.values()] + [float(v.double().abs().sum()) for v in p.values()])


def make_model(p, n_img, device, precision="fp32", radiometric=True):
    """EONerfMLP (CUDA product module) loaded with the oracle's parameter dict."""
    from eonerf_code_b200.radiance_fields import EONerfMLP
    m = EONerfMLP(n_img, radiometric_normalization=radiometric, precision=precision)
    missing, unexpected = m.load_state_dict(p, strict=False)
    assert not unexpected and all("scales" in k for k in missing), (missing, unexpected)
    return m.to(device)


def t(a, device=None, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x if device is None else x.to(device)


def rel_err(a, b, floor=1e-6):
    """max |a-b| / max(|b|, floor-scaled magnitude): a relative error robust to zeros."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(float(b.abs().max()), floor)
    return float((a - b).abs().max()) / scale


def close(a, b, rtol, atol=0.0):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{int(bad.sum())}/{bad.numel()} mismatches, worst |d|={float(err.max()):.3e} at ref={float(b.flatten()[err.argmax()]):.3e} (rtol={rtol}, atol={atol})"


def satrays(rays, ts):
    return O.satrays_from_table(rays, ts)
