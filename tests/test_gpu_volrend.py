"""GPU: the nerfacc v0.5.2 operator trio (render_transmittance/weight_from_density, accumulate_along_rays) through the
C ABI vs the oracle restatement — forward 1e-5 relative, backward against oracle autograd; ragged / empty rays and the
1e10 last interval (radiance_fields/eonerf.py:220) included."""
import pytest
import torch

from helpers import close, t
from oracle import nerfacc_v052 as nv

pytestmark = pytest.mark.gpu


def ragged(B, max_len, seed, last_inf=True):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(0, max_len + 1, (B,), generator=g)
    counts[0] = 0
    counts[-1] = max_len
    if B > 3:
        counts[2] = 1
    ri = torch.arange(B).repeat_interleave(counts)
    N = int(counts.sum())
    ts = torch.rand(N, generator=g) * 2
    te = ts + torch.rand(N, generator=g) * 0.05
    if last_inf:
        last = torch.cumsum(counts, 0)[counts > 0] - 1
        te[last] = 1e10
    sig = torch.rand(N, generator=g) * 20 * (torch.rand(N, generator=g) > 0.3)
    return ri, ts, te, sig, counts


@pytest.mark.parametrize("B,max_len,last_inf", [(64, 63, True), (333, 127, True), (50, 200, False), (1, 5, True)])
def test_weights_forward_backward(cuda, B, max_len, last_inf):
    from eonerf_code_b200 import nerfacc_compat as nc
    ri, ts, te, sig, counts = ragged(B, max_len, B + max_len, last_inf)
    g = torch.Generator().manual_seed(7)
    gw, gT, ga = (torch.randn(ts.numel(), generator=g) for _ in range(3))
    so = sig.clone().requires_grad_(True)
    w, T, a = nv.render_weight_from_density(ts, te, so, ray_indices=ri, n_rays=B)
    (w * gw + T * gT + a * ga).sum().backward()
    sc = sig.to(cuda).requires_grad_(True)
    w2, T2, a2 = nc.render_weight_from_density(ts.to(cuda), te.to(cuda), sc, ray_indices=ri.to(cuda), n_rays=B)
    (w2 * gw.to(cuda) + T2 * gT.to(cuda) + a2 * ga.to(cuda)).sum().backward()
    close(w2, w, 1e-5, 1e-7); close(T2, T, 1e-5, 1e-7); close(a2, a, 1e-5, 1e-7)
    close(sc.grad, so.grad, 2e-4, 2e-5 * float(so.grad.abs().max()))
    if last_inf:    # sum of weights == 1 for every ray whose last sigma > 0 (eonerf.py:214-220)
        s = nc.accumulate_along_rays(w2, None, ri.to(cuda), B).cpu().flatten()
        s_ref = nv.accumulate_along_rays(w.detach(), None, ri, B).flatten()
        close(s, s_ref, 1e-5, 1e-6)
    T3, a3 = nc.render_transmittance_from_density(ts.to(cuda), te.to(cuda), sig.to(cuda), ray_indices=ri.to(cuda), n_rays=B)
    close(T3, T, 1e-5, 1e-7)


@pytest.mark.parametrize("C", [1, 3, 5, 40])
def test_accumulate_forward_backward(cuda, C):
    from eonerf_code_b200 import nerfacc_compat as nc
    B = 97
    ri, ts, te, sig, counts = ragged(B, 90, 11)
    g = torch.Generator().manual_seed(C)
    N = ts.numel()
    w = torch.rand(N, generator=g)
    v = torch.randn(N, C, generator=g)
    go = torch.randn(B, C, generator=g)
    wo, vo = w.clone().requires_grad_(True), v.clone().requires_grad_(True)
    out = nv.accumulate_along_rays(wo, vo, ri, B)
    (out * go).sum().backward()
    wc, vc = w.to(cuda).requires_grad_(True), v.to(cuda).requires_grad_(True)
    out2 = nc.accumulate_along_rays(wc, vc, ri.to(cuda), B)
    (out2 * go.to(cuda)).sum().backward()
    close(out2, out, 1e-5, 1e-5)
    close(wc.grad, wo.grad, 1e-5, 1e-5)
    close(vc.grad, vo.grad, 1e-5, 1e-6)
    if C == 1:
        s = nc.accumulate_along_rays(wc.detach(), None, ri.to(cuda), B)
        close(s, nv.accumulate_along_rays(w, None, ri, B), 1e-5, 1e-5)
    assert torch.all(out2[counts == 0] == 0)


def test_golden_dense_twin(cuda, golden):
    """Against the reference's own dense weights_from_sigma (eonerf.py:37-54) via tests/golden/volrend.npz."""
    from eonerf_code_b200 import nerfacc_compat as nc
    g = golden["volrend"]
    z, sig = t(g["z"]), t(g["sigma"])
    B, n = z.shape
    ts = z.flatten()
    te = torch.cat([z[:, 1:], torch.full((B, 1), 1e10)], 1).flatten()
    ri = torch.arange(B).repeat_interleave(n)
    w, T, a = nc.render_weight_from_density(ts.to(cuda), te.to(cuda), sig.flatten().to(cuda), ray_indices=ri.to(cuda), n_rays=B)
    close(w.view(B, n), t(g["weights"]), 1e-4, 1e-6)
    close(T.view(B, n), t(g["trans"]), 1e-4, 1e-6)
    pi = nc.pack_info(ri.to(cuda), B).cpu()
    assert pi[:, 1].tolist() == [n] * B and pi[:, 0].tolist() == [i * n for i in range(B)]
