"""GPU: the fused tcgen05 MLP kernels (precision "bf16_fused", csrc/field_fused.cu) against the layer-by-layer bf16
tensor-core path (same rounding points: they may differ only by fp32 summation order and the ReLU / bf16 rounding flips
that follow from it), against the golden vectors of the reference, and the tile-blocked stash against the row-major one."""
import pytest
import torch

from helpers import make_model, t
from oracle import eonerf_oracle as O

pytestmark = pytest.mark.gpu

TILE, BLK = 128, 16384


def unblock(buf, off, n_tiles, nb):
    """tile-blocked bf16 array -> [n_tiles*128, nb*64] (undo the 128-byte swizzle)."""
    a = buf[off:off + n_tiles * nb * BLK].view(torch.bfloat16).view(n_tiles, nb, TILE, 8, 8)
    r = torch.arange(TILE, device=buf.device)[:, None]
    c = torch.arange(8, device=buf.device)[None, :]
    phys = (c ^ (r & 7)).view(1, 1, TILE, 8, 1).expand(n_tiles, nb, TILE, 8, 8)
    a = torch.gather(a, 3, phys)
    return a.permute(0, 2, 1, 3, 4).reshape(n_tiles * TILE, nb * 64)


def fused_stash_offsets(n, density_only):
    """Python twin of fused_stash_layout (csrc/field_fused.cuh)."""
    n_tiles = (n + TILE - 1) // TILE
    mpad = n_tiles * TILE
    off = 0

    def take(nbytes):
        nonlocal off
        o = off
        off = (off + nbytes + 1023) // 1024 * 1024
        return o
    L = dict(n_tiles=n_tiles, mpad=mpad, xf=take(mpad * 12), cls=take(mpad * 4), arr={}, mask={})
    for a in range(14):
        nb = 1 if a == 13 else (2 if 10 <= a < 13 else 4)
        if a < 8 or a == 13 or not density_only:
            L["arr"][a] = (take(n_tiles * nb * BLK), nb)
    for m in range(12):
        if m < 8 or not density_only:
            L["mask"][m] = take(mpad * 32)
    L["total"] = off
    return L


def run(model, n, density_only, x, img, keep=True):
    e = model._engine()
    return e.fwd(n, density_only, x=x, img_idx=None if density_only else img, keep=keep)


@pytest.mark.parametrize("n", [1, 127, 128, 300, 1000, 4099])
@pytest.mark.parametrize("density_only", [False, True])
def test_fused_forward_matches_layered(cuda, n, density_only):
    n_img = 7
    p = O.init_params(n_img, seed=3, bias_scale=0.1)
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).to(cuda)
    img = torch.randint(0, n_img, (n, 1), generator=g).to(cuda)
    ml, mf = make_model(p, n_img, cuda, "bf16"), make_model(p, n_img, cuda, "bf16_fused")
    a = run(ml, n, density_only, x, img)
    for keep in (True, False):
        b = run(mf, n, density_only, x, img, keep=keep)
        torch.cuda.synchronize()
        for k in ("sigma",) if density_only else ("sigma", "rgb", "transient_s", "transient_beta"):
            err = (a[k].double() - b[k].double()).abs()
            assert torch.isfinite(b[k]).all(), k
            assert float(err.max()) <= 2e-3 and float(err.mean()) <= 5e-5, (k, keep, float(err.max()), float(err.mean()))


def test_tmem_operand_inference_kernel_matches_layered(cuda):
    """The opt-in TMEM-operand inference kernel (csrc/field_fused_ts.cu, EONERF_FUSED_TS=1; the switch is read once per process):
    the forward parity cases above, re-run in a child process with the kernel enabled (keep=False rows go through it)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, EONERF_FUSED_TS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-k", "test_fused_forward_matches_layered or test_fused_forward_golden",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_fused_stash_is_the_blocked_image_of_the_layered_stash(cuda):
    n, n_img = 1000, 5
    p = O.init_params(n_img, seed=4, bias_scale=0.1)
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).to(cuda)
    img = torch.randint(0, n_img, (n, 1), generator=g).to(cuda)
    ml, mf = make_model(p, n_img, cuda, "bf16"), make_model(p, n_img, cuda, "bf16_fused")
    a, b = run(ml, n, False, x, img), run(mf, n, False, x, img)
    torch.cuda.synchronize()
    L = fused_stash_offsets(n, False)
    from eonerf_code_b200 import _capi as K
    assert L["total"] == K.lib().eonerf_field_stash_bytes(K.FIELD_EONERF, K.PREC_BF16_FUSED, n, 0)
    st = b["stash"]
    # layered stash: xf [n,3] fp32 first (csrc/field_layout.cuh), then cls, then H0..H3 [n,256], H4E [n,320], ...
    xf = st[L["xf"]:L["xf"] + n * 12].view(torch.float32).view(n, 3)
    assert torch.equal(xf, x)
    cls = st[L["cls"]:L["cls"] + n * 4].view(torch.int32)
    assert torch.equal(cls.long(), img[:, 0])
    lay = a["stash"]

    def layered(off_elems_bytes, ld, cols):
        return lay[off_elems_bytes:off_elems_bytes + n * ld * 2].view(torch.bfloat16).view(n, ld)[:, :cols]
    al = lambda v: (v + 255) // 256 * 256
    off = al(n * 12)
    off = off + al(n * 4)
    lay_h = []
    for i in range(8):
        ld = 320 if i == 4 else 256
        lay_h.append((off, ld))
        off += al(n * ld * 2)
    enc_f = unblock(st, L["arr"][13][0], L["n_tiles"], 1)[:n].float()
    enc_l = layered(lay_h[4][0], 320, 320)[:, 256:320].float()
    # the fused kernel evaluates the sines with a Cody-Waite reduction + MUFU.SIN (abs error < 6e-7) instead of sinf():
    # after rounding to bf16 a value may land on the neighbouring bf16 number, rarely
    d = (enc_f - enc_l).abs()
    assert float(d.max()) <= 2 ** -8 and float((d > 0).float().mean()) < 2e-3
    h0_f = unblock(st, L["arr"][0][0], L["n_tiles"], 4)[:n].float()
    h0_l = layered(lay_h[0][0], 256, 256).float()
    d = (h0_f - h0_l).abs()
    assert float(d.max()) <= 0.02 * float(h0_l.abs().max()) and float((d > 0).float().mean()) < 0.05
    h7_f = unblock(st, L["arr"][7][0], L["n_tiles"], 4)[:n].float()
    h7_l = layered(lay_h[7][0], 256, 256).float()
    assert float((h7_f - h7_l).abs().max()) <= 0.05 * float(h7_l.abs().max())
    # ReLU sign bits agree with the stashed activations they were taken from
    for stage, arr in ((0, h0_f), (7, h7_f)):
        m = st[L["mask"][stage]:L["mask"][stage] + n * 32].view(torch.int32).view(n, 8)
        j = torch.arange(16, device=cuda)
        pos = torch.stack([15 - j, 31 - j], 1).reshape(32)            # column 2j -> bit 15-j, column 2j+1 -> bit 31-j
        bits = ((m[:, :, None] >> pos[None, None, :]) & 1).reshape(n, 256).bool()
        assert torch.equal(bits, arr > 0)


@pytest.mark.parametrize("precision", ["bf16_fused"])
def test_fused_forward_golden(cuda, golden, precision):
    from test_gpu_field import NAMES, check_outputs
    g = golden["field"]
    n_img = int(g["n_img"])
    p = O.init_params(n_img, seed=int(g["seed"]), bias_scale=float(g["bias_scale"]))
    m = make_model(p, n_img, cuda, precision)
    x, sun, img = t(g["x"], cuda), t(g["sun"], cuda), t(g["img"], cuda)
    with torch.no_grad():
        outs = m(x, sun, img)
        dens = m.query_density(x)
    check_outputs(outs, [t(g[k]) for k in NAMES], "bf16")
    check_outputs([dens], [t(g["density"])], "bf16")


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.mark.parametrize("n", [100, 1000, 4099])
@pytest.mark.parametrize("density_only", [False, True])
def test_fused_backward_matches_layered(cuda, n, density_only):
    """Parameter gradients and dL/dx of the fused chain vs the layer-by-layer bf16 kernels on the same inputs and the same
    upstream gradients (both round G to bf16 at every layer; they differ by summation order and rounding flips)."""
    n_img = 6
    p = O.init_params(n_img, seed=5, bias_scale=0.1)
    g = torch.Generator().manual_seed(100 + n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).to(cuda)
    img = torch.sort(torch.randint(0, n_img, (n,), generator=g))[0][:, None].to(cuda)
    gs, g3 = torch.randn(n, generator=g).to(cuda), torch.randn(n, 3, generator=g).to(cuda)
    gts, gtb = torch.randn(n, generator=g).to(cuda), torch.randn(n, generator=g).to(cuda)
    res = {}
    for mode in ("fp32", "bf16", "bf16_fused"):
        m = make_model(p, n_img, cuda, mode)
        e = m._engine()
        f = e.fwd(n, density_only, x=x, img_idx=None if density_only else img)
        flat, views, gstruct = e.new_grads()
        gx = e.bwd(n, density_only, f, g_sigma=gs, g_rgb=None if density_only else g3, g_ts=None if density_only else gts,
                   g_tb=None if density_only else gtb, grads_struct=gstruct, want_gx=density_only)
        torch.cuda.synchronize()
        res[mode] = (views, gx)
    v32, va, vb = res["fp32"][0], res["bf16"][0], res["bf16_fused"][0]
    for k in va:
        if float(va[k].abs().max()) == 0.0:
            assert float(vb[k].abs().max()) == 0.0, k
            continue
        assert torch.isfinite(vb[k]).all(), k
        # With these random upstream gradients both bf16 paths sit ~1e-1 (relative L2) from the fp32 kernels
        # (tools/diag_grad_noise.py); their mutual distance is de-correlated rounding, a fraction of that.  The bar: the fused
        # chain is as close to fp32 as the layer-by-layer chain is, and the two agree to within that noise.
        d_layer, d_fused = l2(va[k], v32[k]), l2(vb[k], v32[k])
        assert d_fused <= 1.25 * d_layer + 5e-3, (k, d_fused, d_layer)
        assert l2(vb[k], va[k]) <= max(5e-2, 0.6 * d_layer), (k, l2(vb[k], va[k]), d_layer)
    if density_only:
        assert l2(res["bf16_fused"][1], res["bf16"][1]) <= 2e-2
