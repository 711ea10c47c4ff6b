"""GPU: BASELINE configs[1] — VanillaNeRFRadianceField (mlp.py:211-250) against the reference class's own outputs
(tests/golden/vanilla.npz), the uniform in-box marcher and the nerfacc.rendering conventions against the oracle (that part of
the reference is missing upstream: product == restated oracle, parity unpinned)."""
import numpy as np
import pytest
import torch

from helpers import close, fingerprint, rel_err, t
from oracle import eonerf_oracle as O

pytestmark = pytest.mark.gpu
AABB = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]


def _model(p, cuda, precision):
    from eonerf_code_b200.radiance_fields import VanillaNeRFRadianceField
    m = VanillaNeRFRadianceField(precision=precision)
    missing, unexpected = m.load_state_dict(p, strict=False)
    assert not unexpected and all("scales" in k for k in missing)
    return m.to(cuda)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16_fused"])
def test_vanilla_field_golden(cuda, golden, precision):
    g = golden["vanilla"]
    p = O.init_vanilla_params(seed=int(g["seed"]), bias_scale=float(g["bias_scale"]))
    np.testing.assert_allclose(fingerprint(p), g["fingerprint"], rtol=1e-12)
    m = _model(p, cuda, precision)
    rgb, sigma = m(t(g["x"], cuda), t(g["viewdirs"], cuda))
    dens = m.query_density(t(g["x"], cuda))
    if precision == "fp32":
        close(rgb, t(g["rgb"]), 1e-5, 2e-6); close(sigma, t(g["sigma"]), 1e-5, 2e-6); close(dens, t(g["sigma"]), 1e-5, 2e-6)
    else:
        assert float((rgb.detach().cpu() - t(g["rgb"])).abs().max()) <= 4e-3
        assert float((sigma.detach().cpu() - t(g["sigma"])).abs().max()) <= 1e-2
    ((rgb * t(g["w_rgb"], cuda)).sum() + (sigma * t(g["w_sigma"], cuda)).sum()).backward()
    names = [str(k) for k in g["grad_names"]]
    assert [k for k, _ in m.named_parameters()] == names
    norms = np.array([float(v.grad.double().norm()) for _, v in m.named_parameters()])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-4 if precision == "fp32" else 0.1)
    if precision == "fp32":
        for k, v in m.named_parameters():
            if "grad__" + k in g.files:
                assert rel_err(v.grad, t(g["grad__" + k])) <= 2e-4, k


@pytest.mark.parametrize("jittered", [False, True])
def test_march_aabb_bit_exact_vs_oracle(cuda, jittered):
    from eonerf_code_b200 import ops
    from eonerf_code_b200.datasets.synthetic import make_pinhole_rays
    B, step = 1000, 5e-3
    o, d, _ = make_pinhole_rays(B, seed=3)
    d[5] = torch.tensor([0.0, 0.0, 1.0]); o[5] = torch.tensor([0.0, 0.0, 4.0])       # looks away from the box: no samples
    jit = torch.rand(B, generator=torch.Generator().manual_seed(1)) if jittered else None
    ri, ts, te, offs = ops.march_aabb(o.to(cuda), d.to(cuda), AABB, 0.0, 1e10, step, None if jit is None else jit.to(cuda))
    ri_o, ts_o, te_o = O.march_aabb(o, d, AABB, 0.0, 1e10, step, jit)
    assert ri.numel() == ri_o.numel() > 300 * B // 2
    assert torch.equal(ri.cpu(), ri_o) and torch.equal(ts.cpu(), ts_o) and torch.equal(te.cpu(), te_o)
    counts = (offs[1:] - offs[:-1]).cpu()
    assert int(counts[5]) == 0 and torch.equal(counts, torch.bincount(ri_o, minlength=B))
    x = o[ri_o] + d[ri_o] * ((ts_o + te_o) / 2)[:, None]
    assert float(x.abs().max()) <= 1.5 + 1e-3                                          # every sample inside the box


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16_fused"])
def test_vanilla_render_vs_oracle(cuda, precision):
    """render_image_with_occgrid (train_mlp_nerf.py:162-170 signature) forward + smooth-L1 backward vs the oracle."""
    from eonerf_code_b200.datasets.synthetic import make_pinhole_rays
    from eonerf_code_b200.nerfacc_compat import OccGridEstimator
    from eonerf_code_b200.vanilla_rendering import Rays, render_image_with_occgrid
    B, step = 96, 2e-2
    p = O.init_vanilla_params(seed=7, bias_scale=0.05)
    m = _model(p, cuda, precision).train()
    est = OccGridEstimator(roi_aabb=AABB, resolution=16, levels=1).to(cuda)
    o, d, px = make_pinhole_rays(B, seed=8)
    jit = torch.rand(B, generator=torch.Generator().manual_seed(2))
    bkgd = torch.ones(3)
    rgb, acc, depth, nren = render_image_with_occgrid(m, est, Rays(o.to(cuda), d.to(cuda)), near_plane=0.0, render_step_size=step,
                                                      render_bkgd=bkgd.to(cuda), jitter=jit.to(cuda))
    loss = torch.nn.functional.smooth_l1_loss(rgb, px.to(cuda))                       # train_mlp_nerf.py:183
    loss.backward()
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rgb_o, acc_o, depth_o, nren_o = O.vanilla_render(q, o, d, AABB, 0.0, 1e10, step, jit, bkgd)
    loss_o = torch.nn.functional.smooth_l1_loss(rgb_o, px)
    loss_o.backward()
    assert nren == nren_o and rgb.shape == (B, 3) and acc.shape == (B, 1) and depth.shape == (B, 1)
    if precision == "fp32":
        close(rgb, rgb_o, 1e-5, 2e-6); close(acc, acc_o, 1e-5, 2e-6); close(depth, depth_o, 1e-5, 2e-6); close(loss, loss_o, 1e-5)
        bad = {k: rel_err(v.grad, q[k].grad) for k, v in m.named_parameters() if rel_err(v.grad, q[k].grad) > 1e-3}
        assert not bad, bad
    else:
        assert float((rgb.detach().cpu() - rgb_o).abs().max()) <= 5e-3 and float((acc.detach().cpu() - acc_o).abs().max()) <= 5e-3
        assert float((depth.detach().cpu() - depth_o).abs().max()) <= 2e-2
        assert all(torch.isfinite(v.grad).all() for v in m.parameters())
    # eval mode: no stratified offset, chunked, same numbers as one chunk
    m.eval()
    with torch.no_grad():
        a1 = render_image_with_occgrid(m, est, Rays(o.to(cuda), d.to(cuda)), render_step_size=step, test_chunk_size=B)
        a2 = render_image_with_occgrid(m, est, Rays(o.to(cuda).view(4, B // 4, 3), d.to(cuda).view(4, B // 4, 3)), render_step_size=step, test_chunk_size=40)
    assert a2[0].shape == (4, B // 4, 3) and a1[3] == a2[3]
    close(a2[0].reshape(B, 3), a1[0], 1e-6, 1e-7)


def _l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.mark.parametrize("n,per_ray", [(300, False), (5000, False), (4099, True)])
def test_vanilla_fused_matches_layered(cuda, n, per_ray):
    """The fused tcgen05 program of the vanilla field (first ten stages of the EO-NeRF program, view-direction term as a per-row
    bias) vs the layer-by-layer bf16 kernels and the fp32 kernels: outputs, every parameter gradient incl. the view-direction
    columns 256:283 of rgb_layer.hidden_layers.0, with per-sample directions (forward(x, dirs)) and per-ray directions (the
    renderer's form)."""
    from eonerf_code_b200 import ops
    p = O.init_vanilla_params(seed=9, bias_scale=0.1)
    g = torch.Generator().manual_seed(n)
    res = {}
    if per_ray:
        B = 37
        o = (torch.rand(B, 3, generator=g) * 2 - 1).to(cuda)
        d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=1).to(cuda)
        ri = torch.sort(torch.randint(0, B, (n,), generator=g))[0].to(cuda)
        ts = (torch.rand(n, generator=g) * 0.5).to(cuda)
        te = ts + 0.01
    else:
        x = (torch.rand(n, 3, generator=g) * 3 - 1.5).to(cuda)
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1).to(cuda)
    g_rgb, g_sig = torch.randn(n, 3, generator=g).to(cuda), torch.randn(n, 1, generator=g).to(cuda)
    for prec in ("fp32", "bf16", "bf16_fused"):
        m = _model(p, cuda, prec)
        if per_ray:
            e = m._engine()
            sig, rgb, _ = ops._VanillaRaysFn.apply(True, e, o, d, ri, ts, te, None, *e.tensors())
        else:
            rgb, sig = m(x, d)
        ((rgb * g_rgb).sum() + (sig * g_sig).sum()).backward()
        res[prec] = (rgb.detach(), sig.detach(), {k: v.grad.clone() for k, v in m.named_parameters()})
    for a, b in ((res["bf16_fused"][0], res["bf16"][0]), (res["bf16_fused"][1], res["bf16"][1])):
        assert float((a - b).abs().max()) <= 3e-3 * max(1.0, float(b.abs().max()))
    assert float((res["bf16_fused"][0] - res["fp32"][0]).abs().max()) <= 4e-3
    for k in res["fp32"][2]:
        d_layer, d_fused = _l2(res["bf16"][2][k], res["fp32"][2][k]), _l2(res["bf16_fused"][2][k], res["fp32"][2][k])
        assert torch.isfinite(res["bf16_fused"][2][k]).all(), k
        assert d_fused <= 1.25 * d_layer + 1e-2, (k, d_fused, d_layer)
    # the view-direction columns specifically (they take a separate kernel in the fused mode)
    k = "mlp.rgb_layer.hidden_layers.0.weight"
    d_layer = _l2(res["bf16"][2][k][:, 256:], res["fp32"][2][k][:, 256:])
    d_fused = _l2(res["bf16_fused"][2][k][:, 256:], res["fp32"][2][k][:, 256:])
    assert d_fused <= 1.25 * d_layer + 1e-2 and d_fused <= 6e-2, (d_fused, d_layer)


@pytest.mark.parametrize("kind", ["vanilla_rays", "eonerf_field"])
def test_autograd_nodes_release_their_stash_without_the_cycle_collector(cuda, kind):
    """The per-sample stash (6 KB per sample) kept for the backward must die with the autograd graph, by reference counting
    alone: a forward dict holding the very tensors the Function returns closes a cycle (ctx -> dict -> output -> grad_fn ->
    ctx) and leaks one stash per step until Python's cycle collector happens to run."""
    import gc
    from eonerf_code_b200.datasets.synthetic import make_pinhole_rays
    from eonerf_code_b200.nerfacc_compat import OccGridEstimator
    from eonerf_code_b200.vanilla_rendering import Rays, render_image_with_occgrid
    if kind == "vanilla_rays":
        m = _model(O.init_vanilla_params(seed=7, bias_scale=0.05), cuda, "bf16_fused").train()
        est = OccGridEstimator(roi_aabb=AABB, resolution=16, levels=1).to(cuda)
        o, d, px = (v.to(cuda) for v in make_pinhole_rays(256, seed=8))

        def step():
            rgb = render_image_with_occgrid(m, est, Rays(o, d), near_plane=0.0, render_step_size=1e-2, render_bkgd=torch.ones(3, device=cuda))[0]
            torch.nn.functional.smooth_l1_loss(rgb, px).backward()
    else:
        from helpers import make_model
        m = make_model(O.init_params(5, seed=3, bias_scale=0.1), 5, cuda, "bf16_fused").train()
        x = (torch.rand(60000, 3, device=cuda) * 2 - 1)
        sun = torch.nn.functional.normalize(torch.rand(60000, 3, device=cuda), dim=-1)
        img = torch.randint(0, 5, (60000, 1), device=cuda)

        def step():
            outs = m(x, sun, img)
            sum(v.sum() for v in outs).backward()
    gc.collect()
    gc.disable()
    try:
        step()
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        grown = torch.cuda.memory_allocated() - base
    finally:
        gc.enable()
    assert grown < 32 * 2**20, f"{grown / 2**20:.0f} MiB still allocated after three more steps"


def test_vanilla_graph_step_matches_eager_step(cuda):
    """VanillaTrainStep: the sync-free step replayed as one CUDA graph (worst-case sized buffers, sample count on the device, flat
    Adam) against the eager step (exact buffers, host-read count) on the same rays and stratified offsets: sample counts equal,
    losses and parameters after three steps agree to the order-of-summation noise of the atomics."""
    from eonerf_code_b200.datasets.synthetic import make_pinhole_rays
    from eonerf_code_b200.nerfacc_compat import OccGridEstimator
    from eonerf_code_b200.training import VanillaTrainStep
    B, step = 512, 1e-2
    res = {}
    for mode in ("graph", "eager"):
        m = _model(O.init_vanilla_params(seed=7, bias_scale=0.05), cuda, "bf16_fused").train()
        est = OccGridEstimator(roi_aabb=AABB, resolution=16, levels=1).to(cuda)
        st = VanillaTrainStep(m, est, render_step_size=step, render_bkgd=torch.ones(3, device=cuda), graph=mode == "graph")
        losses, counts = [], []
        for i in range(3):
            o, d, px = (v.to(cuda) for v in make_pinhole_rays(B, seed=20 + i))
            jit = torch.rand(B, generator=torch.Generator().manual_seed(5 + i)).to(cuda)
            loss, n = st(o, d, px, jitter=jit)
            losses.append(float(loss))
            counts.append(int(n))
        res[mode] = (losses, counts, torch.cat([p.detach().flatten() for p in m.parameters()]).cpu())
    assert res["graph"][1] == res["eager"][1] and min(res["eager"][1]) > 0
    for a, b in zip(res["graph"][0], res["eager"][0]):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (res["graph"][0], res["eager"][0])
    assert float((res["graph"][2] - res["eager"][2]).abs().max()) <= 3e-3
    init = torch.cat([p.detach().flatten() for p in _model(O.init_vanilla_params(seed=7, bias_scale=0.05), cuda, "bf16_fused").parameters()]).cpu()
    assert float((res["graph"][2] - init).abs().max()) > 1e-4                      # the optimiser did step
