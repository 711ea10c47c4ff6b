"""CPU: pins oracle/ (the restatement) against tests/golden/*.npz, which were produced by running the
unmodified reference Python from /root/reference (oracle/make_golden.py, committed)."""
import numpy as np
import torch

from helpers import fingerprint, t
from oracle import eonerf_oracle as O
from oracle import nerfacc_v052 as nv


def test_sampling_bit_exact(golden):
    g = golden["sampling"]
    for tag in ("n64", "n96", "n128_inside"):
        rays, u, n = t(g[f"{tag}_rays"]), t(g[f"{tag}_u"]), int(g[f"{tag}_n"])
        assert n == O.n_samples_from_step(float(g[f"{tag}_step"]))
        assert torch.equal(torch.linspace(0, 1, n), t(g[f"{tag}_z_steps"]))
        ri, t0, t1, _ = O.satnerf_sampling(rays[:, 0:3], rays[:, 3:6], n, u, near=rays[:, 6:7])
        assert torch.equal(ri, t(g[f"{tag}_ray_indices"]))
        assert torch.equal(t0, t(g[f"{tag}_t_starts"])) and torch.equal(t1, t(g[f"{tag}_t_ends"]))
        assert torch.equal(O.pts_per_ray(ri, rays.shape[0]), t(g[f"{tag}_pts_per_ray"]))
    assert int(golden["sampling"]["n96_n"]) == 95          # fp32 quotient quirk: int(2/fl32(2/96))
    assert (golden["sampling"]["n64_pts_per_ray"] == 0).sum() >= 1   # the fixture holds an empty ray


def test_field_forward_and_density_gradient(golden):
    g = golden["field"]
    p = O.init_params(int(g["n_img"]), seed=int(g["seed"]), bias_scale=float(g["bias_scale"]))
    np.testing.assert_allclose(fingerprint(p), g["fingerprint"], rtol=1e-12)
    x, sun, img = t(g["x"]), t(g["sun"]), t(g["img"])
    with torch.no_grad():
        out = O.field_forward(p, x, sun, img)
    for a, k in zip(out, ("sigma", "albedo", "ambient", "transient_s", "transient_beta")):
        assert torch.allclose(a, t(g[k]), rtol=1e-6, atol=1e-6), k
    xg = x.clone().requires_grad_(True)
    d = O.query_density(p, xg)
    assert torch.allclose(d.detach(), t(g["density"]), rtol=1e-6, atol=1e-6)
    d.sum().backward()
    assert torch.allclose(xg.grad, t(g["d_density_dx"]), rtol=1e-4, atol=1e-5)


def test_volrend_vs_reference_dense_twin(golden):
    """nerfacc-form weights == the reference's own dense `weights_from_sigma` (eonerf.py:37-54)."""
    g = golden["volrend"]
    z, sig = t(g["z"]), t(g["sigma"])
    B, n = z.shape
    # dense twin: deltas = z[1:]-z[:-1], last = 1e10
    ts = z.flatten()
    te = torch.cat([z[:, 1:], torch.full((B, 1), 1e10)], 1).flatten()
    ri = torch.arange(B).repeat_interleave(n)
    w, T, a = nv.render_weight_from_density(ts, te, sig.flatten(), ray_indices=ri, n_rays=B)
    assert torch.allclose(w.view(B, n), t(g["weights"]), rtol=1e-4, atol=1e-6)
    assert torch.allclose(T.view(B, n), t(g["trans"]), rtol=1e-4, atol=1e-6)
    assert torch.allclose(a.view(B, n), t(g["alphas"]), rtol=1e-5, atol=1e-6)
    s = nv.accumulate_along_rays(w, None, ri, B)
    assert torch.allclose(s, torch.ones_like(s), atol=1e-5)   # sum w = 1 when the last interval is 1e10


def test_exclusive_sum_backward_matches_autograd():
    g = torch.Generator().manual_seed(0)
    ri = torch.sort(torch.randint(0, 7, (50,), generator=g)).values
    x = torch.rand(50, generator=g, dtype=torch.float64, requires_grad=True)
    y = nv.exclusive_sum(x, ri, 7)
    ref = torch.stack([x[(ri == ri[i]) & (torch.arange(50) < i)].sum() for i in range(50)])
    assert torch.allclose(y, ref)
    gy = torch.rand(50, generator=g, dtype=torch.float64)
    (gx,) = torch.autograd.grad(y, x, gy)
    (gref,) = torch.autograd.grad(ref, x, gy)
    assert torch.allclose(gx, gref)


def test_render_chunk_against_reference_outputs(golden):
    g = golden["render"]
    for tag in ("train_e2", "train_e0", "eval_e5"):
        n_img, n, epoch, ev = int(g[f"{tag}_n_img"]), int(g[f"{tag}_n"]), int(g[f"{tag}_epoch"]), bool(g[f"{tag}_eval"])
        p = O.init_params(n_img, seed=21, bias_scale=0.05)
        np.testing.assert_allclose(fingerprint(p), g[f"{tag}_fingerprint"], rtol=1e-12)
        rays, ts, pixels = t(g[f"{tag}_rays"]), t(g[f"{tag}_ts"]), t(g[f"{tag}_pixels"])
        sr = O.satrays_from_table(rays, ts)
        u_cam, u_sun = t(g[f"{tag}_u_cam"]), t(g[f"{tag}_u_sun"])
        if ev:
            with torch.no_grad():
                out, nren = O.render_chunk(p, sr, n, epoch, u_cam, u_sun, eval=True)
        else:
            loss, out, grads, nren = O.train_step_grads(p, sr, pixels, n, epoch, u_cam, u_sun)
            assert torch.allclose(loss, t(g[f"{tag}_loss"]), rtol=1e-5)
            norms = np.array([float(v.double().norm()) for v in grads.values()])
            np.testing.assert_allclose(norms, g[f"{tag}_grad_norms"], rtol=2e-4, atol=1e-9)
            heads = np.stack([np.resize(v.flatten()[:16].numpy(), 16) for v in grads.values()])
            np.testing.assert_allclose(heads, g[f"{tag}_grad_heads"], rtol=2e-3, atol=1e-7)
            assert list(grads.keys()) == [str(s) for s in g[f"{tag}_grad_names"]]
        assert nren == int(g[f"{tag}_n_rendering_samples"])
        assert torch.allclose(out, t(g[f"{tag}_out"]), rtol=1e-5, atol=1e-6), tag


def test_render_depth_only(golden):
    g = golden["render"]
    p = O.init_params(6, seed=21, bias_scale=0.05)
    rays, ts, u, n = t(g["depth_rays"]), t(g["depth_ts"]), t(g["depth_u"]), int(g["depth_n"])
    sr = O.satrays_from_table(rays, ts)
    ri, t0, t1, _ = O.satnerf_sampling(sr.origins, sr.viewdirs, n, u, near=sr.t_near)
    with torch.no_grad():
        d = O.render_depth(p, sr, t0, t1, ri)
    assert ri.numel() == int(g["depth_n_rendering_samples"])
    assert torch.allclose(d, t(g["depth_out"]), rtol=1e-5, atol=1e-6)


def test_analytic_identities():
    """rgb == clip(albedo) when epoch<2 with identity radiometric; entropy / opacity columns == 1."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    p = O.init_params(4, seed=3)
    rays, ts, _ = make_rays(16, 4, seed=9)
    u = torch.rand(16, 32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        out, _ = O.render_chunk(p, O.satrays_from_table(rays, ts), 32, 0, u)
    assert torch.equal(out[:, 0:3], torch.clip(out[:, 4:7], 0, 1))
    assert torch.all(out[:, 13] == 1) and torch.all(out[:, 15:18] == 1) and torch.all(out[:, 10] == 1)


def test_redraw_branch_against_reference_outputs(golden):
    """sat_rendering.py:259-262 (zero-sample re-draw) pinned by the reference's own outputs, tests/golden/redraw.npz."""
    g = golden["redraw"]
    n_img, n, epoch = int(g["n_img"]), int(g["n"]), int(g["epoch"])
    p = O.init_params(n_img, seed=21, bias_scale=0.05)
    np.testing.assert_allclose(fingerprint(p), g["fingerprint"], rtol=1e-12)
    rays, ts, pixels = t(g["rays"]), t(g["ts"]), t(g["pixels"])
    sr = O.satrays_from_table(rays, ts)
    assert O.chunk_needs_redraw(sr, n, t(g["u_cam"]))
    loss, out, grads, nren = O.train_step_grads(p, sr, pixels, n, epoch, t(g["u_cam"]), t(g["u_sun"]), t(g["u_cam2"]))
    assert nren == int(g["n_rendering_samples"])
    assert torch.allclose(out, t(g["out"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(loss, t(g["loss"]), rtol=1e-5)
    assert int((out[:, 14] == 0).sum()) == 4            # pts_per_ray keeps the FIRST draw's counts
    norms = np.array([float(grads[str(k)].double().norm()) for k in g["grad_names"]])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-3, atol=1e-9)


def test_vanilla_field_against_reference_outputs(golden):
    """VanillaNeRFRadianceField (mlp.py:211-250) pinned by the reference class's own outputs and gradients."""
    g = golden["vanilla"]
    p = O.init_vanilla_params(seed=int(g["seed"]), bias_scale=float(g["bias_scale"]))
    np.testing.assert_allclose(fingerprint(p), g["fingerprint"], rtol=1e-12)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rgb, sigma = O.vanilla_forward(q, t(g["x"]), t(g["viewdirs"]))
    assert torch.allclose(rgb, t(g["rgb"]), rtol=1e-6, atol=1e-6) and torch.allclose(sigma, t(g["sigma"]), rtol=1e-6, atol=1e-6)
    ((rgb * t(g["w_rgb"])).sum() + (sigma * t(g["w_sigma"])).sum()).backward()
    names = [str(k) for k in g["grad_names"]]
    assert names == list(p.keys())
    np.testing.assert_allclose(np.array([float(q[k].grad.double().norm()) for k in names]), g["grad_norms"], rtol=1e-4)
    for k in names:
        if "grad__" + k in g.files:
            assert torch.allclose(q[k].grad, t(g["grad__" + k]), rtol=1e-4, atol=1e-6), k


def test_sliced_oracle_equals_whole_chunk():
    """render_chunk_sliced / train_step_grads_sliced (what the at-size GPU tests use) == the one-pass forms, including the
    chunk-wide re-draw decision and the eval image index."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    B, n, n_img = 96, 24, 4
    p = O.init_params(n_img, seed=3, bias_scale=0.05)
    rays, ts, pixels = make_rays(B, n_img, seed=4)
    rays[70, 6] = 3.0                                    # one empty ray in the LAST slice -> every slice must re-draw
    g = torch.Generator().manual_seed(5)
    u = [torch.rand(B, n, generator=g) for _ in range(3)]
    sr = O.satrays_from_table(rays, ts)
    for ev in (False, True):
        with torch.no_grad():
            a, na = O.render_chunk(p, sr, n, 2, u[0], u[1], u[2], eval=ev)
            b, nb = O.render_chunk_sliced(p, sr, n, 2, u[0], u[1], u[2], eval=ev, rays_per_slice=32)
        assert na == nb and torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    for emu in (False, True):
        la, _, ga, _ = O.train_step_grads(p, sr, pixels, n, 2, u[0], u[1], u[2], emulate_bf16=emu)
        lb, _, gb, _ = O.train_step_grads_sliced(p, sr, pixels, n, 2, u[0], u[1], u[2], emulate_bf16=emu, rays_per_slice=32)
        assert abs(float(la) - float(lb)) < 1e-5
        for k in ga:
            # bf16 emulation rounds the back-propagated gradients to bf16 at every layer: a slice weight that is not a power of
            # two (32/96) moves the rounding points, so the two forms agree to bf16 round-off only (2^-9 per element)
            d = float((ga[k] - gb[k]).norm() / max(float(ga[k].norm()), 1e-30))
            assert d <= (1e-2 if emu else 1e-4), (k, d)
