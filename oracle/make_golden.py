"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the unmodified reference Python.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The fixtures are committed; the GPU box never runs this script.  While generating, every fixture is
also checked against oracle/eonerf_oracle.py so a drift between reference and restatement fails here.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import eonerf_oracle as O          # noqa: E402
from oracle import ref_harness                 # noqa: E402
from eonerf_code_b200.datasets.synthetic import make_rays  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def param_fingerprint(p):
    return np.array([float(v.double().sum()) for v in p.values()] +
                    [float(v.double().abs().sum()) for v in p.values()])


def ref_model(ref, p, n_img):
    m = ref.eonerf.EONerfMLP(n_img, radiometric_normalization=True)
    missing, unexpected = m.load_state_dict(p, strict=False)
    assert not unexpected and all("scales" in k for k in missing), (missing, unexpected)
    return m


def gen_sampling(ref):
    out = {}
    for tag, (B, n_arg, variant, seed) in {"n64": (48, 64, "spread", 1), "n96": (40, 96, "spread", 2),
                                          "n128_inside": (24, 128, "inside", 3)}.items():
        rays, ts, _ = make_rays(B, 7, seed=seed, variant=variant)
        if tag == "n64":
            rays[5, 0:3] = torch.tensor([3.0, 3.0, 1.0])     # a ray that never enters the cube -> 0 samples
            rays[6, 6] = 0.25                                 # non-zero t_near
        step = (torch.tensor(2.0) / n_arg).item()             # fp32 quotient, train_eonerf.py:50-53
        n = int(2 / step)
        g = torch.Generator().manual_seed(100 + seed)
        u = torch.rand(B, n, generator=g)
        sr = ref.satellite.define_satrays_from_tensors(rays, ts)
        with ref_harness.FixedRand([u]):
            ri, t0, t1 = ref.sat_rendering.satnerf_sampling(sr.origins, sr.viewdirs, {"render_step_size": step},
                                                            near=sr.t_near)
        ppr = ref.sat_rendering.count_number_of_pts_per_nerfacc_ray(sr, ri)
        ori, ot0, ot1, _ = O.satnerf_sampling(sr.origins, sr.viewdirs, n, u, near=sr.t_near)
        assert torch.equal(ri, ori) and torch.equal(t0, ot0) and torch.equal(t1, ot1), tag
        assert torch.equal(ppr, O.pts_per_ray(ori, B))
        out.update({f"{tag}_rays": rays.numpy(), f"{tag}_u": u.numpy(), f"{tag}_step": np.float64(step),
                    f"{tag}_n": np.int64(n), f"{tag}_ray_indices": ri.numpy(), f"{tag}_t_starts": t0.numpy(),
                    f"{tag}_t_ends": t1.numpy(), f"{tag}_pts_per_ray": ppr.numpy(),
                    f"{tag}_z_steps": torch.linspace(0, 1, n).numpy()})
        print(f"sampling {tag}: n={n} kept {ri.numel()} of {B * (n - 1)}; empty rays {(ppr == 0).sum().item()}")
    np.savez_compressed(os.path.join(GOLD, "sampling.npz"), **out)


def gen_field(ref):
    n_img, N = 5, 300
    p = O.init_params(n_img, seed=11, bias_scale=0.1)
    m = ref_model(ref, p, n_img)
    g = torch.Generator().manual_seed(12)
    x = torch.rand(N, 3, generator=g) * 2 - 1
    sun = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1)
    img = torch.randint(0, n_img, (N, 1), generator=g)
    with torch.no_grad():
        ref_out = m(x, sun, img)
        ora_out = O.field_forward(p, x, sun, img)
        dens = m.query_density(x)
    for a, b in zip(ref_out, ora_out):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    # input gradient of the density (needed by the shadow pass, sat_rendering.py:90)
    xg = x.clone().requires_grad_(True)
    m.query_density(xg).sum().backward()
    names = ["sigma", "albedo", "ambient", "transient_s", "transient_beta"]
    np.savez_compressed(os.path.join(GOLD, "field.npz"), n_img=n_img, seed=11, bias_scale=0.1,
                        fingerprint=param_fingerprint(p), x=x.numpy(), sun=sun.numpy(), img=img.numpy(),
                        density=dens.numpy(), d_density_dx=xg.grad.numpy(),
                        **{k: v.numpy() for k, v in zip(names, ref_out)})
    print("field: ok", [tuple(o.shape) for o in ref_out])


def gen_render(ref):
    fixtures = {}
    for tag, (B, n_arg, epoch, ev, variant) in {"train_e2": (64, 32, 2, False, "spread"),
                                                "train_e0": (32, 32, 0, False, "spread"),
                                                "eval_e5": (32, 48, 5, True, "inside")}.items():
        n_img = 6
        p = O.init_params(n_img, seed=21, bias_scale=0.05)
        m = ref_model(ref, p, n_img)
        rays, ts, pixels = make_rays(B, n_img, seed=5, variant=variant)
        step = (torch.tensor(2.0) / n_arg).item()
        n = int(2 / step)
        g = torch.Generator().manual_seed(31)
        u_cam, u_sun = torch.rand(B, n, generator=g), torch.rand(B, n, generator=g)
        sr = ref.satellite.define_satrays_from_tensors(rays, ts)
        m.train(not ev)
        with ref_harness.FixedRand([u_cam, u_sun]):
            res, nren = ref.sat_rendering.render_image(m, None, sr, None, None, epoch_idx=epoch, chunk=B,
                                                       render_step_size=step, eval=ev)
        out_ref = torch.cat([res[k] for k in O.OUT_KEYS], 1)
        if epoch < 2:
            loss = torch.nn.functional.mse_loss(res["rgb"], pixels)
        else:
            loss, _ = ref.metrics.uncertainty_aware_loss(pixels, res["rgb"], res["beta"])
        m.zero_grad()
        loss.backward()
        gref = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in m.named_parameters()}

        osr = O.satrays_from_table(rays, ts)
        if ev:
            with torch.no_grad():
                out_o, nren_o = O.render_chunk(p, osr, n, epoch, u_cam, u_sun, eval=True)
            gora, loss_o = None, O.loss_from_out(out_o, pixels, epoch)
        else:
            loss_o, out_o, gora, nren_o = O.train_step_grads(p, osr, pixels, n, epoch, u_cam, u_sun)
        assert nren == nren_o
        assert torch.allclose(out_ref.detach(), out_o, rtol=1e-5, atol=1e-6), (out_ref.detach() - out_o).abs().max()
        assert torch.allclose(loss.detach(), loss_o, rtol=1e-5)
        fx = {f"{tag}_rays": rays.numpy(), f"{tag}_ts": ts.numpy(), f"{tag}_pixels": pixels.numpy(),
              f"{tag}_u_cam": u_cam.numpy(), f"{tag}_u_sun": u_sun.numpy(), f"{tag}_n": np.int64(n),
              f"{tag}_epoch": np.int64(epoch), f"{tag}_eval": np.int64(ev), f"{tag}_n_img": np.int64(n_img),
              f"{tag}_out": out_ref.detach().numpy(), f"{tag}_loss": loss.detach().numpy(),
              f"{tag}_n_rendering_samples": np.int64(nren), f"{tag}_fingerprint": param_fingerprint(p)}
        if gora is not None:
            for k in gref:
                a, b = gref[k], gora[k]
                assert torch.allclose(a, b, rtol=2e-4, atol=1e-7), (k, (a - b).abs().max(), a.abs().max())
            fx[f"{tag}_grad_norms"] = np.array([float(gref[k].double().norm()) for k in gref])
            fx[f"{tag}_grad_heads"] = np.stack([np.resize(gref[k].flatten()[:16].numpy(), 16) for k in gref])
            fx[f"{tag}_grad_names"] = np.array(list(gref.keys()))
        fixtures.update(fx)
        print(f"render {tag}: n={n} n_rendering_samples={nren} loss={float(loss):.6f}")

    # only_depth branch (sat_rendering.py:227-249)
    B, n_img, n = 32, 6, 32
    p = O.init_params(n_img, seed=21, bias_scale=0.05)
    m = ref_model(ref, p, n_img)
    rays, ts, _ = make_rays(B, n_img, seed=6, variant="spread")
    u = torch.rand(B, n, generator=torch.Generator().manual_seed(32))
    sr = ref.satellite.define_satrays_from_tensors(rays, ts)
    with torch.no_grad(), ref_harness.FixedRand([u]):
        res, nren = ref.sat_rendering.render_image(m, None, sr, None, None, epoch_idx=3, chunk=B,
                                                   render_step_size=2.0 / n, only_depth=True)
    osr = O.satrays_from_table(rays, ts)
    ri, t0, t1, _ = O.satnerf_sampling(osr.origins, osr.viewdirs, n, u, near=osr.t_near)
    with torch.no_grad():
        d_o = O.render_depth(p, osr, t0, t1, ri)
    assert torch.allclose(res["depth"], d_o, rtol=1e-5, atol=1e-6)
    fixtures.update({"depth_rays": rays.numpy(), "depth_ts": ts.numpy(), "depth_u": u.numpy(),
                     "depth_n": np.int64(n), "depth_out": res["depth"].numpy(),
                     "depth_n_rendering_samples": np.int64(nren)})
    np.savez_compressed(os.path.join(GOLD, "render.npz"), **fixtures)


def gen_volrend(ref):
    """nerfacc-form weights vs the reference's own dense dead-code twin weights_from_sigma (eonerf.py:37-54)."""
    g = torch.Generator().manual_seed(41)
    B, n = 16, 24
    z = torch.sort(torch.rand(B, n, generator=g) * 2, dim=1).values
    sig = torch.rand(B, n, generator=g) * 30
    w_dense, t_dense, a_dense = ref.eonerf.weights_from_sigma(z, sig)
    np.savez_compressed(os.path.join(GOLD, "volrend.npz"), z=z.numpy(), sigma=sig.numpy(),
                        weights=w_dense.numpy(), trans=t_dense.numpy(), alphas=a_dense.numpy())
    print("volrend: ok")


def gen_redraw(ref):
    """The zero-sample re-draw branch (sat_rendering.py:259-262): three rays start at t_near=3 (every first-draw sample is
    below the cube) so the whole chunk is sampled again with near=None and a second set of uniforms; one ray never enters
    the cube at all (still empty after the re-draw: zero weights, depth 0, beta = beta_min, geo_shadow = 1).  pts_per_ray
    keeps the FIRST draw's counts (:258, :308)."""
    B, n_arg, epoch, n_img = 48, 32, 2, 6
    p = O.init_params(n_img, seed=21, bias_scale=0.05)
    m = ref_model(ref, p, n_img)
    rays, ts, pixels = make_rays(B, n_img, seed=9, variant="spread")
    rays[[3, 17, 40], 6] = 3.0
    rays[11, 0:3] = torch.tensor([3.0, 3.0, 1.0])
    step = (torch.tensor(2.0) / n_arg).item()
    n = int(2 / step)
    g = torch.Generator().manual_seed(77)
    u_cam, u_cam2, u_sun = (torch.rand(B, n, generator=g) for _ in range(3))
    sr = ref.satellite.define_satrays_from_tensors(rays, ts)
    m.train()
    with ref_harness.FixedRand([u_cam, u_cam2, u_sun]) as fr:
        res, nren = ref.sat_rendering.render_image(m, None, sr, None, None, epoch_idx=epoch, chunk=B, render_step_size=step)
        assert not fr.queue, "the reference did not take the re-draw branch"
    out_ref = torch.cat([res[k] for k in O.OUT_KEYS], 1)
    loss, _ = ref.metrics.uncertainty_aware_loss(pixels, res["rgb"], res["beta"])
    m.zero_grad()
    loss.backward()
    gref = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in m.named_parameters()}
    loss_o, out_o, gora, nren_o = O.train_step_grads(p, O.satrays_from_table(rays, ts), pixels, n, epoch, u_cam, u_sun, u_cam2)
    assert nren == nren_o
    assert (out_ref[:, 14] == 0).sum() == 4 and float(out_ref[11, 3]) == 0.0
    assert torch.allclose(out_ref.detach(), out_o, rtol=1e-5, atol=1e-6), (out_ref.detach() - out_o).abs().max()
    assert torch.allclose(loss.detach(), loss_o, rtol=1e-5)
    for k in gref:
        assert torch.allclose(gref[k], gora[k], rtol=2e-4, atol=1e-7), k
    np.savez_compressed(os.path.join(GOLD, "redraw.npz"), rays=rays.numpy(), ts=ts.numpy(), pixels=pixels.numpy(),
                        u_cam=u_cam.numpy(), u_cam2=u_cam2.numpy(), u_sun=u_sun.numpy(), n=np.int64(n), epoch=np.int64(epoch),
                        n_img=np.int64(n_img), out=out_ref.detach().numpy(), loss=loss.detach().numpy(),
                        n_rendering_samples=np.int64(nren), fingerprint=param_fingerprint(p),
                        grad_norms=np.array([float(gref[k].double().norm()) for k in gref]),
                        grad_names=np.array(list(gref.keys())))
    print(f"redraw: n={n} n_rendering_samples={nren} loss={float(loss):.6f} empty first-draw rays=4")


def gen_vanilla(ref):
    """VanillaNeRFRadianceField (mlp.py:211-250): forward(x, viewdirs) -> (rgb, sigma), query_density, and the parameter
    gradients of a scalar loss, from the reference class itself."""
    N = 300
    p = O.init_vanilla_params(seed=51, bias_scale=0.1)
    m = ref.mlp.VanillaNeRFRadianceField()
    missing, unexpected = m.load_state_dict(p, strict=False)
    assert not unexpected and all("scales" in k for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(52)
    x = torch.rand(N, 3, generator=g) * 3 - 1.5
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1)
    rgb, sigma = m(x, d)
    dens = m.query_density(x)
    with torch.no_grad():
        rgb_o, sigma_o = O.vanilla_forward(p, x, d)
    assert torch.allclose(rgb.detach(), rgb_o, rtol=1e-6, atol=1e-6) and torch.allclose(sigma.detach(), sigma_o, rtol=1e-6, atol=1e-6)
    assert torch.equal(dens, sigma)
    w_rgb, w_sig = torch.rand(N, 3, generator=g), torch.rand(N, 1, generator=g)
    ((rgb * w_rgb).sum() + (sigma * w_sig).sum()).backward()
    gref = {k: v.grad for k, v in m.named_parameters()}
    np.savez_compressed(os.path.join(GOLD, "vanilla.npz"), seed=51, bias_scale=0.1, fingerprint=param_fingerprint(p), x=x.numpy(),
                        viewdirs=d.numpy(), rgb=rgb.detach().numpy(), sigma=sigma.detach().numpy(), w_rgb=w_rgb.numpy(),
                        w_sigma=w_sig.numpy(), grad_names=np.array(list(gref.keys())),
                        grad_norms=np.array([float(v.double().norm()) for v in gref.values()]),
                        **{"grad__" + k: v.numpy() for k, v in gref.items() if v.numel() <= 768})
    print("vanilla: ok", tuple(rgb.shape), tuple(sigma.shape), f"sigma>0: {(sigma > 0).float().mean():.2f}")


def gen_nadir(ref):
    """create_rays_from_nadir / generate_rays_from_virtual_pinhole (eval_eonerf.py:78-249) and the UTM/altitude point cloud of
    get_utmalt_from_nerf_prediction (satellite.py:502-531, utm_sampling branch) from the reference's own functions."""
    import importlib
    import types
    ev = importlib.import_module("eval_eonerf")
    scale, offset = torch.tensor([143.5, 139.25, 51.0], dtype=torch.float64), torch.tensor([435500.5, 3354950.25, 12.5], dtype=torch.float64)
    out = {"scene_scale": scale.numpy(), "scene_offset": offset.numpy()}
    for tag, (h, w, ds, el, az) in {"a": (12, 10, 1.0, 27.5, 151.0), "b": (33, 48, 2.0, 61.0, 110.5)}.items():
        dataset = types.SimpleNamespace(scene_scale=scale, img_downscale=ds)
        rays = ev.create_rays_from_nadir(dataset, h, w, el, az)
        out.update({f"{tag}_h": np.int64(h), f"{tag}_w": np.int64(w), f"{tag}_downscale": np.float64(ds), f"{tag}_sun_el": np.float64(el),
                    f"{tag}_sun_az": np.float64(az), f"{tag}_rays": rays.numpy()})
    rays = torch.from_numpy(out["b_rays"])
    depth = torch.rand(rays.shape[0], 1, generator=torch.Generator().manual_seed(5)) * 2
    ds_self = types.SimpleNamespace(scene_scale=scale, scene_offset=offset, utm_sampling=True)
    e, n, a = ref.satellite.SatelliteDataset.get_utmalt_from_nerf_prediction(ds_self, rays, depth)
    out.update({"utm_depth": depth.numpy(), "utm_easts": e.numpy(), "utm_norths": n.numpy(), "utm_alts": a.numpy()})
    np.savez_compressed(os.path.join(GOLD, "nadir.npz"), **out)
    print("nadir: ok", out["a_rays"].shape, out["b_rays"].shape, e.dtype)


if __name__ == "__main__":
    torch.set_num_threads(8)
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_harness.load()
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:                       # python -m oracle.make_golden redraw vanilla
        for name in only:
            globals()["gen_" + name](ref)
        sys.exit(0)
    gen_sampling(ref)
    gen_field(ref)
    gen_volrend(ref)
    gen_render(ref)
    gen_redraw(ref)
    gen_vanilla(ref)
    gen_nadir(ref)
    print("golden fixtures written to", GOLD)
