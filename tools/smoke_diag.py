import sys, torch
sys.path.insert(0, "/root/repo")
from eonerf_code_b200 import _capi, metrics, sat_rendering
from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
from eonerf_code_b200.datasets.synthetic import make_rays
from eonerf_code_b200.radiance_fields import EONerfMLP
from oracle import eonerf_oracle as O
dev = torch.device("cuda:0")
for B in (256, 1024):
    n, n_img, epoch = 64, 5, 2
    p = O.init_params(n_img, seed=1, bias_scale=0.05)
    rays, ts, pixels = make_rays(B, n_img, seed=2)
    g = torch.Generator().manual_seed(3)
    u_cam, u_sun = torch.rand(B, n, generator=g), torch.rand(B, n, generator=g)
    _, _, g32, _ = O.train_step_grads(p, O.satrays_from_table(rays, ts), pixels, n, epoch, u_cam, u_sun)
    _, _, g16, _ = O.train_step_grads(p, O.satrays_from_table(rays, ts), pixels, n, epoch, u_cam, u_sun, emulate_bf16=True)
    G = {}
    for precision in ("bf16", "bf16_fused"):
        m = EONerfMLP(n_img, radiometric_normalization=True, precision=precision)
        m.load_state_dict(p, strict=False)
        m = m.to(dev)
        sr = define_satrays_from_tensors(rays.to(dev), ts.to(dev))
        res, nren = sat_rendering.render_image(m, None, sr, None, None, epoch_idx=epoch, chunk=B, render_step_size=2.0 / n,
                                               uniforms=[dict(u_cam=u_cam.to(dev), u_sun=u_sun.to(dev))], z_steps=torch.linspace(0, 1, n).to(dev))
        loss, _ = metrics.uncertainty_aware_loss(pixels.to(dev), res["rgb"], res["beta"])
        loss.backward()
        G[precision] = {k: v.grad.cpu() for k, v in m.named_parameters() if v.grad is not None}
    l2 = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
    print(f"B={B}: parameter | layered vs emu | fused vs emu | fused vs layered | emu vs fp32")
    for k in g16:
        if k in G["bf16"] and float(g16[k].norm()) > 0:
            print(f"  {k:45s} {l2(G['bf16'][k], g16[k]):9.2e} {l2(G['bf16_fused'][k], g16[k]):9.2e} {l2(G['bf16_fused'][k], G['bf16'][k]):9.2e} {l2(g16[k], g32[k]):9.2e}")
