"""GPU: oracle-anchored parity of the BENCHMARKED path (precision bf16_fused, static=True, CUDA graph) at the sizes
BASELINE.json names — not a chain of small links.

Every test renders through the product with explicit uniforms and compares with oracle/eonerf_oracle.py (the restatement
pinned against the unmodified reference by tests/golden/) on the SAME rays, weights and uniforms:

* sample indices: n_rendering_samples and pts_per_ray bit-exact; sc_pts_per_ray bit-exact given the same rendered depth
  (the sun rays start at the rendered surface point, sat_rendering.py:90, so with a bf16 MLP the cube filter of a sun
  sample next to a face can flip: against the fp32 oracle's own depth the counts must still agree on >= 99 % of the rays
  and never differ by more than 2);
* floating-point outputs vs the fp32 oracle: >= 99 % of the elements of every output within 1e-3 absolute (north_star's
  bf16-MLP tolerance; it asks for >= 90 %), all within MAX_ABS below;
* gradients of one whole step vs the oracle's bf16-emulating autograd (same rounding points, fp32 accumulation):
  relative L2 distance per parameter tensor.

The oracle evaluates large chunks in 1024-ray slices (rays are independent; the two chunk-wide decisions of the reference
are taken over the whole chunk first: oracle.render_chunk_sliced)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import close, fingerprint, make_model, t
from oracle import eonerf_oracle as O

pytestmark = pytest.mark.gpu

FLOAT_KEYS = ("rgb", "depth", "albedo_rgb", "ambient_rgb", "geo_shadows", "transient_s", "beta", "shadowless_rgb")
# stated maxima of |product - fp32 oracle| per output (bf16 MLP, composited over <= 127 samples; depth is in ray units 0..2)
# (measured, profiles/r2a_parity_report.jsonl: rgb 2.3e-3, depth 3.4e-4, albedo 4.5e-4, geo_shadows 9.3e-3, beta 5.3e-4; the
# bounds leave a factor ~3).  geo_shadows = exp(-sum sigma*delta) over <= 127 sun samples amplifies the bf16 density error.
MAX_ABS = {"rgb": 6e-3, "depth": 1e-3, "albedo_rgb": 1.5e-3, "ambient_rgb": 1e-5, "geo_shadows": 3e-2, "transient_s": 1e-3,
           "beta": 2e-3, "shadowless_rgb": 1.5e-3}
MIN_WITHIN_1E3 = 0.99      # north_star asks for >= 0.90; measured >= 0.9966 on every output


def _report(name, payload):
    """Append the measured distances to gpurun_out/parity_report.jsonl (kept under profiles/ per round)."""
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, **payload}) + "\n")


def _uniforms(B, n, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.rand(B, n, generator=g) for k in ("u_cam", "u_sun", "u_cam2")}


def _render_product(m, rays, ts, n, epoch, us, cuda, eval=False, chunk=None, shape=None):
    from eonerf_code_b200 import sat_rendering
    from eonerf_code_b200.datasets.satellite import SatRays, define_satrays_from_tensors
    sr = define_satrays_from_tensors(rays.to(cuda), ts.to(cuda))
    if shape is not None:
        sr = SatRays(*[x.reshape(*shape, -1) for x in sr])
    res, nren = sat_rendering.render_image(m, None, sr, None, None, epoch_idx=epoch, chunk=chunk or rays.shape[0],
                                           render_step_size=2.0 / n, eval=eval, uniforms=[{k: v.to(cuda) for k, v in us.items()}],
                                           static=True)
    return res, int(nren)


def _compare_outputs(name, res, out_o, rows=None, rays=None, us=None, n=None):
    """res: product dict (flattened to [B,C]); out_o: oracle [b,21] for `rows` (None = all)."""
    from eonerf_code_b200 import sat_rendering
    stats = {}
    for k, a, b in sat_rendering.OUT_SLICES:
        mine = res[k].detach().reshape(-1, b - a).cpu()
        mine = mine if rows is None else mine[rows]
        ref = out_o[:, a:b]
        if k == "pts_per_ray":
            assert torch.equal(mine, ref), "pts_per_ray must be bit-exact"
        elif k == "sc_pts_per_ray":
            d = (mine - ref).abs()
            stats[k] = {"equal_frac": float((d == 0).float().mean()), "max_diff": float(d.max())}
            assert stats[k]["equal_frac"] >= 0.99 and stats[k]["max_diff"] <= 2, stats[k]
        elif k in ("entropy", "opacity_after_surface"):
            assert torch.equal(mine, ref), k
        else:
            d = (mine - ref).abs()
            stats[k] = {"within_1e-3": float((d <= 1e-3).float().mean()), "max_abs": float(d.max()), "mean_abs": float(d.mean())}
    _report(name, {"outputs": stats})
    bad = {k: v for k, v in stats.items() if k in MAX_ABS and (v["within_1e-3"] < MIN_WITHIN_1E3 or v["max_abs"] > MAX_ABS[k])}
    assert not bad, bad
    return stats


def _sun_counts_given_depth(rays, depth, n, u_sun):
    """sc_pts_per_ray the reference's sampler gives for THIS depth (sat_rendering.py:90-96): bit-exact target."""
    sr = O.satrays_from_table(rays, torch.zeros(rays.shape[0], 1, dtype=torch.long))
    sc_o = sr.origins + torch.hstack([depth, depth, depth]) * sr.viewdirs
    ri, _, _, _ = O.satnerf_sampling(sc_o, -1.0 * sr.sundirs, n, u_sun, near=None)
    return O.pts_per_ray(ri, rays.shape[0])


def test_cfg3_render_static_fused_vs_oracle(cuda):
    """BASELINE configs[2] (the bench workload): 8192 rays x n=128, epoch 2, 19 images, the bench's seeds."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    B, n, n_img, epoch = 8192, 128, 19, 2
    p = O.init_params(n_img, seed=42)
    m = make_model(p, n_img, cuda, "bf16_fused")
    rays, ts, _ = make_rays(B, n_img, seed=42)
    us = _uniforms(B, n, 1)
    with torch.no_grad():
        res, nren = _render_product(m, rays, ts, n, epoch, us, cuda)
        out_o, nren_o = O.render_chunk_sliced(p, O.satrays_from_table(rays, ts), n, epoch, **us)
    assert nren == nren_o
    _compare_outputs("cfg3_render", res, out_o)
    assert torch.equal(res["sc_pts_per_ray"].cpu()[:, 0], _sun_counts_given_depth(rays, res["depth"].cpu(), n, us["u_sun"]))


def test_cfg4_strip_eval_static_fused_vs_oracle(cuda):
    """A strip of BASELINE configs[3]: 128 rows x 1024 columns rendered as ONE 131072-ray chunk with eval=True under no_grad
    (the bench's render arm); the oracle checks every 64th ray (rays are independent; eval broadcasts the image index of the
    chunk's first ray, sat_rendering.py:288-289)."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    H, W, n, n_img, epoch = 128, 1024, 128, 19, 2
    B = H * W
    p = O.init_params(n_img, seed=42, bias_scale=0.02)
    m = make_model(p, n_img, cuda, "bf16_fused").eval()
    rays, ts, _ = make_rays(B, n_img, seed=7, eval_mode=True)
    us = _uniforms(B, n, 2)
    with torch.no_grad():
        res, nren = _render_product(m, rays, ts, n, epoch, us, cuda, eval=True, chunk=131072, shape=(H, W))
    for k, v in res.items():
        assert v.shape[:2] == (H, W) and v.dtype == torch.float32, k
    assert nren == int(res["pts_per_ray"].sum())
    rows = torch.arange(0, B, 64)
    sub = O.SatRays(*[r[rows] for r in O.satrays_from_table(rays, ts)])
    with torch.no_grad():
        out_o, _ = O.render_chunk_sliced(p, sub, n, epoch, us["u_cam"][rows], us["u_sun"][rows], us["u_cam2"][rows], eval=True)
    _compare_outputs("cfg4_strip", res, out_o, rows=rows)
    sc = _sun_counts_given_depth(rays[rows], res["depth"].reshape(B, 1).cpu()[rows], n, us["u_sun"][rows])
    assert torch.equal(res["sc_pts_per_ray"].reshape(B).cpu()[rows], sc)


def test_redraw_branch_static_fused_vs_oracle(cuda):
    """>= 3 camera rays keep no sample under the first draw (t_near = 3): the whole chunk is drawn again with near=None and the
    second set of uniforms (sat_rendering.py:259-262) — decided on the device in the static form.  pts_per_ray keeps the
    first draw's counts (zeros for those rays), n_rendering_samples counts the second draw."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    B, n, n_img, epoch = 2048, 128, 19, 2
    p = O.init_params(n_img, seed=43)
    m = make_model(p, n_img, cuda, "bf16_fused")
    rays, ts, _ = make_rays(B, n_img, seed=44)
    rays[[5, 700, 2047], 6] = 3.0
    rays[1000, 0:3] = torch.tensor([3.0, 3.0, 1.0])            # empty under both draws
    us = _uniforms(B, n, 3)
    sr = O.satrays_from_table(rays, ts)
    assert O.chunk_needs_redraw(sr, n, us["u_cam"])
    with torch.no_grad():
        res, nren = _render_product(m, rays, ts, n, epoch, us, cuda)
        out_o, nren_o = O.render_chunk_sliced(p, sr, n, epoch, **us)
    assert nren == nren_o
    assert int((res["pts_per_ray"] == 0).sum()) == 4
    _compare_outputs("redraw", res, out_o)


@pytest.mark.parametrize("precision,static", [("fp32", False), ("bf16_fused", False), ("bf16_fused", True)])
def test_redraw_golden(cuda, golden, precision, static):
    """The re-draw branch against the unmodified reference's own outputs (tests/golden/redraw.npz)."""
    from eonerf_code_b200 import metrics, sat_rendering
    from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
    g = golden["redraw"]
    n_img, n, epoch = int(g["n_img"]), int(g["n"]), int(g["epoch"])
    p = O.init_params(n_img, seed=21, bias_scale=0.05)
    np.testing.assert_allclose(fingerprint(p), g["fingerprint"], rtol=1e-12)
    m = make_model(p, n_img, cuda, precision)
    rays, ts = t(g["rays"], cuda), t(g["ts"], cuda)
    us = [dict(u_cam=t(g["u_cam"], cuda), u_cam2=t(g["u_cam2"], cuda), u_sun=t(g["u_sun"], cuda))]
    res, nren = sat_rendering.render_image(m, None, define_satrays_from_tensors(rays, ts), None, None, epoch_idx=epoch, chunk=rays.shape[0],
                                           render_step_size=2.0 / n, uniforms=us, z_steps=torch.linspace(0, 1, n).to(cuda), static=static)
    assert int(nren) == int(g["n_rendering_samples"])
    ref = t(g["out"])
    assert torch.equal(res["pts_per_ray"].cpu(), ref[:, 14:15])
    for k, a, b in sat_rendering.OUT_SLICES:
        if precision == "fp32":
            close(res[k], ref[:, a:b], 1e-5, 2e-6)
        elif k in FLOAT_KEYS:
            assert float((res[k].detach().cpu() - ref[:, a:b]).abs().max()) <= 5e-3, k
    if precision == "fp32":
        assert torch.equal(res["sc_pts_per_ray"].cpu(), ref[:, 15:16])
        loss = metrics.uncertainty_aware_loss(t(g["pixels"], cuda), res["rgb"], res["beta"])[0]
        close(loss, t(g["loss"]), 1e-5)
        loss.backward()
        assert [k for k, _ in m.named_parameters()] == [str(s) for s in g["grad_names"]]
        norms = np.array([float(v.grad.double().norm()) if v.grad is not None else 0.0 for _, v in m.named_parameters()])
        np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-2, atol=1e-8)


def _l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


# relative-L2 bars of the whole-step gradient vs the oracle's bf16-emulating autograd.  The per-image 9-vector rows and the
# narrow heads see every ray of the batch (well averaged); the sun-pass position gradient feeds the trunk's first layers
# through 2^k-weighted pos-enc terms, the worst conditioned part (see test_full_gradients_vs_oracle_fp32).
GRAD_L2 = 3e-2          # mid-size eager test (measured 1.0e-2)
GRAD_L2_CFG3 = 1e-2     # full-size captured step (measured 2.7e-3: more samples per parameter average the rounding)


def test_cfg3_step_gradients_graph_vs_oracle_bf16(cuda):
    """One cfg3-sized training step as the bench runs it (TrainStep(graph=True): captured and REPLAYED) with explicit uniforms
    and lr = 0 (Adam then leaves the parameters alone, so the replayed step's gradient buffer is the gradient at the initial
    parameters): loss and every parameter gradient vs O.train_step_grads_sliced(emulate_bf16=True)."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    from eonerf_code_b200.training import TrainStep
    B, n, n_img, epoch = 8192, 128, 19, 2
    p = O.init_params(n_img, seed=42, bias_scale=0.02)
    m = make_model(p, n_img, cuda, "bf16_fused")
    rays, ts, pixels = make_rays(B, n_img, seed=42)
    us = _uniforms(B, n, 4)
    step = TrainStep(m, n_samples=n, graph=True, lr=0.0)
    dev = lambda x: x.to(cuda)
    usd = {k: dev(v) for k, v in us.items()}
    for _ in range(3):                                   # eager, capture + replay, replay
        loss, nren = step(dev(rays), dev(ts), dev(pixels), epoch, uniforms=usd)
    torch.cuda.synchronize()
    for k, v in m.named_parameters():                    # lr = 0: parameters untouched
        assert torch.equal(v.detach().cpu(), p[k]), k
    loss_o, out_o, grads_o, nren_o = O.train_step_grads_sliced(p, O.satrays_from_table(rays, ts), pixels, n, epoch, emulate_bf16=True, **us)
    assert int(nren) == nren_o
    assert abs(float(loss) - float(loss_o)) <= 2e-3 * abs(float(loss_o)), (float(loss), float(loss_o))
    dist = {k: _l2(v.grad, grads_o[k]) for k, v in m.named_parameters() if float(grads_o[k].abs().max()) > 0}
    zero = [k for k, v in m.named_parameters() if float(grads_o[k].abs().max()) == 0 and float(v.grad.abs().max()) != 0]
    _report("cfg3_step_gradients", {"loss": float(loss), "loss_oracle": float(loss_o), "rel_l2": dist})
    assert not zero, zero
    bad = {k: e for k, e in dist.items() if e > GRAD_L2_CFG3}
    assert not bad, bad


def test_render_level_gradients_eager_vs_oracle_bf16(cuda):
    """The sun-pass -> depth -> compositing gradient chain in bf16 mode, eager form, mid size (1024 rays x n=64, both epochs'
    losses): every parameter gradient vs the bf16-emulating oracle."""
    from eonerf_code_b200.datasets.synthetic import make_rays
    from eonerf_code_b200.training import TrainStep
    B, n, n_img = 1024, 64, 7
    rays, ts, pixels = make_rays(B, n_img, seed=61)
    us = _uniforms(B, n, 5)
    out = {}
    for epoch in (0, 2):
        p = O.init_params(n_img, seed=62, bias_scale=0.05)
        m = make_model(p, n_img, cuda, "bf16_fused")
        step = TrainStep(m, n_samples=n, lr=0.0)
        loss, nren = step.eager(rays.to(cuda), ts.to(cuda), pixels.to(cuda), epoch, uniforms={k: v.to(cuda) for k, v in us.items()})
        loss_o, _, grads_o, nren_o = O.train_step_grads(p, O.satrays_from_table(rays, ts), pixels, n, epoch, emulate_bf16=True, **us)
        assert nren == nren_o
        assert abs(float(loss) - float(loss_o)) <= 2e-3 * abs(float(loss_o))
        dist = {k: _l2(v.grad, grads_o[k]) for k, v in m.named_parameters() if float(grads_o[k].abs().max()) > 0}
        out[epoch] = dist
        for k, v in m.named_parameters():
            if float(grads_o[k].abs().max()) == 0:
                assert float(v.grad.abs().max()) == 0.0, k
    _report("render_level_gradients", {"rel_l2": {str(e): d for e, d in out.items()}})
    bad = {(e, k): v for e, d in out.items() for k, v in d.items() if v > GRAD_L2}
    assert not bad, bad
