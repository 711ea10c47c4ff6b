"""GPU: the sync-free ("static") form of the hot path and the CUDA-graph training step.

static=True keeps the sample counts on the device (kernels read P / Q there), turns the reference's "some ray kept no
sample -> draw again" branch (sat_rendering.py:259-262) into a device-side condition and accumulates the parameter
gradients straight into the flat gradient buffer.  Every result must equal the eager path's on the same uniforms."""
import pytest
import torch

from helpers import close, make_model, rel_err
from oracle import eonerf_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(B, n, n_img, cuda, seed=5, empty_rays=0):
    from eonerf_code_b200.datasets.synthetic import make_rays
    rays, ts, pixels = make_rays(B, n_img, seed=seed)
    if empty_rays:                                  # origins far outside the cube, marching away: every sample is discarded
        rays[:empty_rays, 0:3] = torch.tensor([5.0, 5.0, 5.0])
        rays[:empty_rays, 3:6] = torch.tensor([0.0, 0.0, 1.0])
    g = torch.Generator().manual_seed(seed + 1)
    us = [dict(u_cam=torch.rand(B, n, generator=g).to(cuda), u_sun=torch.rand(B, n, generator=g).to(cuda),
               u_cam2=torch.rand(B, n, generator=g).to(cuda))]
    return rays.to(cuda), ts.to(cuda), pixels.to(cuda), us


def _step(m, rays, ts, pixels, n, epoch, us, cuda, static):
    from eonerf_code_b200 import metrics, sat_rendering
    from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
    for p in m.parameters():
        p.grad = None
    res, nren = sat_rendering.render_image(m, None, define_satrays_from_tensors(rays, ts), None, None, epoch_idx=epoch,
                                           chunk=rays.shape[0], render_step_size=2.0 / n, uniforms=us,
                                           z_steps=torch.linspace(0, 1, n).to(cuda), static=static)
    loss = metrics.mse(pixels, res["rgb"]) if epoch < 2 else metrics.uncertainty_aware_loss(pixels, res["rgb"], res["beta"])[0]
    loss.backward()
    return res, int(nren), loss.detach(), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}


@pytest.mark.parametrize("epoch,empty", [(2, 0), (0, 0), (2, 3)])
def test_static_render_equals_eager(cuda, epoch, empty):
    B, n, n_img = 300, 64, 5
    p = O.init_params(n_img, seed=3, bias_scale=0.05)
    m = make_model(p, n_img, cuda, "bf16_fused")
    rays, ts, pixels, us = _inputs(B, n, n_img, cuda, empty_rays=empty)
    res_e, n_e, loss_e, g_e = _step(m, rays, ts, pixels, n, epoch, us, cuda, static=False)
    res_s, n_s, loss_s, g_s = _step(m, rays, ts, pixels, n, epoch, us, cuda, static=True)
    assert n_s == n_e and n_e > 0
    for k in res_e:
        close(res_s[k], res_e[k], 1e-6, 1e-7)           # same kernels on the same samples
    close(loss_s, loss_e, 1e-6)
    assert g_s.keys() == g_e.keys()
    for k in g_e:                                       # atomic accumulation order differs: fp32 round-off only
        assert rel_err(g_s[k], g_e[k]) < 2e-4, k


def test_grad_sink_equals_autograd_accumulation(cuda):
    """TrainStep routes every backward kernel into one flat gradient buffer; the sums must be those autograd builds."""
    from eonerf_code_b200.training import TrainStep
    B, n, n_img = 256, 64, 5
    p = O.init_params(n_img, seed=4, bias_scale=0.05)
    rays, ts, pixels, us = _inputs(B, n, n_img, cuda, seed=9)
    m_ref = make_model(p, n_img, cuda, "bf16_fused")
    _, _, loss_ref, g_ref = _step(m_ref, rays, ts, pixels, n, 2, us, cuda, static=False)
    m = make_model(p, n_img, cuda, "bf16_fused")
    step = TrainStep(m, n_samples=n)
    assert m._engine().grad_sink is not None
    step.grads.zero()
    _, _, loss, g = _step_keep_grads(m, rays, ts, pixels, n, 2, us, cuda)
    close(loss, loss_ref, 1e-6)
    for k in g_ref:
        assert rel_err(g[k], g_ref[k]) < 2e-4, k
        assert g[k].data_ptr() >= step.grads.flat.data_ptr()                   # still views of the flat buffer
        assert g[k].data_ptr() < step.grads.flat.data_ptr() + step.grads.flat.numel() * 4


def _step_keep_grads(m, rays, ts, pixels, n, epoch, us, cuda):
    from eonerf_code_b200 import metrics, sat_rendering
    from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
    res, nren = sat_rendering.render_image(m, None, define_satrays_from_tensors(rays, ts), None, None, epoch_idx=epoch,
                                           chunk=rays.shape[0], render_step_size=2.0 / n, uniforms=us,
                                           z_steps=torch.linspace(0, 1, n).to(cuda))
    loss = metrics.uncertainty_aware_loss(pixels, res["rgb"], res["beta"])[0]
    loss.backward()
    return res, nren, loss.detach(), {k: v.grad for k, v in m.named_parameters()}


def test_graph_train_step_matches_eager_steps(cuda):
    """Captured step == eager static step: same parameters after the same number of Adam steps on the same batches.
    (torch replays the device RNG with the offsets an eager run would use, so the stratified samples agree.)"""
    from eonerf_code_b200.training import TrainStep
    B, n, n_img, steps = 512, 64, 5, 5
    p = O.init_params(n_img, seed=6, bias_scale=0.05)
    batches = [_inputs(B, n, n_img, cuda, seed=20 + i)[:3] for i in range(2)]
    out = {}
    for mode in ("graph", "eager"):
        m = make_model(p, n_img, cuda, "bf16_fused")
        step = TrainStep(m, n_samples=n, graph=True)
        torch.manual_seed(123)
        losses = []
        for i in range(steps):
            r, t_, px = batches[i % 2]
            if mode == "graph":
                loss, nr = step(r, t_, px, 2)
            else:                                        # the same sync-free step, never captured
                loss, nr = step._forward_backward(r, t_, px, 2, static=True)
                step._update(averaged=True)
            losses.append(float(loss))
            assert int(nr) > 0
        out[mode] = (losses, {k: v.detach().clone() for k, v in m.named_parameters()})
        if mode == "graph":
            assert step.launches_per_step and step.launches_per_step > 10
            assert int(step.n_rendered_total) > steps * B
    l_g, p_g = out["graph"]
    l_e, p_e = out["eager"]
    assert all(abs(a - b) <= 2e-3 * max(1.0, abs(b)) for a, b in zip(l_g, l_e)), (l_g, l_e)
    assert l_g[-1] < l_g[0]                              # it trains
    lr = 5e-4
    for k in p_e:                                        # Adam's first steps are sign-like: where a gradient entry is ~0 the
        d = (p_g[k] - p_e[k]).abs()                      # atomic-order round-off decides the sign, so compare in bulk
        assert float(d.max()) <= 2 * lr * steps + 1e-7, k
        assert float((d > 0.5 * lr).float().mean()) < 0.05, (k, float((d > 0.5 * lr).float().mean()))


def test_flat_adam_matches_torch_adam(cuda):
    """csrc/optim.cu vs torch.optim.Adam (the reference's optimiser, train_eonerf.py:57): same parameters and moments after
    several steps on the same gradients, and a state_dict torch.optim.Adam can load."""
    from eonerf_code_b200.optim import FlatAdam
    from eonerf_code_b200.parallel import FlatGrads
    torch.manual_seed(0)
    shapes = [(19, 4), (19, 9), (256, 63), (256,), (1, 256), (1,), (3, 128), (128, 260)]
    pa = [torch.nn.Parameter(torch.randn(*s, device=cuda)) for s in shapes]
    pb = [torch.nn.Parameter(t.detach().clone()) for t in pa]
    fg = FlatGrads(pa)
    opt_a = FlatAdam(pa, fg.flat, lr=5e-4)
    opt_b = torch.optim.Adam(pb, lr=5e-4)
    for it in range(7):
        for a, b in zip(pa, pb):
            g = torch.randn_like(b) * (10.0 ** (it - 3))
            a.grad.copy_(g)
            b.grad = g.clone()
        opt_a.step()
        opt_b.step()
    for a, b in zip(pa, pb):
        close(a, b, 2e-6, 1e-7)
        for key in ("exp_avg", "exp_avg_sq"):           # fma contraction may differ from torch's kernels: fp32 round-off only
            ref = opt_b.state[b][key]
            close(opt_a.state[a][key], ref, 1e-5, 1e-6 * float(ref.abs().max()))
        assert float(opt_a.state[a]["step"]) == 7.0
    # checkpoint round trip in both directions (train_eonerf.py:185-191 saves optimizer.state_dict())
    opt_c = torch.optim.Adam([torch.nn.Parameter(t.detach().clone()) for t in pa], lr=5e-4)
    opt_c.load_state_dict(opt_a.state_dict())
    opt_a.load_state_dict(opt_b.state_dict())
    for a, b in zip(pa, pb):
        close(opt_a.state[a]["exp_avg_sq"], opt_b.state[b]["exp_avg_sq"], 0.0, 0.0)
        assert opt_a.state[a]["exp_avg"].data_ptr() >= opt_a.flat_exp_avg.data_ptr()     # still views of the flat buffer
    # gradient scaling (data-parallel mean after a sum all-reduce)
    for a, b in zip(pa, pb):
        g = torch.randn_like(b)
        a.grad.copy_(2.0 * g)
        b.grad = g.clone()
    opt_a.step(grad_scale=0.5)
    opt_b.step()
    for a, b in zip(pa, pb):
        close(a, b, 2e-6, 1e-7)


def test_device_ray_loader_is_a_shuffled_epoch(cuda):
    """DeviceRayLoader == DataLoader(shuffle=True) over the reference's training dataset (satellite.py:799-807): every row
    exactly once per epoch, same dict keys / shapes / dtypes, rows bit-identical to indexing the table, short last batch."""
    from eonerf_code_b200.datasets.device_loader import DeviceRayLoader
    from eonerf_code_b200.datasets.synthetic import make_rays
    n, B = 10_007, 1024
    rays, ts, rgbs = make_rays(n, 7, seed=31)
    gen = torch.Generator(device=cuda).manual_seed(5)
    loader = DeviceRayLoader(rays, rgbs, ts, B, device=cuda, generator=gen)
    assert len(loader) == (n + B - 1) // B
    seen = []
    for k, batch in enumerate(loader):
        b = batch["rays"].shape[0]
        assert b == (B if k < len(loader) - 1 else n - B * (len(loader) - 1))
        assert batch["rays"].shape == (b, 11) and batch["rgbs"].shape == (b, 3) and batch["ts"].shape == (b, 1)
        assert batch["ts"].dtype == torch.int64 and batch["idx"].dtype == torch.int64 and batch["rays"].is_cuda
        idx = batch["idx"].cpu()
        assert torch.equal(batch["rays"].cpu(), rays[idx]) and torch.equal(batch["rgbs"].cpu(), rgbs[idx])
        assert torch.equal(batch["ts"].cpu(), ts[idx])
        seen.append(idx)
    seen = torch.cat(seen)
    assert torch.equal(torch.sort(seen)[0], torch.arange(n))              # a permutation: without replacement
    assert not torch.equal(seen, torch.arange(n))                          # ... and shuffled
    second = torch.cat([b["idx"].cpu() for b in loader])
    assert not torch.equal(second, seen)                                   # a new permutation every epoch
    fixed = DeviceRayLoader(rays, rgbs, ts, B, shuffle=False, device=cuda)
    assert torch.equal(torch.cat([b["idx"].cpu() for b in fixed]), torch.arange(n))


@pytest.mark.parametrize("epoch", [0, 2])
def test_packed_loss_matches_reference_losses(cuda, epoch):
    """eonerf_loss_fwd_bwd vs metrics.uncertainty_aware_loss / mse (metrics.py:17-22, train_eonerf.py:139-143): value, the two
    terms of the loss_dict and the gradient wrt the packed renderer output."""
    from eonerf_code_b200 import metrics
    torch.manual_seed(3 + epoch)
    B = 5000
    out = torch.rand(B, 21, device=cuda)
    out[:, 12] = out[:, 12] * 2 + 0.05                      # beta >= beta_min
    gt = torch.rand(B, 3, device=cuda)
    a = out.clone().requires_grad_(True)
    b = out.clone().requires_grad_(True)
    loss_a, terms = metrics.packed_loss(a, gt, epoch)
    if epoch < 2:
        loss_b = metrics.mse(gt, b[:, 0:3])
    else:
        loss_b, ref_terms = metrics.uncertainty_aware_loss(gt, b[:, 0:3], b[:, 12:13])
        close(terms["coarse_color"], ref_terms["coarse_color"], 2e-6)
        close(terms["coarse_logbeta"], ref_terms["coarse_logbeta"], 2e-6)
    close(loss_a, loss_b, 2e-6)
    (3.0 * loss_a).backward()
    (3.0 * loss_b).backward()
    close(a.grad, b.grad, 1e-5, 2e-6 * float(b.grad.abs().max()))    # the two terms of d/d beta cancel: absolute floor
    assert float(a.grad[:, 3:12].abs().max()) == 0.0 and float(a.grad[:, 13:].abs().max()) == 0.0


def test_micro_batches_accumulate_to_the_full_batch_gradient(cuda):
    """TrainStep(micro_batch=b): slices of the batch rendered and back-propagated one after the other, gradients accumulated in
    the flat buffer, one Adam step — same loss and gradient as the whole batch at once (the losses are batch means)."""
    from eonerf_code_b200.training import TrainStep
    B, n, n_img = 384, 64, 5
    p = O.init_params(n_img, seed=12, bias_scale=0.05)
    rays, ts, pixels, us = _inputs(B, n, n_img, cuda, seed=40)
    res = {}
    for mb in (None, 128, 100):
        m = make_model(p, n_img, cuda, "bf16_fused")
        step = TrainStep(m, n_samples=n, micro_batch=mb)
        loss, nren = step._forward_backward(rays, ts, pixels, 2, static=False, uniforms=us[0])
        res[mb] = (float(loss), nren, step.grads.flat.clone())
    for mb in (128, 100):
        assert res[mb][1] == res[None][1]
        assert abs(res[mb][0] - res[None][0]) <= 2e-6 * abs(res[None][0])
        assert rel_err(res[mb][2], res[None][2]) < 3e-4, mb


@pytest.mark.parametrize("epoch,micro", [(0, None), (2, 256)])
def test_graph_train_step_variants(cuda, epoch, micro):
    """The captured step without the shadow pass (epoch < 2: MSE loss) and with micro-batches: it runs, trains, and counts
    the samples it rendered."""
    from eonerf_code_b200.training import TrainStep
    B, n, n_img, steps = 512, 64, 5, 6
    p = O.init_params(n_img, seed=9, bias_scale=0.05)
    rays, ts, pixels, _ = _inputs(B, n, n_img, cuda, seed=50)
    m = make_model(p, n_img, cuda, "bf16_fused")
    step = TrainStep(m, n_samples=n, graph=True, micro_batch=micro)
    torch.manual_seed(7)
    losses = []
    for _ in range(steps):
        loss, nr = step(rays, ts, pixels, epoch)
        losses.append(float(loss))
        assert int(nr) > B
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    assert int(step.n_rendered_total) > steps * B


def test_static_eval_render_equals_eager_and_keeps_no_activations(cuda):
    """Evaluation (eval=True under torch.no_grad(), the bench's render arm): the sync-free form gives the eager form's
    outputs, chunked, and neither keeps an activation stash (inference kernels: memory stays flat)."""
    from eonerf_code_b200 import sat_rendering
    from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
    from eonerf_code_b200.datasets.synthetic import make_rays
    B, n, n_img = 700, 64, 5
    p = O.init_params(n_img, seed=2, bias_scale=0.05)
    m = make_model(p, n_img, cuda, "bf16_fused").eval()
    rays, ts, _ = make_rays(B, n_img, seed=3, eval_mode=True)
    rays, ts = rays.to(cuda), ts.to(cuda)
    g = torch.Generator().manual_seed(8)
    us = [dict(u_cam=torch.rand(c, n, generator=g).to(cuda), u_sun=torch.rand(c, n, generator=g).to(cuda),
               u_cam2=torch.rand(c, n, generator=g).to(cuda)) for c in (256, 256, 188)]
    out = {}
    with torch.no_grad():
        for static in (False, True):
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.memory_allocated()
            res, nren = sat_rendering.render_image(m, None, define_satrays_from_tensors(rays, ts), None, None, epoch_idx=2, chunk=256,
                                                   render_step_size=2.0 / n, eval=True, uniforms=us,
                                                   z_steps=torch.linspace(0, 1, n).to(cuda), static=static)
            out[static] = (res, int(nren), torch.cuda.max_memory_allocated() - base)
    assert out[True][1] == out[False][1] > 0
    for k in out[False][0]:
        close(out[True][0][k], out[False][0][k], 1e-6, 1e-7)
    # a training-mode stash for one 256-ray chunk would be ~100 MB (6.3 KB x 16 k samples); inference stays far below
    assert out[True][2] < 40e6 and out[False][2] < 40e6, (out[True][2], out[False][2])


def test_graph_step_follows_lr_changes(cuda):
    """A scheduler (StepLR(gamma=0.9), train_eonerf.py:64,304) or a manual param_groups[0]['lr'] change must reach the
    captured Adam: lr lives in a device double the captured kernel reads, refreshed before each replay."""
    from eonerf_code_b200.training import TrainStep
    B, n, n_img = 256, 32, 4
    p = O.init_params(n_img, seed=14, bias_scale=0.05)
    rays, ts, pixels, us = _inputs(B, n, n_img, cuda, seed=70)
    lrs = [5e-4, 5e-4, 5e-4, 5e-5, 5e-5, 1e-3]
    res = {}
    for mode in ("graph", "eager"):
        m = make_model(p, n_img, cuda, "bf16_fused")
        step = TrainStep(m, n_samples=n, graph=(mode == "graph"))
        sched_seen = []
        for lr in lrs:
            step.optimizer.param_groups[0]["lr"] = lr
            if mode == "graph":
                step(rays, ts, pixels, 2, uniforms=us[0])
            else:
                step._forward_backward(rays, ts, pixels, 2, static=True, uniforms=us[0])
                step._update(averaged=True)
            sched_seen.append(float(step.optimizer._lr_buf))
        assert sched_seen == lrs
        res[mode] = {k: v.detach().clone() for k, v in m.named_parameters()}
    for k in res["eager"]:                               # same tolerance logic as test_graph_train_step_matches_eager_steps
        d = (res["graph"][k] - res["eager"][k]).abs()
        assert float(d.max()) <= 2 * sum(lrs) + 1e-7, k
        assert float((d > 0.5 * 5e-4).float().mean()) < 0.05, (k, float((d > 0.5 * 5e-4).float().mean()))
    # and the change is not a no-op: with lr frozen at capture (5e-4) the last three steps would move differently
    m = make_model(p, n_img, cuda, "bf16_fused")
    step = TrainStep(m, n_samples=n, graph=True)
    for _ in lrs:
        step(rays, ts, pixels, 2, uniforms=us[0])
    k = "base_mlp.hidden_layers.3.weight"
    frozen = dict(m.named_parameters())[k].detach()
    assert float((frozen - res["graph"][k]).abs().max()) > 1e-4


def test_eval_between_graph_steps_sees_current_weights(cuda):
    """Graph replays run Adam behind Python's back; an eval render between replays must use the weights the optimiser just
    wrote (the operand-layout cache is keyed on the parameters' versions, which TrainStep bumps after every replay)."""
    from eonerf_code_b200 import sat_rendering
    from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
    from eonerf_code_b200.training import TrainStep
    B, n, n_img = 256, 32, 4
    p = O.init_params(n_img, seed=15, bias_scale=0.05)
    rays, ts, pixels, us = _inputs(B, n, n_img, cuda, seed=71)
    m = make_model(p, n_img, cuda, "bf16_fused")
    step = TrainStep(m, n_samples=n, graph=True, lr=5e-3)

    def evaluate(model):
        with torch.no_grad():
            res, _ = sat_rendering.render_image(model, None, define_satrays_from_tensors(rays, ts), None, None, epoch_idx=2, chunk=B,
                                                render_step_size=2.0 / n, eval=True, uniforms=us,
                                                z_steps=torch.linspace(0, 1, n).to(cuda))
        return res["rgb"].clone()

    for round_ in range(3):
        for _ in range(3):
            step(rays, ts, pixels, 2, uniforms=us[0])
        got = evaluate(m)
        fresh = make_model({k: v.detach().cpu() for k, v in m.state_dict().items() if "scales" not in k}, n_img, cuda, "bf16_fused")
        want = evaluate(fresh)                           # a new module built from the current master weights
        assert torch.equal(got, want), round_


def test_flat_adam_state_dict_loads_into_stock_adam(cuda):
    """FlatAdam.state_dict() must not leak its shared step counter: a stock torch.optim.Adam that loads it increments `step`
    once per parameter per iteration (aliased tensors would be bumped 8x here)."""
    from eonerf_code_b200.optim import FlatAdam
    from eonerf_code_b200.parallel import FlatGrads
    torch.manual_seed(1)
    shapes = [(19, 4), (256, 63), (256,), (1, 256), (1,), (3, 128), (128, 260), (7,)]
    pa = [torch.nn.Parameter(torch.randn(*s, device=cuda)) for s in shapes]
    pb = [torch.nn.Parameter(t.detach().clone()) for t in pa]
    pc = [torch.nn.Parameter(t.detach().clone()) for t in pa]
    fg = FlatGrads(pa)
    opt_a, opt_b = FlatAdam(pa, fg.flat, lr=5e-4), torch.optim.Adam(pb, lr=5e-4)
    grads = [[torch.randn_like(b) for b in pb] for _ in range(5)]
    for it in range(3):
        for a, b, g in zip(pa, pb, grads[it]):
            a.grad.copy_(g)
            b.grad = g.clone()
        opt_a.step()
        opt_b.step()
    sd = opt_a.state_dict()
    steps = [st["step"] for st in sd["state"].values()]
    assert len({s.data_ptr() for s in steps}) == len(steps) and all(float(s) == 3.0 for s in steps)
    import io
    buf = io.BytesIO()
    torch.save(sd, buf)                                  # through a checkpoint file, as train_eonerf.py:185-191 does
    buf.seek(0)
    opt_c = torch.optim.Adam(pc, lr=5e-4)
    for c, a in zip(pc, pa):
        c.data.copy_(a.data)
    opt_c.load_state_dict(torch.load(buf))
    for it in (3, 4):
        for b, c, g in zip(pb, pc, grads[it]):
            b.grad = g.clone()
            c.grad = g.clone()
        opt_b.step()
        opt_c.step()
    for b, c in zip(pb, pc):
        assert float(opt_c.state[c]["step"]) == 5.0
        close(c, b, 2e-6, 1e-7)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_non_default_device(cuda):
    """gpu_id != 0 as the reference selects it (device=f"cuda:{args.gpu_id}", no set_device): kernels must run on the tensors'
    device and stream, and give the same numbers as on device 0."""
    from eonerf_code_b200 import sat_rendering
    from eonerf_code_b200.datasets.satellite import define_satrays_from_tensors
    B, n, n_img = 128, 32, 4
    p = O.init_params(n_img, seed=16, bias_scale=0.05)
    rays, ts, pixels, us = _inputs(B, n, n_img, cuda, seed=72)
    outs = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        m = make_model(p, n_img, dev, "bf16_fused")
        u = [{k: v.to(dev) for k, v in us[0].items()}]
        assert torch.cuda.current_device() == 0
        res, _ = sat_rendering.render_image(m, None, define_satrays_from_tensors(rays.to(dev), ts.to(dev)), None, None, epoch_idx=2,
                                            chunk=B, render_step_size=2.0 / n, uniforms=u, z_steps=torch.linspace(0, 1, n).to(dev))
        res["rgb"].sum().backward()
        outs.append((res["rgb"].detach().cpu(), m.base_mlp.hidden_layers[3].weight.grad.cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel_err(outs[1][1], outs[0][1]) < 2e-4
