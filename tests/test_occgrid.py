"""OccGridEstimator: the product's estimator follows nerfacc v0.5.2's update rule and buffer
layout (checked against the oracle's restatement on CPU: the estimator is host-side torch logic around `occ_eval_fn`), and
on the GPU the update runs through the fused density kernel and survives the reference's checkpoint round trip
(train_eonerf.py:185-191 -> eval_eonerf.py:49-71)."""
import io

import pytest
import torch

from oracle import nerfacc_v052 as nv


def _occ_fn(x):
    return (torch.exp(-4.0 * (x ** 2).sum(-1, keepdim=True)) * 0.05)


def test_update_rule_matches_restated_nerfacc_on_cpu():
    from eonerf_code_b200.nerfacc_compat import OccGridEstimator
    aabb = [-1., -1., -1., 1., 1., 1.]
    a, b = OccGridEstimator(roi_aabb=aabb, resolution=8, levels=1), nv.OccGridEstimator(roi_aabb=aabb, resolution=8, levels=1)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys()) == ["resolution", "aabbs", "occs", "binaries"]
    for k, v in a.state_dict().items():
        assert v.shape == b.state_dict()[k].shape and v.dtype == b.state_dict()[k].dtype, k
    assert a.grid_coords.shape == (512, 3) and torch.equal(a.grid_coords[1], torch.tensor([0, 0, 1]))
    for est in (a, b):
        torch.manual_seed(7)
        est.train()
        for step in (0, 1, 50, 100, 300, 350):                       # n=50 as train_eonerf.py:112-119; 300, 350 are past warm-up
            est.update_every_n_steps(step=step, occ_eval_fn=_occ_fn, n=50, occ_thre=1e-2)
    assert torch.equal(a.occs, b.occs) and torch.equal(a.binaries, b.binaries)
    assert 0 < int(a.binaries.sum()) < 512                            # the centre is occupied, the corners are not
    # step 1 is skipped (1 % 50 != 0): after step 0 alone the EMA-max leaves occs = occ at one jittered point per cell
    c = OccGridEstimator(roi_aabb=aabb, resolution=8)
    torch.manual_seed(7)
    c.update_every_n_steps(step=0, occ_eval_fn=_occ_fn, n=50)
    torch.manual_seed(7)
    x = (c.grid_coords + torch.rand_like(c.grid_coords, dtype=torch.float32)) / c.resolution * 2 - 1
    assert torch.allclose(c.occs, _occ_fn(x).squeeze(-1))
    c.eval()
    with pytest.raises(RuntimeError):
        c.update_every_n_steps(step=0, occ_eval_fn=_occ_fn)
    # checkpoint round trip in both directions
    buf = io.BytesIO()
    torch.save(a.state_dict(), buf)
    buf.seek(0)
    d = nv.OccGridEstimator(roi_aabb=aabb, resolution=8)
    d.load_state_dict(torch.load(buf))
    assert torch.equal(d.binaries, a.binaries) and torch.equal(d.occs, a.occs)
    e = OccGridEstimator(roi_aabb=aabb, resolution=8)
    e.load_state_dict(b.state_dict())
    assert torch.equal(e.binaries, b.binaries)


@pytest.mark.gpu
def test_update_with_fused_density_and_checkpoint_roundtrip(cuda, tmp_path):
    """train_eonerf.py:74,112-119 on the product: 128^3 cells = 2.1 M jittered points through query_opacity (fused density
    kernel, no_grad), then the checkpoint dict of :185-191 loaded the way eval_eonerf.py:49-71 does."""
    from helpers import make_model
    from eonerf_code_b200.nerfacc_compat import OccGridEstimator
    from eonerf_code_b200.radiance_fields import EONerfMLP
    from oracle import eonerf_oracle as O
    n_img = 5
    p = O.init_params(n_img, seed=3, bias_scale=0.05)
    m = make_model(p, n_img, cuda, "bf16_fused")
    aabb = [-1., -1., -1., 1., 1., 1.]
    grid = OccGridEstimator(roi_aabb=aabb, resolution=128, levels=1).to(cuda)
    step_size = 2.0 / 128
    torch.manual_seed(11)
    grid.update_every_n_steps(step=0, occ_eval_fn=lambda x: m.query_opacity(x, step_size), n=50, occ_thre=1e-2)
    assert grid.occs.shape == (128 ** 3,) and float(grid.occs.min()) > 0          # softplus density > 0 everywhere
    # same jittered points through the oracle's density on a sample of cells
    torch.manual_seed(11)
    x = (grid.grid_coords + torch.rand_like(grid.grid_coords, dtype=torch.float32)) / grid.resolution * 2 - 1
    sel = torch.arange(0, 128 ** 3, 4099, device=cuda)
    ref = O.query_density(p, x[sel].cpu()).squeeze(-1) * step_size
    got = grid.occs[sel].cpu()
    assert float((got - ref).abs().max()) <= 1e-3 * step_size + 5e-3 * float(ref.abs().max())
    ckpt = tmp_path / "epoch=0.ckpt"
    torch.save({"epoch": 0, "occ_grid_state_dict": grid.state_dict(), "model_state_dict": m.state_dict()}, ckpt)
    checkpoint = torch.load(ckpt)
    model = EONerfMLP(checkpoint["model_state_dict"]["radiometricT_enc.weight"].shape[0], radiometric_normalization=True)
    model.to(cuda)
    model.load_state_dict(checkpoint["model_state_dict"])                           # strict, as eval_eonerf.py:62
    occ = OccGridEstimator(roi_aabb=aabb, resolution=128, levels=1).to(cuda)
    occ.load_state_dict(checkpoint["occ_grid_state_dict"])
    assert torch.equal(occ.binaries, grid.binaries) and torch.equal(occ.occs, grid.occs) and occ.device == cuda
    # ... and into the oracle's stand-in for upstream nerfacc (what the reference's eval would construct)
    ref_occ = nv.OccGridEstimator(roi_aabb=aabb, resolution=128, levels=1)
    ref_occ.load_state_dict({k: v.cpu() for k, v in checkpoint["occ_grid_state_dict"].items()})
    assert int(ref_occ.binaries.sum()) == int(grid.binaries.sum())
