"""CPU: the C-ABI library builds in-tree for sm_100a, loads, and exports every symbol include/eonerf_b200.h
declares; the ctypes structs agree with the header; the product refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "eonerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eonerf_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from eonerf_code_b200 import build, _capi
    path = build.build()
    assert os.path.exists(path)
    l = ctypes.CDLL(path)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/eonerf_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == names, set(_capi.SYMBOLS) ^ set(names)
    assert _capi.lib().eonerf_abi_version() == _capi.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "eonerf_b200.h")).read()
    assert f"#define EONERF_ABI_VERSION {_capi.ABI_VERSION}" in hdr


def test_struct_field_names_follow_header():
    from eonerf_code_b200 import _capi
    src = open(os.path.join(ROOT, "include", "eonerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, body in re.findall(r"typedef struct \{(.*?)\}\s*(\w+);", src, flags=re.S)[::1]:
        pass
    for body, cname in re.findall(r"typedef struct \{(.*?)\}\s*(\w+);", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.search(r"(\w+)\s*(\[\d+\])?$", decl)
            fields.append(m.group(1))
        py = getattr(_capi, cname.replace("Eonerf", ""))
        assert [f[0] for f in py._fields_] == fields, cname


def test_sizes_are_consistent():
    from eonerf_code_b200 import _capi as K
    l = K.lib()
    for prec, es in ((K.PREC_FP32, 4), (K.PREC_BF16, 2)):
        full = l.eonerf_field_stash_bytes(K.FIELD_EONERF, prec, 1000, 0)
        dens = l.eonerf_field_stash_bytes(K.FIELD_EONERF, prec, 1000, 1)
        assert full > dens >= 1000 * (7 * 256 + 320) * es
        assert l.eonerf_field_prepared_bytes(K.FIELD_EONERF, prec, 20) > 2 * 600_000 * es
    assert l.eonerf_field_stash_bytes(7, 0, 10, 0) == -1


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build container")
    from eonerf_code_b200 import sat_rendering
    o = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):
        sat_rendering.satnerf_sampling(o, o, {"render_step_size": 2 / 16})


def test_state_dict_contract():
    """Checkpoint keys/shapes == reference EONerfMLP, in named_parameters() order."""
    from eonerf_code_b200.radiance_fields import EONerfMLP
    from oracle import eonerf_oracle as O
    m = EONerfMLP(20, radiometric_normalization=True)
    sd = m.state_dict()
    ref = O.param_shapes(20)
    assert [k for k in sd if "scales" not in k] == list(ref)
    assert all(tuple(sd[k].shape) == v for k, v in ref.items())
    assert sd["posi_encoder.scales"].tolist() == [2 ** i for i in range(10)] and sd["posi_encoder.scales"].dtype == torch.int64
    assert sum(p.numel() for p in m.parameters()) == 679821
    r = sd["radiometricT_enc.weight"]
    assert torch.equal(r[:, :3], torch.ones(20, 3)) and torch.equal(r[:, 3:], torch.zeros(20, 6))


def test_utmalt_epilogue_has_no_cpu_fallback():
    """datasets/satellite.py:502-531 runs as an sm_100a kernel (tests/test_evalpost.py checks it bit-exactly against the
    reference's own fp64 outputs on the GPU); CPU tensors are refused like everywhere else in the product."""
    from eonerf_code_b200.datasets.satellite import get_utmalt_from_nerf_prediction
    from eonerf_code_b200.datasets.synthetic import make_rays
    rays, _, _ = make_rays(17, 5, seed=3)
    with pytest.raises(RuntimeError):
        get_utmalt_from_nerf_prediction(rays, torch.rand(17, 1), [140.0, 140.0, 50.0], [628000.0, 3357000.0, -20.0])


def test_prior_loss_terms_follow_the_reference_definitions():
    """metrics.py:9-58 (depth / shadow prior terms and the epoch window of auxiliary terms)."""
    from eonerf_code_b200 import metrics
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(200, 1, generator=g) * 2 - 0.3          # some negative = no prior
    pred = torch.rand(200, 1, generator=g)
    conf = torch.randint(0, 8, (200, 1), generator=g).float()
    t, d = metrics.depth_loss_L2(gt, pred, conf, w=100)
    keep = (gt >= 0) & (conf >= 4)
    assert torch.allclose(t, 100 * ((pred[keep] - gt[keep]) ** 2).mean()) and d["depth_weight"] == 100
    sm = (torch.rand(200, 1, generator=g) > 0.4).float()
    geo = torch.rand(200, 1, generator=g)
    t2, d2 = metrics.shadow_loss_L2(sm, geo)
    ref = ((sm <= 0.5).sum() / (sm >= 0).sum()) * (((sm <= 0.5) * (geo - sm) ** 2).sum() / ((sm <= 0.5).sum() + 1e-6))
    assert torch.allclose(t2, ref)
    assert torch.allclose(d2["shadow_vals_to_penalize"], (((geo > 0.2) & (sm < 0.5)).sum(0) / 200.0))
    base, logs = torch.tensor(1.0), {}
    l_in, logs = metrics.update_loss_with_aux_term(base, logs, t, d, epoch=3, start_epoch=2, end_epoch=5)
    l_out, _ = metrics.update_loss_with_aux_term(base, {}, t, d, epoch=5, start_epoch=2, end_epoch=5)
    assert torch.allclose(l_in, base + t) and torch.allclose(l_out, base) and "depth_l2" in logs
