"""GPU: the tcgen05/TMEM/TMA GEMM building blocks (eonerf_linear_fwd / eonerf_linear_dw) against a plain torch fp32
reference of the same contraction on the same bf16-rounded operands (fp32 accumulation: 2e-3 relative to the row scale)
and against the SIMT twin.  Shapes are the ones the MLP issues plus ragged M / K tails."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def linear_fwd(prec, x, w, bias, relu, ldy=None):
    from eonerf_code_b200 import _capi as K
    from eonerf_code_b200.ops import _p, _stream
    M, Kd = x.shape
    N = w.shape[0]
    y = torch.zeros(M, ldy or N, dtype=x.dtype, device=x.device)
    a = K.LinearArgs(prec, _p(x), x.stride(0), _p(w), w.stride(0), _p(bias), M, N, Kd, int(relu), _p(y), y.stride(0))
    K.call("linear_fwd", a, _stream())
    return y[:, :N]


def linear_dw(prec, dy, x, n, k, with_bias=True):
    from eonerf_code_b200 import _capi as K
    from eonerf_code_b200.ops import _p, _stream
    dw = torch.zeros(n, k, dtype=torch.float32, device=x.device)
    db = torch.zeros(n, dtype=torch.float32, device=x.device) if with_bias else None
    a = K.DwArgs(prec, _p(dy), dy.stride(0), _p(x), x.stride(0), dy.shape[0], n, k, _p(dw), dw.stride(0), _p(db))
    K.call("linear_dw", a, _stream())
    return dw, db


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 256, 256), (4096, 256, 320), (777, 128, 128), (300, 256, 256),
                                   (5000, 320, 256), (129, 64, 256), (20000, 256, 256), (1, 256, 64), (640, 128, 288)])
def test_linear_fwd_tensor_core(cuda, M, N, K):
    from eonerf_code_b200 import _capi as Kc
    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g)).to(torch.bfloat16).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    ref = torch.relu(x.float() @ w.float().t() + b)
    y = linear_fwd(Kc.PREC_BF16, x, w, b, True)
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err      # bf16 output rounding: 2^-8 relative
    y2 = linear_fwd(Kc.PREC_BF16_SIMT, x, w, b, True)
    assert (y.float() - y2.float()).abs().max().item() <= 1.6e-2 * max(1.0, ref.abs().max().item())
    # without bias / activation, into a wider row (ldy > N)
    y3 = linear_fwd(Kc.PREC_BF16, x, w, None, False, ldy=N + 64)
    ref3 = x.float() @ w.float().t()
    assert (y3.float() - ref3).abs().max().item() <= 2e-2 * max(1.0, ref3.abs().max().item())


@pytest.mark.parametrize("M,N,K,ldx", [(128, 256, 256, 256), (1000, 256, 64, 320), (4096, 256, 320, 320), (777, 128, 128, 128),
                                       (5000, 128, 256, 256), (65, 256, 63, 320), (30000, 256, 256, 256), (900, 128, 283, 288),
                                       (2000, 256, 319, 320)])
def test_linear_dw_tensor_core(cuda, M, N, K, ldx):
    from eonerf_code_b200 import _capi as Kc
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.randn(M, N, generator=g).to(torch.bfloat16).to(cuda)
    xfull = torch.zeros(M, ldx, dtype=torch.bfloat16)
    xfull[:, :K] = torch.randn(M, K, generator=g).to(torch.bfloat16)
    xfull = xfull.to(cuda)
    ref = dy.float().t() @ xfull[:, :K].float()
    refb = dy.float().sum(0)
    dw, db = linear_dw(Kc.PREC_BF16, dy, xfull, N, K)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (dw - ref).abs().max().item() <= 2e-3 * scale, ((dw - ref).abs().max().item(), scale)
    assert (db - refb).abs().max().item() <= 2e-3 * refb.abs().max().item()
    dw2, _ = linear_dw(Kc.PREC_BF16_SIMT, dy, xfull, N, K)
    assert (dw - dw2).abs().max().item() <= 2e-3 * scale
    # accumulation semantics: a second call adds
    from eonerf_code_b200.ops import _p, _stream
    a = Kc.DwArgs(Kc.PREC_BF16, _p(dy), dy.stride(0), _p(xfull), xfull.stride(0), M, N, K, _p(dw), dw.stride(0), None)
    Kc.call("linear_dw", a, _stream())
    assert (dw - 2 * ref).abs().max().item() <= 4e-3 * scale


def test_misaligned_operands_are_rejected(cuda):
    from eonerf_code_b200 import _capi as Kc
    x = torch.zeros(64, 72, dtype=torch.bfloat16, device=cuda)[:, :63]
    w = torch.zeros(16, 63, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(RuntimeError):
        linear_fwd(Kc.PREC_BF16, x, w, None, False)
