"""How far are the two bf16 gradient paths (layer-by-layer, fused) from the fp32 kernels and from each other?  (GPU)"""
import sys, torch
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from helpers import make_model
from oracle import eonerf_oracle as O
cuda = torch.device('cuda:0')
def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))
for n in (100, 1000):
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
        f = e.fwd(n, False, x=x, img_idx=img)
        flat, views, gstruct = e.new_grads()
        e.bwd(n, False, f, g_sigma=gs, g_rgb=g3, g_ts=gts, g_tb=gtb, grads_struct=gstruct)
        torch.cuda.synchronize()
        res[mode] = views
    for k in ("transient_mlp.hidden_layers.0.weight", "transient_encoder.weight", "transient_mlp.hidden_layers.1.weight", "base_mlp.hidden_layers.3.weight", "albedo_mlp.hidden_layers.0.weight"):
        print(n, k, "layered-vs-fp32 %.4f  fused-vs-fp32 %.4f  fused-vs-layered %.4f" % (l2(res["bf16"][k], res["fp32"][k]), l2(res["bf16_fused"][k], res["fp32"][k]), l2(res["bf16_fused"][k], res["bf16"][k])))
