"""Parameter containers with the reference's state_dict layout (/root/reference/radiance_fields/mlp.py:14-111,
168-250) and the vanilla NeRF field of BASELINE config 2.

These modules hold fp32 master weights only; they have no eager forward.  All arithmetic runs in the
sm_100a kernels behind the C ABI (csrc/field.cu, csrc/gemm_tc.cu)."""
import torch
import torch.nn as nn

from .. import _capi as K
from .. import ops

PRECISIONS = {"bf16": K.PREC_BF16, "fp32": K.PREC_FP32, "bf16_simt": K.PREC_BF16_SIMT, "bf16_fused": K.PREC_BF16_FUSED}


class LayerStack(nn.Module):
    """`hidden_layers.{i}` (+ `output_layer`): the key layout of the reference's MLP / DenseLayer
    (mlp.py:44-63,104-111).  Xavier-uniform weights, zero biases (mlp.py:22,25,28,67-85)."""

    def __init__(self, in_dims, width, out_dim=None, out_in=None):
        super().__init__()
        self.hidden_layers = nn.ModuleList([nn.Linear(k, width) for k in in_dims])
        if out_dim is not None:
            self.output_layer = nn.Linear(out_in if out_in is not None else width, out_dim)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, *a, **k):
        raise RuntimeError("LayerStack is a parameter container: the field modules run it through the CUDA engine")


class SinusoidalEncoder(nn.Module):
    """Keeps the `scales` buffer so state_dicts match (mlp.py:177-179); the encoding itself is computed in
    csrc/field.cu::encode_kernel."""

    def __init__(self, x_dim, min_deg, max_deg, use_identity=True):
        super().__init__()
        self.x_dim, self.min_deg, self.max_deg, self.use_identity = x_dim, min_deg, max_deg, use_identity
        self.register_buffer("scales", torch.tensor([2 ** i for i in range(min_deg, max_deg)]))

    @property
    def latent_dim(self):
        return (int(self.use_identity) + (self.max_deg - self.min_deg) * 2) * self.x_dim


def trunk_in_dims(enc_dim, width=256):
    # skip concat after layer index 4 (mlp.py:49-56 with skip_layer=4)
    return [enc_dim] + [width] * 4 + [width + enc_dim] + [width] * 2


class _EngineMixin:
    def _engine(self):
        eng = self.__dict__.get("_eng")
        names = dict(self.named_parameters())
        key = tuple((k, v.data_ptr()) for k, v in names.items()) + (self.precision,)
        if eng is None or self.__dict__.get("_eng_key") != key:
            eng = ops.FieldEngine(names, field=self._field_kind, precision=PRECISIONS[self.precision],
                                  n_images=getattr(self, "n_input_images", 0))
            self.__dict__["_eng"], self.__dict__["_eng_key"] = eng, key
        return eng


class NerfMLP(nn.Module):
    """Key layout of the reference's NerfMLP (mlp.py:114-165): base / sigma_layer / bottleneck_layer / rgb_layer."""

    def __init__(self, input_dim, condition_dim, net_width=256, net_width_condition=128):
        super().__init__()
        self.base = LayerStack(trunk_in_dims(input_dim, net_width), net_width)
        self.sigma_layer = LayerStack([], net_width, 1)
        self.bottleneck_layer = LayerStack([], net_width, net_width)
        self.rgb_layer = LayerStack([net_width + condition_dim], net_width_condition, 3)


class VanillaNeRFRadianceField(nn.Module, _EngineMixin):
    """mlp.py:211-250.  forward(x, condition) -> (rgb[N,3], sigma[N,1]); query_density(x) -> [N,1].
    precision: "bf16_fused" runs the first ten stages of the fused tcgen05 program (trunk, sigma head, bottleneck, rgb hidden layer
    with the view-direction term as a per-row bias, rgb head); "bf16" the layer-by-layer tcgen05 GEMMs; "fp32" the exactness mode."""
    _field_kind = K.FIELD_VANILLA

    def __init__(self, net_depth=8, net_width=256, skip_layer=4, net_depth_condition=1, net_width_condition=128,
                 precision="bf16"):
        super().__init__()
        if (net_depth, net_width, skip_layer, net_depth_condition, net_width_condition) != (8, 256, 4, 1, 128):
            raise ValueError("the sm_100a kernels are built for the 8x256 (+1x128) network every reference config uses")
        self.precision = precision
        self.posi_encoder = SinusoidalEncoder(3, 0, 10, True)
        self.view_encoder = SinusoidalEncoder(3, 0, 4, True)
        self.mlp = NerfMLP(self.posi_encoder.latent_dim, self.view_encoder.latent_dim, net_width, net_width_condition)

    def query_density(self, x):
        e = self._engine()
        return ops._FieldFn.apply(torch.is_grad_enabled(), e, True, x, None, None, *e.tensors())

    def query_opacity(self, x, step_size):
        return self.query_density(x) * step_size

    def forward(self, x, condition=None):
        if condition is None:
            raise ValueError("view directions are required (mlp.py:153-165)")
        e = self._engine()
        sigma, rgb = ops._FieldFn.apply(torch.is_grad_enabled(), e, False, x, None, condition, *e.tensors())
        return rgb, sigma
