"""EONerfMLP with the reference's constructor, methods and checkpoint layout
(/root/reference/radiance_fields/eonerf.py:69-248; state_dict contract in SURVEY.md Appendix B).

The module owns fp32 master parameters under the reference's names; every forward/backward runs in the
sm_100a kernels behind the C ABI.  `precision` selects the fused bf16 tensor-core kernels ("bf16_fused", production),
the layer-by-layer bf16 tensor-core kernels ("bf16", cross-check) or the fp32 exactness mode ("fp32") used for the 1e-5
parity tests."""
import torch
import torch.nn as nn

from .. import _capi as K
from .. import ops
from .mlp import LayerStack, SinusoidalEncoder, _EngineMixin, trunk_in_dims


class EONerfMLP(nn.Module, _EngineMixin):
    _field_kind = K.FIELD_EONERF

    def __init__(self, n_input_images, net_depth=8, net_width=256, skip_layer=4, radiometric_normalization=False,
                 precision="bf16_fused"):
        super().__init__()
        if (net_depth, net_width, skip_layer) != (8, 256, 4):
            raise ValueError("the sm_100a kernels are built for the 8x256 skip-4 network: the reference never builds "
                             "another one (--fc_units/--fc_layers are not read, train_eonerf.py:60-61)")
        self.n_input_images = n_input_images
        self.precision = precision
        self.pos_enc_L, self.view_enc_L = 10, 4
        self.posi_encoder = SinusoidalEncoder(3, 0, self.pos_enc_L, True)
        self.view_encoder = SinusoidalEncoder(3, 0, self.view_enc_L, True)
        self.transient_encoder = nn.Embedding(n_input_images, 4)
        self.beta_min = ops.BETA_MIN
        self.radiometric_normalization = radiometric_normalization
        if radiometric_normalization:                      # [1,1,1,0,...,0] per image, trainable (eonerf.py:91-94)
            init = torch.cat([torch.ones(n_input_images, 3), torch.zeros(n_input_images, 6)], dim=1)
            self.radiometricT_enc = nn.Embedding.from_pretrained(init, freeze=False)
        w, h = net_width, net_width // 2
        self.base_mlp = LayerStack(trunk_in_dims(self.posi_encoder.latent_dim, w), w)
        self.sigma_layer = LayerStack([], w, 1)
        self.bottleneck_layer = LayerStack([], w, w)
        self.albedo_mlp = LayerStack([w], h, 3)
        self.transient_mlp = LayerStack([w + 4, h, h, h], h)
        self.transient_scalar = LayerStack([], h, 1, out_in=h)
        self.transient_beta = LayerStack([], h, 1, out_in=h)
        self.ambient_mlp = LayerStack([self.view_encoder.latent_dim], h, 3)

    # --- point-wise API (eonerf.py:141-170) ---------------------------------------------------
    def query_density(self, x):
        e = self._engine()
        return ops._FieldFn.apply(torch.is_grad_enabled(), e, True, x, None, None, *e.tensors())

    def query_opacity(self, x, step_size):
        return self.query_density(x) * step_size

    def forward(self, x, sun_dirs=None, img_indices=None):
        e = self._engine()
        p = e.tensors()
        sigma, albedo, ts, tb = ops._FieldFn.apply(torch.is_grad_enabled(), e, False, x, img_indices.reshape(-1, 1), None, *p)
        ambient = ops._AmbientFn.apply(e, sun_dirs, *p)
        return sigma, albedo, ambient, ts, tb

    # --- per-ray API (eonerf.py:172-248) ------------------------------------------------------
    def _camera_pass(self, chunk_rays, t_starts, t_ends, ray_indices, only_depth):
        e = self._engine()
        n_rays = chunk_rays.origins.shape[0]
        offs = ops.pack_info(ray_indices, n_rays)
        return ops._CameraPassFn.apply(torch.is_grad_enabled(), e, only_depth, chunk_rays.origins, chunk_rays.viewdirs, chunk_rays.sundirs,
                                       chunk_rays.img_idx, ray_indices, t_starts, t_ends, offs, None, *e.tensors())

    def render_depth(self, chunk_rays, t_starts, t_ends, ray_indices):
        return self._camera_pass(chunk_rays, t_starts, t_ends, ray_indices, True)[:, 3:4]

    def rendering(self, chunk_rays, t_starts, t_ends, ray_indices, epoch_idx=100):
        """-> (albedo[B,3], depth[B,1], beta[B,1], transient_s[B,1], ambient[B,3], entropy[B,1]); like the reference,
        t_ends is updated in place (last sample of every ray -> 1e10, eonerf.py:220)."""
        comp = self._camera_pass(chunk_rays, t_starts, t_ends, ray_indices, False)
        depth = comp[:, 3:4]
        return comp[:, 0:3], depth, comp[:, 4:5], comp[:, 5:6], comp[:, 6:9], torch.ones_like(depth)
