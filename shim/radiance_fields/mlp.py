"""`radiance_fields.mlp` (train_mlp_nerf.py:14: VanillaNeRFRadianceField) -> the B200 product."""
from eonerf_code_b200.radiance_fields.mlp import NerfMLP, SinusoidalEncoder, VanillaNeRFRadianceField  # noqa: F401
