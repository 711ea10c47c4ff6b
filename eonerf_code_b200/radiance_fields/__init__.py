from .eonerf import EONerfMLP  # noqa: F401
from .mlp import VanillaNeRFRadianceField  # noqa: F401
