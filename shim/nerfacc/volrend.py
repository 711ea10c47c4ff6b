"""`nerfacc.volrend` (radiance_fields/eonerf.py:15) -> eonerf_code_b200.nerfacc_compat."""
from eonerf_code_b200.nerfacc_compat import (accumulate_along_rays, render_transmittance_from_density,  # noqa: F401
                                             render_weight_from_density)
