"""`nerfacc` v0.5.2 as the reference imports it (sat_rendering.py:2, radiance_fields/eonerf.py:15, train_eonerf.py:13,
eval_eonerf.py:47, utils.py:12) -> the sm_100a kernels behind eonerf_code_b200.nerfacc_compat."""
from eonerf_code_b200.nerfacc_compat import (OccGridEstimator, accumulate_along_rays, pack_info,  # noqa: F401
                                             render_transmittance_from_density, render_weight_from_density)
from . import volrend  # noqa: F401

__version__ = "0.5.2+eonerf_b200"
rendering = None            # imported by sat_rendering.py:2 / utils.py, never called on the EO-NeRF path
