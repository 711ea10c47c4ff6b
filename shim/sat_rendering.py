"""`import sat_rendering` of the reference's entry points (train_eonerf.py:24, eval_eonerf.py:9) -> the B200 product."""
from eonerf_code_b200.sat_rendering import (SatRays, compute_geometric_shadows, count_number_of_pts_per_nerfacc_ray,  # noqa: F401
                                            namedtuple_map, render_image, satnerf_sampling)

# train_eonerf.py:24 also imports render_image_old (sat_rendering.py:337-391, never called): same signature, same renderer
render_image_old = render_image
