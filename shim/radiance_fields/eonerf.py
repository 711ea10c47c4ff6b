"""`from radiance_fields.eonerf import EONerfMLP` (train_eonerf.py:10, eval_eonerf.py:46) -> the B200 product."""
from eonerf_code_b200.radiance_fields.eonerf import EONerfMLP  # noqa: F401
