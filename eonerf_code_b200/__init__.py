"""B200-native EO-NeRF per-ray rendering hot path (drop-in for the reference's sat_rendering /
radiance_fields.eonerf / nerfacc operator trio).  See DESIGN.md and include/eonerf_b200.h."""
__version__ = "0.1.0"
