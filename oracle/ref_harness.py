"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference Python from /root/reference.

Works only where /root/reference exists (the build container); nothing that runs on the GPU box
(`-m gpu` tests, smoke(), bench.py) imports this module.  The reference cannot be imported as-is
because it depends on packages that are absent here (nerfacc v0.5.2, rasterio, rpcm, ...).  We
register:
  * `nerfacc`, `nerfacc.volrend`  -> oracle/nerfacc_v052.py (restated published algorithm),
  * empty stub modules for the geo-IO imports that `datasets/satellite.py:1-19` executes at load.
No reference source is copied; modules are imported from where they lie.
"""
import importlib
import os
import sys
import types

REF = os.environ.get("EONERF_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "radiance_fields"))


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def load():
    """Returns a namespace with the reference's sat_rendering, eonerf, mlp, satellite, metrics modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    from . import nerfacc_v052
    nerfacc_v052.install_as_nerfacc()
    for name in ("rasterio", "rpcm", "utm", "pyproj", "plyflatten", "affine", "fire", "imageio",
                 "matplotlib", "matplotlib.pyplot", "cv2", "osgeo", "srtm4", "plyfile", "numba"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # the reference has a top-level `datasets` and `utils`; make sure ours/HF's do not shadow them
    for m in ("datasets", "datasets.satellite", "datasets.utils", "utils", "metrics", "sat_rendering",
              "radiance_fields", "radiance_fields.eonerf", "radiance_fields.mlp", "sat_utils"):
        sys.modules.pop(m, None)
    ns = types.SimpleNamespace()
    ns.mlp = importlib.import_module("radiance_fields.mlp")
    ns.eonerf = importlib.import_module("radiance_fields.eonerf")
    ns.satellite = importlib.import_module("datasets.satellite")
    ns.sat_rendering = importlib.import_module("sat_rendering")
    ns.metrics = importlib.import_module("metrics")
    return ns


class FixedRand:
    """Context manager: make torch.rand_like return queued tensors (the reference draws
    `torch.rand_like(z_vals)` inside perturb_z_vals, sat_rendering.py:52)."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        import torch
        self._orig = torch.rand_like
        q = self.queue

        def fake(t, *a, **k):
            u = q.pop(0)
            assert u.shape == t.shape, (u.shape, t.shape)
            return u.to(t.dtype)
        torch.rand_like = fake
        return self

    def __exit__(self, *exc):
        import torch
        torch.rand_like = self._orig
        return False
