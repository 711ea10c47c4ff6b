"""Zero-edit drop-in: make the reference's own entry points import the B200 product for the rendering hot path.

    cd /path/to/eonerf_code                      # the unmodified reference checkout
    PYTHONPATH=/path/to/this/repo/shim python train_eonerf.py --root_dir ... --model eo-nerf ...
    PYTHONPATH=/path/to/this/repo/shim python eval_eonerf.py ...

Python puts the script's directory in front of PYTHONPATH, so a plain path entry would lose against the reference's own
`sat_rendering.py` / `radiance_fields/` next to the script.  `sitecustomize.py` (imported automatically at interpreter
start-up because this directory is on PYTHONPATH) therefore calls `install()`, which registers a meta-path finder IN FRONT
of the path-based one for exactly the modules of the hot path:

    sat_rendering                  (/root/reference/sat_rendering.py)           -> eonerf_code_b200.sat_rendering
    radiance_fields[.eonerf|.mlp]  (/root/reference/radiance_fields/*.py)       -> eonerf_code_b200.radiance_fields.*
    nerfacc[.volrend]              (pip nerfacc v0.5.2, /root/reference/setup_env.sh:10) -> eonerf_code_b200.nerfacc_compat

Everything else the entry points import (opt, utils, metrics, datasets.satellite, sat_utils: RPC ray generation, geo I/O,
logging) stays the reference's own code (out of scope).  No CPU fallback: importing works anywhere,
calling anything needs a B200 and the built libeonerf_b200.so."""
import importlib.abc
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

_MODULES = {
    "sat_rendering": ("sat_rendering.py", False),
    "radiance_fields": (os.path.join("radiance_fields", "__init__.py"), True),
    "radiance_fields.eonerf": (os.path.join("radiance_fields", "eonerf.py"), False),
    "radiance_fields.mlp": (os.path.join("radiance_fields", "mlp.py"), False),
    "nerfacc": (os.path.join("nerfacc", "__init__.py"), True),
    "nerfacc.volrend": (os.path.join("nerfacc", "volrend.py"), False),
}


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        entry = _MODULES.get(fullname)
        if entry is None:
            return None
        rel, is_pkg = entry
        file = os.path.join(HERE, rel)
        return importlib.util.spec_from_file_location(fullname, file, submodule_search_locations=[os.path.dirname(file)] if is_pkg else None)


def install():
    if REPO not in sys.path:
        sys.path.append(REPO)                                  # for `import eonerf_code_b200`
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    for name in _MODULES:                                      # forget copies imported before install()
        m = sys.modules.get(name)
        if m is not None and not str(getattr(m, "__file__", "")).startswith(HERE):
            del sys.modules[name]
