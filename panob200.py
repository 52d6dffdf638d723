"""Import shim: the package directory is named `img-stitching_b200` (not a valid Python
identifier), so load it under the module name `img_stitching_b200` and re-export it."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "img-stitching_b200")
_NAME = "img_stitching_b200"


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = _load()
capi = pkg.capi
ocvStitcher = pkg.ocvStitcher
StitcherConfig = pkg.StitcherConfig
nvCamFrontEnd = pkg.nvCamFrontEnd
RingComposer = pkg.RingComposer
FitCanvas = pkg.FitCanvas
sharding = pkg.sharding
strips = pkg.strips
PanoError = pkg.PanoError
