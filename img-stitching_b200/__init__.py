"""B200-native per-frame compose path of Img-Stitching (package dir `img-stitching_b200`).

The directory name carries a hyphen (the reference's name), so import it through the
repo-root shim:  `from panob200 import pkg`  or  `import panob200` (see /panob200.py).

Layout: csrc/ (sm_100a kernels + C ABI), lib/ (built libpanob200.so), capi.py (ctypes binding of
include/panob200.h), stitcher.py (`ocvStitcher`-shaped host mirror), nvcam.py (`nvCam` pixel
pipeline mirror), ring.py (two-ring epilogue), sharding.py (frame-set sharding across ranks).
"""
from . import capi  # noqa: F401
from .capi import PanoError, build_library, library_path  # noqa: F401
from .stitcher import ocvStitcher, StitcherConfig  # noqa: F401
from .nvcam import nvCamFrontEnd  # noqa: F401
from .ring import FitCanvas, RingComposer  # noqa: F401
from . import sharding  # noqa: F401
from . import strips  # noqa: F401
