"""The C-ABI library loads and exports every symbol include/panob200.h declares (no compute)."""
import ctypes
import os
import re

import panob200
import util


def header_functions():
    txt = open(os.path.join(util.ROOT, "include", "panob200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pano_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    names = header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(panob200.capi.build_library())
    for n in names:
        assert hasattr(lib, n), "libpanob200.so does not export %s" % n
    assert sorted(panob200.capi.exported_symbols()) == names


def test_version_and_no_device_failure_is_loud():
    import torch
    lib = panob200.capi.lib()
    assert b"sm_100a" in lib.pano_version()
    if torch.cuda.is_available():
        return
    # no GPU here: creating a handle must FAIL (there is no CPU fallback), with a message
    import numpy as np
    import pytest
    from golden import calib
    Ks, Rs, sc = calib.rig("2222", 240)
    st = panob200.ocvStitcher(panob200.StitcherConfig(width=240, height=135, num_images=4, Ks=Ks, Rs=Rs,
                                                      warped_image_scale=sc, num_bands=5, blender="multiband", seam="no"))
    assert st.initSeam([np.zeros((135, 240, 3), np.uint8)] * 4) == -1
    assert "CUDA" in st.last_error
    with pytest.raises(panob200.PanoError):
        st.process([np.zeros((135, 240, 3), np.uint8)] * 4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(util.ROOT, "img-stitching_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("parity oracle", ""), "%s mentions the oracle" % f
