#!/usr/bin/env python
"""Latency of the updateMask landing (pano_set_mask x N cameras + the table sync of the next process call) at
BASELINE config-1 size.  PANO_HOST_WEIGHTS=1 selects the host builder for comparison (read once per process)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import panob200  # noqa: E402
import util      # noqa: E402
from golden import calib  # noqa: E402
from oracle import compose  # noqa: E402

W, H = 1920, 1080
Ks, Rs, scale = calib.rig("2222", W)
t = compose.build_tables(Ks, Rs, scale, (W, H), "spherical")
masks = util.soft_masks(t)
st = panob200.ocvStitcher(panob200.StitcherConfig(width=W, height=H, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                                  blender="multiband", num_bands=5, cut=[0, 64, 5336, 896]))
assert st.initTables(masks) == 0, st.last_error
imgs = util.synth_set(4, H, W, 5)
ref = st.process(imgs)
ts = []
for r in range(5):
    t0 = time.perf_counter()
    for i, m in enumerate(masks):
        st.set_mask(i, m)
    ts.append(1e3 * (time.perf_counter() - t0))
assert np.array_equal(st.process(imgs), ref)
print(json.dumps({"what": "pano_set_mask x4 cameras, config-1 size (feed rects ~1.9k x 1.06k, 5 bands)",
                  "builder": "host" if os.environ.get("PANO_HOST_WEIGHTS") else "device",
                  "ms_median": float(np.median(ts)), "ms_all": [round(v, 2) for v in ts]}))
