#!/usr/bin/env python
"""Throughput of the single-pass blenders (feather + gain, no-blend) at BASELINE config-3 shape (4 x 1920x1080)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, panob200, util
from golden import calib
from oracle import compose

dev = torch.device("cuda", 0)
B = 64
frames = bench.synth_batch_torch(B, 99, dev)
Ks, Rs, scale = bench.calibration(os.environ.get("WORKLOAD", "config1"))
t = compose.build_tables(Ks, Rs, scale, (1920, 1080), "spherical")
masks = util.soft_masks(t)
res = {}
for blender in ("feather", "no"):
    st = panob200.ocvStitcher(panob200.StitcherConfig(width=1920, height=1080, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                                      blender=blender, num_bands=0, sharpness=0.02, cut=bench.CUT if os.environ.get("WORKLOAD", "config1") != "config3" else None, max_batch=64))
    assert st.initTables(masks) == 0, st.last_error
    for gains in (False, True):
        if gains:
            rng = np.random.default_rng(1)
            st.set_gain_maps([(0.8 + 0.4 * rng.random((s[1], s[0]))).astype(np.float32) for s in st.m_sizes])
        ow, oh = st.out_size
        out = torch.empty((B, oh, ow, 3), dtype=torch.uint8, device=dev)
        for _ in range(3):
            st.process_device(frames, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            st.process_device(frames, out)
        e1.record(); torch.cuda.synchronize()
        st.enable_profile(True); st.process_device(frames, out); torch.cuda.synchronize()
        prof = {p["name"]: round(p["ms"], 3) for p in st.read_profile()}; st.enable_profile(False)
        res["%s%s" % (blender, "+gain" if gains else "")] = {"panoramas_per_s": B * 10 / (e0.elapsed_time(e1) / 1e3), "ms_per_64": e0.elapsed_time(e1) / 10, "kernels_ms": prof}
print(json.dumps(res))
