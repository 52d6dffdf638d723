#!/usr/bin/env python
"""Experiment: does running two independent handles on two streams (what the reference's callers do with the
upper / lower ring stitchers, src/replay.cpp:284-288) raise device-resident throughput over one handle?
Config 2 (front end + 5-band compose), 64 frame-sets per step."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import panob200  # noqa: E402

dev = torch.device("cuda", 0)
B = 64
frames = bench.synth_batch_torch(B, 1234, dev)
frames = torch.cat([frames, torch.full(frames.shape[:-1] + (1,), 255, dtype=torch.uint8, device=dev)], dim=-1).contiguous()
Ks, Rs, scale = bench.calibration()


def make(max_batch):
    fe = bench.make_front_end(0, max_batch * 4)
    st = panob200.ocvStitcher(panob200.StitcherConfig(width=1920, height=1080, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                                      blender="multiband", num_bands=5, cut=bench.CUT, device=0, max_batch=max_batch, initMode=2))
    assert st.initTables() == 0, st.last_error
    st.attach_frontend(fe)
    return st, fe


def timeit(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return B * steps / (time.perf_counter() - t0)


res = {}
for mb in (8, 16):
    st, fe = make(mb)
    ow, oh = st.out_size
    out = torch.empty((B, oh, ow, 3), dtype=torch.uint8, device=dev)
    s0 = torch.cuda.Stream()
    res["one_handle_wave%d" % mb] = timeit(lambda: st.process_device(frames, out, s0.cuda_stream))
    ref = out.clone()
    st2, fe2 = make(mb)
    s1 = torch.cuda.Stream()
    h = B // 2

    def two():
        # interleave waves of the two handles so both streams always have work queued
        for b0 in range(0, h, mb):
            st.process_device(frames[b0:b0 + mb], out[b0:b0 + mb], s0.cuda_stream)
            st2.process_device(frames[h + b0:h + b0 + mb], out[h + b0:h + b0 + mb], s1.cuda_stream)
    res["two_handles_wave%d" % mb] = timeit(two)
    assert torch.equal(out, ref)
    st.close(); st2.close()
print(json.dumps(res))
