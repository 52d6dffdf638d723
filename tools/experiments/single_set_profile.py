#!/usr/bin/env python
"""Per-kernel device time of ONE frame-set (CUDA events between the launches, pano_profile_enable): where the 0.25 ms of a
single pano_process call's kernel chain goes.  Prints one JSON line per workload."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream(dev)
for wl in ("config2", "config1"):
    b = bench.build(wl, 0, dev, 1, 1, 1234)
    st = b["st"]
    for _ in range(5):
        st.process_device(b["frames"], b["out"], stream.cuda_stream)
    torch.cuda.synchronize()
    acc = bench.profile_kernels(b, stream, reps=20)
    rec = {k: round(1000.0 * v["ms"] / v["launches"], 1) for k, v in acc.items()}
    print(json.dumps({"workload": wl, "us_per_kernel_one_frame_set": rec, "sum_us": round(sum(rec.values()), 1)}))
    st.close()
