#!/usr/bin/env python
"""pano_process latency with PAGEABLE host buffers (what the reference's cv::Mat frames are), config 2 and config 1.
Environment knobs of the library: PANO_NO_HOST_STAGING=1 (the driver's own pageable path), PANO_HOST_THREADS=n,
PANO_HOST_NO_STREAM=1 (plain memcpy instead of streaming stores into the bounce buffer).  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
rec = {"env": {k: v for k, v in os.environ.items() if k.startswith("PANO_")}}
for wl in ("config2", "config1"):
    b = bench.build(wl, 0, dev, 4, 4, 1234)
    lat, _ = bench.time_latency(b, 100)
    rec[wl] = {"pinned_ms_p50": lat["pano_process_ms_p50"], "pageable_ms_p50": lat["pageable_host_buffers"]["pano_process_ms_p50"],
               "pageable_ms_p99": lat["pageable_host_buffers"]["pano_process_ms_p99"], "device_ms": lat["device_ms_per_frame_set"]}
    b["st"].close()
print(json.dumps(rec))
