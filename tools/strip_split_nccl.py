#!/usr/bin/env python
"""BASELINE config 4 across GPUs: ONE 8-camera 4K cylindrical 7-band panorama split into column
strips; pyramid halos exchanged (a) with NCCL point-to-point (torch.distributed batch_isend_irecv), (b) through
peer-memory mailboxes (P2P stores over NVLink + flags, no collective library on the data path), or (c) not at all
(redundant halo).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tools/strip_split_nccl.py [--small] [--steps K]

Checks on every rank that its own columns equal the undivided single-GPU result (computed locally
with a second handle), then times `exchange` (NCCL halo), `p2p` (peer-memory halo) and `redundant` (recomputed halo):
device time per panorama, max over ranks.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true", help="960x540 cameras, 5 bands (quick check)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import panob200
    import util
    from golden import calib
    from oracle import compose  # only used for geometry-independent synthetic masks (init-time, host)

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H, nb = (960, 540, 5) if args.small else (3840, 2160, 7)
    Ks, Rs, scale = calib.ring(8, W, H, 65.2, 40.0, focal=None if args.small else 3000.0)
    t = compose.build_tables(Ks, Rs, scale, (W, H), "cylindrical")
    masks = util.soft_masks(t)
    imgs = [util.synth_frame(H, W, 400 + i, cell=64) for i in range(8)]
    frames = torch.from_numpy(np.stack(imgs)).to(dev)

    def stitcher():
        st = panob200.ocvStitcher(panob200.StitcherConfig(width=W, height=H, num_images=8, Ks=Ks, Rs=Rs,
                                                          warped_image_scale=scale, warp="cylindrical",
                                                          blender="multiband", num_bands=nb, device=local))
        assert st.initTables(masks) == 0, st.last_error
        return st

    ref = stitcher()
    ow, oh = ref.out_size
    want = torch.empty((1, oh, ow, 3), dtype=torch.uint8, device=dev)
    ref.process_device(frames.unsqueeze(0), want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ref.process_device(frames.unsqueeze(0), want)
    e1.record()
    torch.cuda.synchronize()
    single_ms = e0.elapsed_time(e1) / args.steps

    res = {}
    for mode in ("exchange", "p2p", "redundant"):
        r = panob200.strips.StripRank(stitcher(), rank, world, "exchange" if mode == "p2p" else mode)
        pano = torch.zeros((oh, ow, 3), dtype=torch.uint8, device=dev)
        if mode == "p2p":
            panob200.strips.p2p_setup_distributed(r)
            side = torch.cuda.Stream(dev)            # a capturable stream: the frame is replayed as a CUDA graph

            def run(bufs=None):
                side.wait_stream(torch.cuda.current_stream(dev))
                panob200.strips.compose_p2p(r, frames, pano, side.cuda_stream)
                torch.cuda.current_stream(dev).wait_stream(side)
        else:
            run = lambda bufs=None: panob200.strips.compose_nccl(r, frames, pano, bufs)  # noqa: E731
        bufs = run()
        torch.cuda.synchronize()
        c0, c1 = r.own_output_columns()
        ok = bool(torch.equal(pano[:, c0:c1], want[0][:, c0:c1]))
        for _ in range(args.warmup):
            run(bufs)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(args.steps):
            run(bufs)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / args.steps, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res[mode] = {"ms_per_panorama": float(tt[0]), "all_ranks_match_undivided": bool(tt[1] == 0.0),
                     "halo_bytes_per_rank_side": int(sum(r.halo_bytes(p) for p in range(r.phases)))}
    if rank == 0:
        print(json.dumps({"workload": "config4%s: 8x%dx%d cylindrical ring, %d bands, one panorama split into %d column strips"
                                      % (" (small)" if args.small else "", W, H, nb, world),
                          "n_gpus": world, "single_gpu_ms_per_panorama": single_ms, "modes": res,
                          "strips": panob200.sharding.strip_columns(r.padded[0], nb, world)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
