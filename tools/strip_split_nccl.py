#!/usr/bin/env python
"""BASELINE config 4 across GPUs: ONE 8-camera 4K cylindrical 7-band panorama split into column
strips; pyramid halos exchanged (a) with NCCL point-to-point (torch.distributed batch_isend_irecv), (b) through
peer-memory mailboxes (P2P stores over NVLink + flags, no collective library on the data path), or (c) not at all
(redundant halo).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tools/strip_split_nccl.py [--small] [--steps K]

Checks on every rank that its own columns equal the undivided single-GPU result (computed locally
with a second handle), then times `exchange` (NCCL halo), `p2p` (peer-memory halo) and `redundant` (recomputed halo):
device time per panorama, max over ranks.  Prints one JSON line on rank 0.  The measurement itself lives in the
package (img-stitching_b200/strips.py: bench_config4), which bench.py also calls for its `strip_split` record."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true", help="960x540 cameras, 5 bands (quick check)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import panob200

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rec = panob200.pkg.strips.bench_config4(rank, world, local, small=args.small, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        print(json.dumps(rec))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
