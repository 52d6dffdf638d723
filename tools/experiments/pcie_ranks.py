#!/usr/bin/env python
"""Raw PCIe rates with N ranks copying at the same time (one process per GPU, torchrun): what the host side of this
box can feed.  Per rank 1 GiB pinned buffers; H2D alone, D2H alone, both at once; plain pinned memory
(cudaHostAllocDefault) and write-combined input buffers (cudaHostAllocWriteCombined).  Rank 0 prints one JSON line
with per-rank minima and aggregates.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/experiments/pcie_ranks.py
"""
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist

N = 1 << 30


def host_alloc(nbytes, flags):
    """cudaHostAlloc through the runtime torch already loaded (torch has no write-combined allocator)."""
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        import glob
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        rt = ctypes.CDLL(cands[0])
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError("cudaHostAlloc failed: %d" % rc)
    return rt, p


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    d_in = torch.empty(N, dtype=torch.uint8, device=dev)
    d_out = torch.empty(N, dtype=torch.uint8, device=dev)
    h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    rt, wc = host_alloc(N, 0x04)                          # cudaHostAllocWriteCombined
    ctypes.memset(wc, 1, N)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, src="pinned", seconds=1.0):
        torch.cuda.synchronize()
        dist.barrier()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            if h2d:
                with torch.cuda.stream(s1):
                    if src == "pinned":
                        d_in.copy_(h_in, non_blocking=True)
                    else:
                        rt.cudaMemcpyAsync(ctypes.c_void_p(d_in.data_ptr()), wc, ctypes.c_size_t(N), ctypes.c_int(1),
                                           ctypes.c_void_p(s1.cuda_stream))
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            s1.synchronize(); s2.synchronize()
            n += 1
        dt = time.perf_counter() - t0
        gb = n * N / dt / 1e9
        t = torch.tensor([gb, -gb], dtype=torch.float64, device=dev)
        s = t.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"aggregate_GBps": float(s[0]), "min_rank_GBps": float(-t[1]), "max_rank_GBps": float(t[0])}

    run(True, True, seconds=0.3)
    rec = {"what": "raw cudaMemcpyAsync of 1 GiB pinned buffers, all ranks at once (tools/experiments/pcie_ranks.py)", "n_gpus": world,
           "h2d_only": run(True, False), "d2h_only": run(False, True), "both_directions_each": run(True, True),
           "h2d_only_write_combined": run(True, False, "wc"), "both_directions_write_combined_input": run(True, True, "wc")}
    try:
        rec["cpus"] = os.cpu_count()
        rec["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:
        pass
    if rank == 0:
        print(json.dumps(rec))
    rt.cudaFreeHost(wc)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
