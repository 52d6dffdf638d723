"""Column-strip split of ONE large panorama across GPUs (BASELINE config 4).

Every rank holds the full static tables, owns the padded-dst columns given by
`sharding.strip_columns`, and runs the compose phase by phase (`pano_strip_run_phase`).  After each
phase that produces pyramid data a neighbour needs, the ranks exchange the edge columns:

* `NcclExchange`   -- torch.distributed point-to-point (`batch_isend_irecv`, NCCL over NVLink);
* peer memory       -- `p2p_setup_*` + `compose_p2p`: every rank's mailbox is mapped by its neighbours (CUDA IPC); a
                      push kernel stores the edge columns straight into the neighbours' HBM over NVLink and raises a
                      flag, a wait/unpack kernel consumes them -- no collective library, no host sync on the data path;
* `LocalExchange`  -- all "ranks" are handles inside one process on one GPU (used to test the
                      decomposition bit-for-bit on a single device: ranks are stepped in lockstep,
                      no kernel ever waits on another);
* `mode="redundant"` -- no exchange at all: each rank widens its window by 3*2^nb columns and
                      recomputes the halo (SURVEY 8e option B).
* `mode="hybrid"`    -- thin redundant halos (3*2^split columns) below pyramid level `split`, the full width above it,
                      and ONE all-gather of the camera pyramids' level `split` per panorama (`compose_hybrid`): the 2 nb
                      latency-bound neighbour exchanges collapse into a single collective of a few hundred KB.
"""
import ctypes as C

from . import capi, sharding


class StripRank:
    """One rank's share: an initialised `ocvStitcher` + its column window."""

    def __init__(self, stitcher, rank, world, mode="exchange", split=None):
        self.st, self.rank, self.world, self.mode = stitcher, rank, world, mode
        self.lib = capi.lib()
        self.h = stitcher._h
        nb, padded, _ = stitcher.blend_geometry()
        self.nb, self.padded = nb, padded
        self.strips = sharding.strip_columns(padded[0], nb, world)
        self.x0, self.x1 = self.strips[rank]
        self.split = None
        if mode == "hybrid":
            self.split = max(1, min(nb, split if split is not None else nb - 2))
            capi.check(self.lib.pano_strip_set_window_hybrid(self.h, self.x0, self.x1, self.split), self.h)
        else:
            margin = 8 if mode == "exchange" else 3 * (1 << nb)
            capi.check(self.lib.pano_strip_set_window(self.h, self.x0, self.x1, margin), self.h)
        self.phases = self.lib.pano_strip_phase_count(self.h)
        self.has_left, self.has_right = rank > 0, rank < world - 1

    def cameras(self):
        """Indices of the cameras whose warped ROI meets this rank's window: the only frames the rank has to hold
        (the rest of the `frames` buffer is never read)."""
        n = self.st.m_cfg.num_images
        need = (C.c_int * n)()
        capi.check(self.lib.pano_strip_cameras(self.h, need), self.h)
        return [i for i in range(n) if need[i]]

    def halo_bytes(self, phase):
        return 0 if self.mode != "exchange" else int(self.lib.pano_strip_halo_bytes(self.h, phase))

    def run_phase(self, phase, frames, pano, stream):
        capi.check(self.lib.pano_strip_run_phase(self.h, phase, capi.ptr(frames), capi.ptr(pano), C.c_void_p(stream)), self.h)

    def pack(self, phase, side, buf, stream):
        capi.check(self.lib.pano_strip_halo_pack(self.h, phase, side, capi.ptr(buf), C.c_void_p(stream)), self.h)

    def unpack(self, phase, side, buf, stream):
        capi.check(self.lib.pano_strip_halo_unpack(self.h, phase, side, capi.ptr(buf), C.c_void_p(stream)), self.h)

    def own_output_columns(self):
        """Columns of the (cut) panorama this rank is responsible for: [c0, c1)."""
        cut = self.st.m_cutParams
        c0 = min(max(self.x0 - cut[0], 0), cut[2])
        c1 = min(max(self.x1 - cut[0], 0), cut[2])
        if self.rank == self.world - 1:
            c1 = cut[2]
        return c0, c1


def _buffers(torch, rank_obj, device):
    """send/recv staging per phase and side (allocated once)."""
    bufs = {}
    for p in range(rank_obj.phases):
        nbytes = rank_obj.halo_bytes(p)
        if nbytes:
            bufs[p] = {k: torch.empty(nbytes, dtype=torch.uint8, device=device) for k in ("sl", "sr", "rl", "rr")}
    return bufs


def compose_nccl(rank_obj, frames, pano, bufs=None):
    """One frame-set on this rank of a torch.distributed (NCCL) job.  frames: [N,H,W,3] uint8 on
    this rank's GPU; pano: [cut_h, cut_w, 3] uint8 (own columns valid afterwards)."""
    import torch
    import torch.distributed as dist
    r = rank_obj
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    if bufs is None:
        bufs = _buffers(torch, r, frames.device)
    for p in range(r.phases):
        r.run_phase(p, frames, pano, stream)
        if p not in bufs:
            continue
        b, ops = bufs[p], []
        if r.has_left:
            r.pack(p, 0, b["sl"], stream)
            ops += [dist.P2POp(dist.isend, b["sl"], r.rank - 1), dist.P2POp(dist.irecv, b["rl"], r.rank - 1)]
        if r.has_right:
            r.pack(p, 1, b["sr"], stream)
            ops += [dist.P2POp(dist.isend, b["sr"], r.rank + 1), dist.P2POp(dist.irecv, b["rr"], r.rank + 1)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if r.has_left:
            r.unpack(p, 0, b["rl"], stream)
        if r.has_right:
            r.unpack(p, 1, b["rr"], stream)
    return bufs


def compose_hybrid(rank_obj, frames, pano, bufs=None, gather=None):
    """One frame-set on this rank in hybrid mode.  `gather(send, recv)` all-gathers the ranks' chunks (default:
    torch.distributed.all_gather_into_tensor on the current stream, i.e. ncclAllGather); returns the reusable buffers."""
    import torch
    r = rank_obj
    L = r.split
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    lo = [x0 >> L for x0, _ in r.strips]
    n = [(x1 >> L) - (x0 >> L) for x0, x1 in r.strips]
    cmax = max(n)
    if bufs is None:
        nbytes = int(r.lib.pano_strip_level_bytes(r.h, L, cmax))
        bufs = {"send": torch.empty(nbytes, dtype=torch.uint8, device=frames.device),
                "recv": torch.empty(nbytes * r.world, dtype=torch.uint8, device=frames.device),
                "lo": (C.c_int * r.world)(*lo), "n": (C.c_int * r.world)(*n)}
    capi.check(r.lib.pano_strip_run_phases(r.h, 0, L, capi.ptr(frames), capi.ptr(pano), C.c_void_p(stream)), r.h)
    capi.check(r.lib.pano_strip_level_pack(r.h, L, lo[r.rank], cmax, capi.ptr(bufs["send"]), C.c_void_p(stream)), r.h)
    if gather is None:
        import torch.distributed as dist
        dist.all_gather_into_tensor(bufs["recv"], bufs["send"])
    else:
        gather(bufs["send"], bufs["recv"])
    capi.check(r.lib.pano_strip_level_unpack_all(r.h, L, capi.ptr(bufs["recv"]), cmax, bufs["lo"], bufs["n"], r.world, r.rank,
                                                 C.c_void_p(stream)), r.h)
    capi.check(r.lib.pano_strip_run_phases(r.h, L, r.phases, capi.ptr(frames), capi.ptr(pano), C.c_void_p(stream)), r.h)
    return bufs


def compose_hybrid_local(ranks, frames, panos):
    """All hybrid ranks inside ONE process/GPU in lockstep (tests): the all-gather is a device copy of every rank's chunk."""
    import torch
    L = ranks[0].split
    device = (frames[0] if isinstance(frames, (list, tuple)) else frames).device
    stream = torch.cuda.current_stream(device).cuda_stream
    world = len(ranks)
    lo = [x0 >> L for x0, _ in ranks[0].strips]
    n = [(x1 >> L) - (x0 >> L) for x0, x1 in ranks[0].strips]
    cmax = max(n)
    nbytes = int(ranks[0].lib.pano_strip_level_bytes(ranks[0].h, L, cmax))
    recv = torch.empty(nbytes * world, dtype=torch.uint8, device=device)
    per_rank = list(frames) if isinstance(frames, (list, tuple)) else [frames] * world      # each rank may hold its own cameras only
    for r, pano in zip(ranks, panos):
        capi.check(r.lib.pano_strip_run_phases(r.h, 0, L, capi.ptr(per_rank[r.rank]), capi.ptr(pano), C.c_void_p(stream)), r.h)
        capi.check(r.lib.pano_strip_level_pack(r.h, L, lo[r.rank], cmax, capi.ptr(recv[r.rank * nbytes:]), C.c_void_p(stream)), r.h)
    lo_c, n_c = (C.c_int * world)(*lo), (C.c_int * world)(*n)
    for r, pano in zip(ranks, panos):
        capi.check(r.lib.pano_strip_level_unpack_all(r.h, L, capi.ptr(recv), cmax, lo_c, n_c, world, r.rank, C.c_void_p(stream)), r.h)
        capi.check(r.lib.pano_strip_run_phases(r.h, L, r.phases, capi.ptr(per_rank[r.rank]), capi.ptr(pano), C.c_void_p(stream)), r.h)


def compose_local(ranks, frames, panos):
    """All ranks inside ONE process/GPU, stepped in lockstep (phase by phase); halo columns move
    with device-to-device copies.  Bit-identical to the NCCL path by construction: same phases,
    same pack/unpack kernels, same bytes."""
    import torch
    per_rank = list(frames) if isinstance(frames, (list, tuple)) else [frames] * len(ranks)
    device = per_rank[0].device
    stream = torch.cuda.current_stream(device).cuda_stream
    allbufs = [_buffers(torch, r, device) for r in ranks]
    for p in range(ranks[0].phases):
        for r, pano in zip(ranks, panos):
            r.run_phase(p, per_rank[r.rank], pano, stream)
        if p not in allbufs[0]:
            continue
        for i, r in enumerate(ranks):
            if r.has_left:
                r.pack(p, 0, allbufs[i][p]["sl"], stream)
            if r.has_right:
                r.pack(p, 1, allbufs[i][p]["sr"], stream)
        for i, r in enumerate(ranks):
            if r.has_left:
                allbufs[i][p]["rl"].copy_(allbufs[i - 1][p]["sr"])
                r.unpack(p, 0, allbufs[i][p]["rl"], stream)
            if r.has_right:
                allbufs[i][p]["rr"].copy_(allbufs[i + 1][p]["sl"])
                r.unpack(p, 1, allbufs[i][p]["rr"], stream)


# ------------------------------------------------------------------ peer-memory exchange (NVLink P2P stores + flags)

def p2p_setup_distributed(rank_obj):
    """One process per GPU (torch.distributed initialised): create the mailbox, ship its CUDA IPC handle to the
    neighbours and map theirs.  Collective: every rank must call it."""
    import torch.distributed as dist
    r = rank_obj
    hd = C.create_string_buffer(64)
    capi.check(r.lib.pano_strip_p2p_create(r.h, hd, None), r.h)
    handles = [None] * r.world
    dist.all_gather_object(handles, bytes(hd.raw))
    r._peer_handles = handles                      # keep the buffers alive
    capi.check(r.lib.pano_strip_p2p_connect(r.h, 0, handles[r.rank - 1] if r.has_left else None), r.h)
    capi.check(r.lib.pano_strip_p2p_connect(r.h, 1, handles[r.rank + 1] if r.has_right else None), r.h)
    dist.barrier()


def p2p_setup_local(ranks):
    """All ranks are handles of ONE process (tests): neighbours are connected by pointer."""
    for r in ranks:
        capi.check(r.lib.pano_strip_p2p_create(r.h, None, None), r.h)
    for i, r in enumerate(ranks):
        capi.check(r.lib.pano_strip_p2p_connect_local(r.h, 0, ranks[i - 1].h if r.has_left else None), r.h)
        capi.check(r.lib.pano_strip_p2p_connect_local(r.h, 1, ranks[i + 1].h if r.has_right else None), r.h)


def compose_p2p(rank_obj, frames, pano, stream=None):
    """One frame-set on this rank, halos through peer memory; asynchronous on `stream`."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream(frames.device).cuda_stream
    r = rank_obj
    capi.check(r.lib.pano_strip_run_p2p(r.h, capi.ptr(frames), capi.ptr(pano), C.c_void_p(stream)), r.h)


def compose_p2p_local(ranks, frames, panos, concurrent=False):
    """All ranks inside one process on one GPU.  concurrent=False: lockstep on one stream (every push precedes the
    matching wait, so no kernel ever spins).  concurrent=True: one stream per rank, each running its whole frame
    (`pano_strip_run_p2p`) -- the wait kernels really spin on flags that kernels of other streams raise."""
    import torch
    if concurrent:
        streams = getattr(compose_p2p_local, "_streams", None)
        if streams is None or len(streams) < len(ranks):
            streams = compose_p2p_local._streams = [torch.cuda.Stream(frames.device) for _ in ranks]
        cur = torch.cuda.current_stream(frames.device)
        # build every rank's graph BEFORE anything spins: instantiating a graph may synchronise the device, which must
        # not happen while another rank's wait kernel is waiting for this rank (one process = one context here)
        for s, r, pano in zip(streams, ranks, panos):
            capi.check(r.lib.pano_strip_p2p_prepare(r.h, capi.ptr(frames), capi.ptr(pano), C.c_void_p(s.cuda_stream)), r.h)
        for s, r, pano in zip(streams, ranks, panos):
            s.wait_stream(cur)
            compose_p2p(r, frames, pano, s.cuda_stream)
        for s in streams[:len(ranks)]:
            cur.wait_stream(s)
        return
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    for r in ranks:
        capi.check(r.lib.pano_strip_p2p_begin(r.h, C.c_void_p(stream)), r.h)
    for p in range(ranks[0].phases):
        for r, pano in zip(ranks, panos):
            r.run_phase(p, frames, pano, stream)
        if not ranks[0].halo_bytes(p):
            continue
        for r in ranks:
            capi.check(r.lib.pano_strip_p2p_push(r.h, p, C.c_void_p(stream)), r.h)
        for r in ranks:
            capi.check(r.lib.pano_strip_p2p_wait_unpack(r.h, p, C.c_void_p(stream)), r.h)


def assemble(ranks, panos):
    """Stitch the ranks' own columns back into one panorama (host side)."""
    import torch
    out = torch.empty_like(panos[0])
    for r, p in zip(ranks, panos):
        c0, c1 = r.own_output_columns()
        out[:, c0:c1] = p[:, c0:c1]
    return out


# ---------------------------------------------------------------------------------- BASELINE config 4 measurement
def ring_calibration(n, width, height, focal, step_deg):
    """Synthetic cylindrical ring of BASELINE config 4: shared K (f, width/2, height/2), R_i = R_y((i - (n-1)/2) * step)."""
    import numpy as np
    f = np.float32(focal)
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1]], np.float32)
    Rs = []
    for i in range(n):
        a = np.radians((i - (n - 1) / 2.0) * step_deg)
        c, s = np.cos(a), np.sin(a)
        Rs.append(np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32))
    return [K.copy() for _ in range(n)], Rs, float(f)


def soft_band_masks(stitcher):
    """Seam-like soft masks without a seam finder: each camera keeps the central vertical band of its warped footprint
    with a linear 0..255 ramp, min-ed with the warped validity mask the handle holds after initTables()."""
    import numpy as np
    out = []
    for i, (w, h) in enumerate(stitcher.m_sizes):
        x = np.arange(w)
        lo, hi = int(w * 0.18), int(w * 0.82)
        ramp = np.clip(np.minimum(x - lo, hi - x) * 12 + 128, 0, 255).astype(np.uint8)
        out.append(np.minimum(np.broadcast_to(ramp[None, :], (h, w)), stitcher.get_mask(i)).astype(np.uint8))
    return out


def bench_config4(rank, world, local, small=False, steps=10, warmup=3, modes=("exchange", "p2p", "redundant", "hybrid")):
    """ONE 8-camera cylindrical 7-band panorama (8 x 3840x2160; small: 8 x 960x540, 5 bands) split into `world` column
    strips, one per rank of an initialised torch.distributed NCCL job.  Every rank checks its own columns against the
    undivided panorama (computed locally by a second handle), then each halo mode is timed on the device, max over
    ranks: "exchange" = NCCL point-to-point, "p2p" = peer-memory mailboxes, "redundant" = recomputed halo.
    Returns the record on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import stitcher as stm
    dev = torch.device("cuda", local)
    W, H, nb, focal = (960, 540, 5, 750.0) if small else (3840, 2160, 7, 3000.0)
    Ks, Rs, scale = ring_calibration(8, W, H, focal, 40.0)

    masks = None

    def make():
        nonlocal masks
        st = stm.ocvStitcher(stm.StitcherConfig(width=W, height=H, num_images=8, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                                warp="cylindrical", blender="multiband", num_bands=nb, device=local))
        if st.initTables() != 0:
            raise capi.PanoError(st.last_error)
        if masks is None:
            masks = soft_band_masks(st)
        for i, m in enumerate(masks):
            st.set_mask(i, m)
        return st

    g = torch.Generator(device=dev)
    g.manual_seed(4242)                                  # every rank generates the SAME frame-set
    low = torch.rand((8, 3, H // 32 + 2, W // 32 + 2), generator=g, device=dev)
    up = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * 255.0
    frames = up.clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    frames_all = frames
    del low, up
    ref = make()
    ow, oh = ref.out_size
    want = torch.empty((1, oh, ow, 3), dtype=torch.uint8, device=dev)
    ref.process_device(frames.unsqueeze(0), want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ref.process_device(frames.unsqueeze(0), want)
    e1.record()
    torch.cuda.synchronize()
    single_ms = e0.elapsed_time(e1) / steps
    res = {}
    r = None
    for mode in modes:
        r = StripRank(make(), rank, world, "exchange" if mode == "p2p" else mode)
        pano = torch.zeros((oh, ow, 3), dtype=torch.uint8, device=dev)
        # the rank holds only the cameras its window meets (pano_strip_cameras); the others are noise it must never read
        need = r.cameras()
        frames = torch.full_like(frames_all, 0x5a)
        for i in need:
            frames[i] = frames_all[i]
        if mode in ("hybrid", "redundant"):
            side = torch.cuda.Stream(dev)            # a capturable stream: phase ranges are replayed as CUDA graphs

            def run(bufs=None, r=r, pano=pano, side=side, mode=mode, frames=frames):
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):        # NCCL's all-gather follows torch's current stream
                    if mode == "hybrid":
                        bufs = compose_hybrid(r, frames, pano, bufs)
                    else:                            # every phase from ONE library call (no exchange in between)
                        capi.check(r.lib.pano_strip_run_phases(r.h, 0, r.phases, capi.ptr(frames), capi.ptr(pano),
                                                               C.c_void_p(side.cuda_stream)), r.h)
                torch.cuda.current_stream(dev).wait_stream(side)
                return bufs
        elif mode == "p2p":
            p2p_setup_distributed(r)
            side = torch.cuda.Stream(dev)            # a capturable stream: the frame is replayed as a CUDA graph

            def run(bufs=None, r=r, pano=pano, side=side, frames=frames):
                side.wait_stream(torch.cuda.current_stream(dev))
                compose_p2p(r, frames, pano, side.cuda_stream)
                torch.cuda.current_stream(dev).wait_stream(side)
        else:
            def run(bufs=None, r=r, pano=pano, frames=frames):
                return compose_nccl(r, frames, pano, bufs)
        bufs = run()
        torch.cuda.synchronize()
        c0, c1 = r.own_output_columns()
        ok = bool(torch.equal(pano[:, c0:c1], want[0][:, c0:c1]))
        for _ in range(warmup):
            run(bufs)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(steps):
            run(bufs)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        if mode == "p2p":
            ok = ok and r.lib.pano_strip_p2p_check(r.h) == 0
        tt = torch.tensor([e0.elapsed_time(e1) / steps, 0.0 if ok else 1.0, float(len(need))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res[mode] = {"ms_per_panorama": float(tt[0]), "all_ranks_match_undivided": bool(tt[1] == 0.0),
                     "cameras_held_per_rank_max": int(tt[2]),
                     "halo_bytes_per_rank_side": int(sum(r.halo_bytes(p) for p in range(r.phases)))}
        if mode == "hybrid":
            L = r.split
            res[mode].update(split_level=L, redundant_margin_px=3 * (1 << L),
                             all_gather_bytes_per_rank=int(r.lib.pano_strip_level_bytes(r.h, L, max((x1 >> L) - (x0 >> L) for x0, x1 in r.strips))))
    t1 = torch.tensor([single_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t1, op=dist.ReduceOp.MAX)
    return {"workload": "config4%s: 8x%dx%d cylindrical ring, %d bands, one panorama split into %d column strips"
                        % (" (small)" if small else "", W, H, nb, world),
            "n_gpus": world, "single_gpu_ms_per_panorama": float(t1[0]), "modes": res,
            "halo_modes": {"exchange": "NCCL point-to-point (batch_isend_irecv) after each of the 2 nb phases", "p2p": "peer-memory mailboxes over NVLink, frame replayed as a CUDA graph",
                           "redundant": "no exchange, halo of 3 * 2^nb columns recomputed",
                           "hybrid": "thin redundant halo (3 * 2^split columns) below level `split`, full width above, ONE ncclAllGather of g[split] per panorama"},
            "strips": sharding.strip_columns(r.padded[0], nb, world)}
