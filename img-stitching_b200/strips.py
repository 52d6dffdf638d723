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
"""
import ctypes as C

from . import capi, sharding


class StripRank:
    """One rank's share: an initialised `ocvStitcher` + its column window."""

    def __init__(self, stitcher, rank, world, mode="exchange"):
        self.st, self.rank, self.world, self.mode = stitcher, rank, world, mode
        self.lib = capi.lib()
        self.h = stitcher._h
        nb, padded, _ = stitcher.blend_geometry()
        self.nb, self.padded = nb, padded
        self.x0, self.x1 = sharding.strip_columns(padded[0], nb, world)[rank]
        margin = 8 if mode == "exchange" else 3 * (1 << nb)
        capi.check(self.lib.pano_strip_set_window(self.h, self.x0, self.x1, margin), self.h)
        self.phases = self.lib.pano_strip_phase_count(self.h)
        self.has_left, self.has_right = rank > 0, rank < world - 1

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


def compose_local(ranks, frames, panos):
    """All ranks inside ONE process/GPU, stepped in lockstep (phase by phase); halo columns move
    with device-to-device copies.  Bit-identical to the NCCL path by construction: same phases,
    same pack/unpack kernels, same bytes."""
    import torch
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    allbufs = [_buffers(torch, r, frames.device) for r in ranks]
    for p in range(ranks[0].phases):
        for r, pano in zip(ranks, panos):
            r.run_phase(p, frames, pano, stream)
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
