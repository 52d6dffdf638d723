#!/usr/bin/env python
"""panoramas/sec of the per-frame compose path (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU OpenCV path

A "step" = one pass of the hot path over one batch of `--waves-per-step` x `--batch` synthetic frame-sets per GPU
(default 16 waves x 64 frame-sets; 4 x 1920x1080 each; the 64 distinct frame-sets resident in HBM are cycled).
Default workload = BASELINE config 2 (imx390 undistort + crop + resize front end -> spherical warp -> 5-band
MultiBandBlender -> cut 5336x896), a superset of config 1; `also.config1` carries config 1 measured in the same process.
Frame-sets are sharded across ranks with no data-path collective (weak scaling: per-GPU work is fixed).

  value      whole-job frame-sets/s with inputs resident in HBM, timed on the device (CUDA events, barrier +
             synchronize on both sides, max over ranks); K steps of ~115 ms -> a timed region of > 2 s
  e2e        the same metric through the host-buffer API (pano_process_batch: pinned host frames in, panoramas back to
             pinned host memory, H2D/D2H inside the timed region), next to the raw PCIe rate N ranks reach at once
  latency    the drop-in call: pano_process, ONE frame-set per call, host buffers in and out (what src/replay.cpp:284-292
             calls), p50 / p99 per call
  roofline   dominant kernel: algorithmic bytes / CUDA-event time (frac) and ncu DRAM bytes / time (frac_dram)
  parity     the first GPU panoramas against the CPU arm's panoramas computed from the same bytes in this run
  cpu_baseline  cv2-driven reference call sequence on the host cores: faithful (maps rebuilt per call, the reference's
             behaviour) and cached-maps variants
  also.config3  BASELINE config 3 (imx424 rig, BlocksGainCompensator gains + FeatherBlender), device-resident
  also.config5  BASELINE config 5: ONE 256-frame-set batch sharded over the ranks (strong scaling), device-resident and
             streamed from pinned host memory
  strip_split   (N > 1) BASELINE config 4: one 8 x 4K cylindrical 7-band panorama split into N column strips, halos by
             NCCL point-to-point / peer-memory mailboxes / recomputed

One wave reads 2.1 GB of frames, far more than the 126 MB L2, so no L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "panoramas_per_sec"
UNIT = "panoramas/s"
W, H, NCAM, NBANDS = 1920, 1080, 4, 5
CUT = [0, 64, 5336, 896]
WORKLOADS = {
    "config1": "config1: 4x1920x1080 BGR -> spherical warp + 5-band MultiBandBlender -> cut 5336x896 (imx390 rig, 2222 calibration x4)",
    "config2": "config2: 4x1920x1080 BGRA camera frames -> imx390 undistort (INTER_CUBIC) + crop [69,103,1782,889] + resize -> "
               "spherical warp + 5-band MultiBandBlender -> cut 5336x896",
    # BASELINE config 3 (a parity-test configuration; measured on request)
    "config3": "config3: 4x1920x1080 BGR, imx424 rig (cfg/424camcfg/cameraparaout_1.txt x3) -> spherical warp -> BlocksGainCompensator "
               "apply -> FeatherBlender(sharpness = 1/blend_width, strength 5) -> whole panorama 6383x1131",
    # the YUYVCAM ingest (include/nvcam.hpp:880-886): frames cross PCIe as 8UC2, cvtColor(COLOR_YUV2BGRA_YUYV) runs on the device
    "config2-yuyv": "config2 with 8UC2 YUYV 4:2:2 camera frames (2 B/px over PCIe): YUV2BGRA_YUYV + imx390 undistort (INTER_CUBIC) + crop "
                    "[69,103,1782,889] + resize -> spherical warp + 5-band MultiBandBlender -> cut 5336x896",
    # NOT the parity path: the single-gather variant north_star asks to report separately, with its own PSNR
    "config2-fused": "config2 inputs, FUSED single-gather variant: undistort+crop+resize+warp maps composed into one table, one "
                     "bilinear gather from the BGRA frame -> 5-band MultiBandBlender -> cut 5336x896 (not bit-exact; see variant.psnr)",
}
# what ncu shows as the on-chip limiter of each hot kernel (profiles/r2*_ncu_summary.txt): the HBM
# roofline is the denominator the task asks for, but the exact-arithmetic gathers are not DRAM-limited
LIMITERS = {
    "fe_cubic_undistort": "L1 data pipes: TEX wavefronts 94 % (the 32-byte weight entry per pixel: 15 wavefronts per 16-byte fetch) + LSU wavefronts 77 % (16 shared-memory taps per pixel at ~1.8 wavefronts each)",
    "fe_resize": "memory latency at 38 % occupancy (72 registers, four output columns per lane), issue slots 57 %, DRAM 65 %",
    "warp": "L1 LSU wavefronts (4 shared-memory taps per pixel at ~2 wavefronts each) + issue slots",
    "pyrdown_l0": "memory latency at ~30 % occupancy (92 registers)",
    "collapse_l0": "issue slots + memory latency at ~40 % occupancy",
    "fe_yuyv_to_bgra": "HBM",
}
NEWK_FALLBACK = [[1627.5076, 0, 943.1681], [0, 1622.9720, 571.5369], [0, 0, 1]]   # SURVEY A10 (used when cv2 is absent)


def camera_entry():
    from golden import calib
    return calib.CAM_LIJING_390_FOV60_1920


def make_front_end(device, max_batch, src_format="bgra"):
    """nvCam front end of config 2 (cfg/cameras.yaml:80-88 entry, 1920x1080 in/out)."""
    import panob200
    cam = camera_entry()
    newK = None
    try:
        import cv2  # noqa: F401  (prepareUndistorMap's getOptimalNewCameraMatrix: one-time host init)
    except ImportError:
        newK = NEWK_FALLBACK
    cfg = panob200.pkg.nvcam.CamConfig(K=cam["K"], distorParams=cam["distorParams"], rect=cam["rect"], newK=newK,
                                       device=device, max_batch=max_batch, srcFormat=src_format)
    return panob200.nvCamFrontEnd(cfg)


def calibration(workload="config1"):
    from golden import calib
    return calib.rig("424" if workload == "config3" else "2222", W)


def config3_sharpness(dst_roi):
    """FeatherBlender::setSharpness(1 / blend_width), blend_width = sqrt(dst area) * strength / 100 (src/stitching_detailed.cpp:865-869)."""
    bw = np.float32(np.sqrt(np.float32(dst_roi[2] * dst_roi[3]))) * np.float32(5.0) / np.float32(100.0)
    return float(np.float32(1.0) / bw)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic():
    """ncu DRAM bytes per frame-set per kernel of the committed --set full capture (profiles/traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def synth_batch_torch(batch, seed, device):
    """Smooth full-range synthetic frames generated on the device: low-res noise, bilinear
    upsample, + mild per-pixel noise.  [batch, NCAM, H, W, 3] uint8."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    low = torch.rand((batch * NCAM, 3, H // 16 + 2, W // 16 + 2), generator=g, device=device)
    up = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
    up = up * 255.0 + (torch.rand(up.shape, generator=g, device=device) - 0.5) * 12.0
    return up.clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().view(batch, NCAM, H, W, 3)


def synth_numpy(seed):
    import util
    return util.synth_set(NCAM, H, W, seed)


def to_yuyv_numpy(f):
    y = np.empty(f.shape[:2] + (2,), np.uint8)
    y[:, :, 0] = f[:, :, 0]; y[:, 0::2, 1] = f[:, 0::2, 1]; y[:, 1::2, 1] = f[:, 0::2, 2]
    return y


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_setup(frames0, workload="config1"):
    """Static tables of the reference init flow (initSeam) for the CPU arm; cv2 path if present."""
    Ks, Rs, scale = calibration(workload)
    try:
        import cv2
        from oracle import cv2_reference as ref
        cv2.setNumThreads(os.cpu_count() or 1)
        t = ref.init_seam(frames0, Ks, Rs, scale, warp="spherical", seam="gc_color", want_gains=workload == "config3")
        return "cv2", t
    except ImportError:
        from oracle import compose
        t = compose.build_tables(Ks, Rs, scale, (W, H), "spherical")
        return "c_port", t


def cpu_reference_run(kind, t, frame_sets, repeats, front=False, workload="config1", cached_maps=False, keep=0):
    """The reference's per-frame path (ocvStitcher::process restated call for call; with front=True preceded by nvCam's
    resize/undistort/crop/resize per camera) on the host cores.
    -> dict(value, cores, sample, panoramas=[first `keep` results])."""
    n, panos = 0, []
    if kind == "cv2":
        import cv2
        from oracle import cv2_reference as ref
        fe = None
        if front:
            cam = camera_entry()
            _, mx, my = ref.undistort_tables(cam["K"], cam["distorParams"], (W, H))
            fe = lambda a: ref.front_end(a, (W, H), mx, my, cam["rect"], (W, H))     # noqa: E731
        maps = ref.build_warp_maps(t) if cached_maps else None

        def to_bgra(f):
            if f.shape[2] == 2:        # YUYVCAM ingest (include/nvcam.hpp:880-886)
                return cv2.cvtColor(np.ascontiguousarray(f), cv2.COLOR_YUV2BGRA_YUYV)
            return np.dstack([f, np.full(f.shape[:2], 255, np.uint8)]) if f.shape[2] == 3 else f

        def one(fs):
            if fe is not None:
                fs = [fe(to_bgra(f)) for f in fs]
            elif fs[0].shape[2] == 4:
                fs = [np.ascontiguousarray(f[:, :, :3]) for f in fs]
            if workload == "config3":    # warp -> compensator->apply -> FeatherBlender (src/stitching_detailed.cpp:829-871)
                return ref.process(t, fs, "feather", sharpness=config3_sharpness(t.dst_roi), apply_gain=True, maps=maps)
            return ref.process(t, fs, "multiband", NBANDS, cut=CUT, maps=maps)   # maps=None: rebuilt per call (:1171)
        for k in range(min(keep, len(frame_sets))):
            panos.append(one(frame_sets[k]))          # doubles as warm-up
        if not panos:
            one(frame_sets[0])
        t0 = time.perf_counter()
        for r in range(repeats):
            one(frame_sets[r % len(frame_sets)])
            n += 1
        dt = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        what = "cached maps (cv2.remap over maps built once)" if cached_maps else \
               "faithful (warper->warp rebuilds the maps on every call, as include/ocvstitcher.hpp:1171 does)"
        return {"value": n / dt, "cores": cores, "panoramas": panos,
                "sample": "%d panoramas in %.1f s, cv2 %s cv::detail classes driven by the restated ocvStitcher::process call "
                          "sequence, %s, cv2.setNumThreads(%d)" % (n, dt, cv2.__version__, what, cores)}
    from oracle import compose
    for k in range(min(keep, len(frame_sets))):
        panos.append(compose.process(t, [f[:, :, :3] for f in frame_sets[k]], "multiband", NBANDS, cut=CUT))
    t0 = time.perf_counter()
    for r in range(repeats):
        compose.process(t, [f[:, :, :3] for f in frame_sets[r % len(frame_sets)]], "multiband", NBANDS, cut=CUT)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "cores": 1, "panoramas": panos,
            "sample": "%d panoramas in %.1f s, scalar C port (oracle/pano_oracle.c), cached maps, 1 thread" % (n, dt)}


def cpu_baseline_record(kind, t, sets, front, workload, repeats, keep=0):
    """Both CPU variants of BASELINE.md section 3.  `value` is the faithful one (what the reference does)."""
    a = cpu_reference_run(kind, t, sets, repeats, front, workload, cached_maps=False, keep=keep)
    rec = {"value": a["value"], "unit": UNIT, "cores": a["cores"], "kind": "port", "sample": a["sample"],
           "faithful": {"value": a["value"], "sample": a["sample"]}}
    if kind == "cv2":
        b = cpu_reference_run(kind, t, sets, repeats, front, workload, cached_maps=True)
        rec["cached_maps"] = {"value": b["value"], "sample": b["sample"]}
    return rec, a["panoramas"]


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    per_step = 2
    sets = [[np.ascontiguousarray(f) for f in synth_numpy(1000 + s)] for s in range(2)]
    kind, t = cpu_reference_setup(sets[0], args.workload)
    if args.workload == "config2-yuyv":
        sets = [[to_yuyv_numpy(f) for f in s] for s in sets]
    front = args.workload.startswith("config2")
    cpu_reference_run(kind, t, sets, max(1, args.warmup), front, args.workload)
    t0 = time.perf_counter()
    a = cpu_reference_run(kind, t, sets, per_step * args.steps, front, args.workload)
    dt = time.perf_counter() - t0
    cpu = {"value": a["value"], "unit": UNIT, "cores": a["cores"], "kind": "port", "sample": a["sample"],
           "faithful": {"value": a["value"], "sample": a["sample"]}}
    if kind == "cv2":
        b = cpu_reference_run(kind, t, sets, max(4, per_step * args.steps // 2), front, args.workload, cached_maps=True)
        cpu["cached_maps"] = {"value": b["value"], "sample": b["sample"]}
    line = {"impl": "reference", "metric": METRIC, "value": a["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/s16 fixed-point + f32 weights", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "frame_sets_per_step": per_step},
            "cpu_baseline": cpu,
            "e2e": {"value": a["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ the CUDA arm
def numa_bind(local_rank):
    """Keep this rank's threads (and, by first touch, its pinned buffers) on the NUMA node its GPU hangs off."""
    info = {"bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        bus = out.lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        info["gpu_numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["numa_nodes"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
            info["bound"] = True
    except Exception as e:   # topology files absent in some containers: report, do not fail
        info["error"] = repr(e)[:120]
    return info


def build(workload, local_rank, dev, batch, max_batch, seed):
    """One-time init (host): calibration -> initSeam -> tables uploaded once.  -> dict of everything a measurement needs."""
    import torch
    import panob200
    Ks, Rs, scale = calibration(workload)
    frames = synth_batch_torch(batch, seed, dev)
    front = None
    if workload.startswith("config2"):
        # camera frames are 8UC4 (the VIC's ARGB output, include/nvcam.hpp:889-893): BGR + alpha 255
        if workload == "config2-yuyv":
            # valid YUYV bytes from the synthetic frames: Y = channel 0, U / V = channels 1 / 2 of the even pixel
            yuyv = torch.empty(frames.shape[:-1] + (2,), dtype=torch.uint8, device=dev)
            yuyv[..., 0] = frames[..., 0]
            yuyv[..., 0::2, 1] = frames[..., 0::2, 1]
            yuyv[..., 1::2, 1] = frames[..., 0::2, 2]
            frames = yuyv.contiguous()
            del yuyv
        else:
            frames = torch.cat([frames, torch.full(frames.shape[:-1] + (1,), 255, dtype=torch.uint8, device=dev)], dim=-1).contiguous()
        front = make_front_end(local_rank, max_batch * NCAM, "yuyv" if workload == "config2-yuyv" else "bgra")
    if front is None:
        set0 = [frames[0, i].cpu().numpy() for i in range(NCAM)]
    else:   # the stitcher calibrates on what the front end delivers
        set0 = [front.getFrame(frames[0, i].cpu().numpy()) for i in range(NCAM)]
    masks_how = "GraphCut seam masks (cv2, host init)"
    if workload == "config3":
        # init like src/stitching_detailed.cpp: seam masks + block gains from frame-set 0 (host, cv2), uploaded once
        from oracle import cv2_reference as ref
        t3 = ref.init_seam(set0, Ks, Rs, scale, warp="spherical", seam="gc_color", want_gains=True)
        sharp = config3_sharpness(t3.dst_roi)
        cfg = panob200.StitcherConfig(width=W, height=H, num_images=NCAM, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                      blender="feather", num_bands=0, sharpness=sharp, device=local_rank, max_batch=max_batch)
        st = panob200.ocvStitcher(cfg)
        rc = st.initTables(t3.blend_masks, None, ref.feather_weights(t3, sharp))
        if rc == 0:
            st.set_gain_maps(ref.full_res_gain_maps(t3))
    else:
        cfg = panob200.StitcherConfig(width=W, height=H, num_images=NCAM, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                      blender="multiband", num_bands=NBANDS, cut=CUT, device=local_rank,
                                      max_batch=max_batch, initMode=2)
        st = panob200.ocvStitcher(cfg)
        try:
            rc = st.calibration(set0)
        except ImportError:
            rc, masks_how = st.initTables(), "warped all-255 masks (cv2 absent)"
    if rc != 0:
        raise SystemExit("stitcher init failed: " + st.last_error)
    if front is not None:
        st.attach_frontend(front)
    ow, oh = st.out_size
    out = torch.empty((batch, oh, ow, 3), dtype=torch.uint8, device=dev)
    return dict(st=st, frames=frames, out=out, front=front, set0=set0, masks_how=masks_how, workload=workload)


def time_device(b, steps, waves, stream, barrier, sampler_rank=None):
    """K steps of `waves` waves each, CUDA events on the launching stream.  -> (ms, clocks or None)"""
    import torch
    st, frames, out = b["st"], b["frames"], b["out"]

    def step():
        for _ in range(waves):
            st.process_device(frames, out, stream.cuda_stream)

    sampler = None
    if sampler_rank is not None:
        # clocks / throttle reasons are sampled under the same load: a pre-roll of identical waves, then the K timed
        # steps, then a post-roll so that nvidia-smi samples land inside the loaded interval
        sampler = ClockSampler(sampler_rank)
        sampler.start()
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 0.3:
            st.process_device(frames, out, stream.cuda_stream)
            torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = None
    if sampler is not None:
        t_post = time.perf_counter()
        while time.perf_counter() - t_post < 0.2:
            st.process_device(frames, out, stream.cuda_stream)
            torch.cuda.synchronize()
        clocks = sampler.summary()
    return ms, clocks


def profile_kernels(b, stream, reps=3):
    """Per-kernel device time (CUDA events on the launching stream, pano_profile_enable) of single waves."""
    import torch
    st = b["st"]
    st.enable_profile(True)
    acc = {}
    for _ in range(reps):
        st.process_device(b["frames"], b["out"], stream.cuda_stream)
        torch.cuda.synchronize()
        for p in st.read_profile():
            a = acc.setdefault(p["name"], dict(ms=0.0, launches=0, bytes=0.0))
            a["ms"] += p["ms"]; a["launches"] += p["launches"]; a["bytes"] += p["alg_bytes"]
    st.enable_profile(False)
    return acc


def roofline_record(acc, sets_per_launch, peak, peak_src):
    total = sum(a["ms"] for a in acc.values()) or 1.0
    traffic = ncu_traffic()
    kernels = {}
    for k, a in acc.items():
        rec = {"ms_per_launch": a["ms"] / a["launches"], "share": a["ms"] / total, "GBps": a["bytes"] / a["ms"] / 1e6,
               "frac": a["bytes"] / a["ms"] / 1e6 / peak}
        if isinstance(traffic.get(k), (int, float)):
            rec["frac_dram"] = traffic[k] * sets_per_launch / (a["ms"] / a["launches"]) / 1e6 / peak
        kernels[k] = rec
    dom = max(acc, key=lambda k: acc[k]["ms"])
    d = acc[dom]
    per_set = traffic.get(dom) if isinstance(traffic.get(dom), (int, float)) else None
    return {"bound": "hbm", "kernel": dom, "achieved": d["bytes"] / d["ms"] / 1e6, "peak": peak, "peak_source": peak_src,
            "unit": "GB/s", "frac": d["bytes"] / d["ms"] / 1e6 / peak,
            "traffic": per_set * sets_per_launch if per_set is not None else None,
            "frac_dram": kernels[dom].get("frac_dram"),
            "traffic_source": traffic.get("_source", "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"),
            "limiter": LIMITERS.get(dom), "alg_bytes_per_launch": d["bytes"] / d["launches"], "ms_per_launch": d["ms"] / d["launches"],
            "whole_path_GBps": sum(a["bytes"] for a in acc.values()) / total / 1e6, "kernels": kernels}


def time_e2e(b, host_in, host_out, steps, barrier):
    import torch
    st = b["st"]
    st.process_batch(host_in, host_out)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        st.process_batch(host_in, host_out)     # synchronous: returns when the panoramas are in host memory
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt


def time_config5(b, host_in, host_out, world, stream, barrier, steps):
    """BASELINE config 5: ONE 256-frame-set batch sharded over the ranks (strong scaling: 256 / world frame-sets per rank per
    step), device-resident and streamed from pinned host memory.  -> (device ms, host-streamed s, frame-sets per rank)"""
    import torch
    st, frames, out = b["st"], b["frames"], b["out"]
    per = max(1, 256 // world)
    B = frames.shape[0]
    chunks = [min(B, per - o) for o in range(0, per, B)]

    def dev_step():
        for n in chunks:
            st.process_device(frames[:n], out[:n], stream.cuda_stream)

    def host_step():
        for n in chunks:
            st.process_batch(host_in[:n], host_out[:n])

    for _ in range(3):
        dev_step()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        dev_step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    host_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        host_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return ms, dt, per


def pcie_ceiling(host_in, host_out, dev_in, dev_out, barrier, seconds=0.6):
    """Raw rate of this rank's pinned buffers over PCIe with BOTH directions busy (what the e2e pipeline needs), all
    ranks copying at the same time.  -> (h2d GB/s, d2h GB/s) of this rank."""
    import torch
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    barrier()
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        with torch.cuda.stream(s1):
            dev_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s2):
            host_out.copy_(dev_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        n += 1
    dt = time.perf_counter() - t0
    barrier()
    return n * host_in.numel() / dt / 1e9, n * host_out.numel() / dt / 1e9


def time_latency(b, calls=200):
    """pano_process: one frame-set per call, host buffers in and out (synchronous) -- the call a drop-in user makes
    (src/replay.cpp:284-292).  Timed with pageable host buffers (what a cv::Mat holds) and with pinned ones; plus the
    device time of one frame-set's kernel chain (CUDA events around pano_process_device with batch 1)."""
    import torch
    st = b["st"]
    fr = b["frames"]
    nset = min(4, fr.shape[0])
    sets = [[np.ascontiguousarray(fr[k, i].cpu().numpy()) for i in range(NCAM)] for k in range(nset)]
    pinned = [[fr[k, i].cpu().pin_memory() for i in range(NCAM)] for k in range(nset)]
    ow, oh = st.out_size
    ret = np.empty((oh, ow, 3), np.uint8)
    ret_pin_t = torch.empty((oh, ow, 3), dtype=torch.uint8).pin_memory()
    ret_pin = ret_pin_t.numpy()

    def run(frame_sets, out, n):
        for k in range(5):
            st.process(frame_sets[k % nset], out)
        ts = []
        for k in range(n):
            t0 = time.perf_counter()
            st.process(frame_sets[k % nset], out)
            ts.append(time.perf_counter() - t0)
        return np.array(ts) * 1000.0

    tp = run(sets, ret, calls)
    tq = run([[t.numpy() for t in s] for s in pinned], ret_pin, calls)
    launches = st.last_launch_count()
    stream = torch.cuda.current_stream(fr.device)
    one_in, one_out = fr[:1], b["out"][:1]
    for _ in range(5):
        st.process_device(one_in, one_out, stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(50):
        st.process_device(one_in, one_out, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    rec = {"api": "pano_process (one frame-set per call, host frames in, panorama back in host memory; the kernel chain is replayed as a CUDA graph)",
           "calls": calls, "pano_process_ms_p50": float(np.percentile(tq, 50)), "pano_process_ms_p99": float(np.percentile(tq, 99)),
           "pano_process_ms_min": float(tq.min()), "panoramas_per_s": float(1000.0 / tq.mean()), "host_buffers": "pinned",
           "pageable_host_buffers": {"pano_process_ms_p50": float(np.percentile(tp, 50)), "pano_process_ms_p99": float(np.percentile(tp, 99)),
                                     "panoramas_per_s": float(1000.0 / tp.mean()),
                                     "how": ("the CUDA driver's own staging of pageable memory (PANO_NO_HOST_STAGING=1)" if os.environ.get("PANO_NO_HOST_STAGING")
                                             else "library worker threads -> pinned bounce buffers, chunked so that DMA overlaps the host copies")},
           "device_ms_per_frame_set": e0.elapsed_time(e1) / 50.0, "launches_per_call": launches,
           "h2d_bytes_per_call": int(fr[0].numel()), "d2h_bytes_per_call": int(ret.size)}
    return rec, ret


def parity_record(gpu_panos, cpu_panos, how):
    d = np.abs(gpu_panos.astype(np.int16) - cpu_panos.astype(np.int16))
    mse = float((d.astype(np.float64) ** 2).mean())
    return {"compared_with": how, "frame_sets_compared": int(gpu_panos.shape[0]), "bytes_compared": int(d.size),
            "max_abs_diff": int(d.max()), "mismatched_bytes": int(np.count_nonzero(d)),
            "psnr_db": None if mse == 0 else 10.0 * float(np.log10(255.0 * 255.0 / mse)), "bit_exact": bool(mse == 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="distinct frame-sets resident per GPU")
    ap.add_argument("--max-batch", type=int, default=64, help="frame-sets per launch wave (workspace: 50 MB per slot)")
    ap.add_argument("--waves-per-step", type=int, default=16, help="waves of --batch frame-sets per step (a step ~115 ms)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the config-1 / latency / strip-split side records")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to fd 1) are pointed at
    # stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = numa_bind(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    B, WAVES = args.batch, args.waves_per_step
    b = build(args.workload, local_rank, dev, B, args.max_batch, 1234 + rank)
    st, frames, out = b["st"], b["frames"], b["out"]
    stream = torch.cuda.current_stream(dev)
    variant = None
    if args.workload == "config2-fused":
        # the sequential (parity) result of the first frame-sets is the yardstick of the fused variant
        nref = min(B, 4)
        st.process_device(frames[:nref], out[:nref], stream.cuda_stream)
        torch.cuda.synchronize()
        seq = out[:nref].clone()
        st.set_frontend_mode(True)
        st.process_device(frames[:nref], out[:nref], stream.cuda_stream)
        torch.cuda.synchronize()
        d = (out[:nref].to(torch.float64) - seq.to(torch.float64))
        mse = float((d * d).mean())
        variant = {"name": "fused single-gather front end (PANO_FRONTEND_FUSED)", "compared_with": "sequential parity path, same inputs",
                   "psnr_db": (10.0 * float(np.log10(255.0 * 255.0 / mse)) if mse > 0 else None),
                   "max_abs_diff": int(d.abs().max()), "frac_bytes_within_1lsb": float((d.abs() <= 1).double().mean()),
                   "frame_sets_compared": nref}

    for _ in range(args.warmup):
        for _ in range(min(WAVES, 2)):
            st.process_device(frames, out, stream.cuda_stream)
    launches_per_wave = st.last_launch_count()
    torch.cuda.synchronize()
    ms, clocks = time_device(b, args.steps, WAVES, stream, barrier, sampler_rank=local_rank)
    prof_acc = profile_kernels(b, stream)
    gpu_first = out[:2].cpu().numpy()                      # panoramas of frame-sets 0, 1 (device path)

    # ---- end to end through the host-buffer API ----
    host_in = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
    host_in.copy_(frames)
    host_out = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
    e2e_steps = args.e2e_steps or max(2, int(round(2.5 * args.steps)))
    e2e_s = time_e2e(b, host_in, host_out, e2e_steps, barrier)
    same = bool(torch.equal(host_out[:2].to(dev), out[:2]))
    h2d_gbs, d2h_gbs = pcie_ceiling(host_in, host_out, frames, out, barrier)
    st.process_device(frames[:2], out[:2], stream.cuda_stream)      # (the raw copies overwrote nothing that is used again)
    torch.cuda.synchronize()

    ms_max, e2e_ms_max = allmax([ms, e2e_s * 1000.0])
    pcie_sum = allsum([h2d_gbs, d2h_gbs])
    pcie_min = [-v for v in allmax([-h2d_gbs, -d2h_gbs])]

    # ---- side records: config 1 in the same process, the drop-in call's latency, the strip split ----
    also, latency, strip = {}, None, None
    if not args.no_also:
        if rank == 0:
            latency, lat_pano = time_latency(b)
            latency["matches_device_path"] = bool(np.array_equal(lat_pano, out[(200 - 1) % min(4, B)].cpu().numpy())) if B >= 4 else None
        barrier()
        c5_steps = max(4, args.steps)
        c5_ms, c5_s, c5_per = time_config5(b, host_in, host_out, world, stream, barrier, c5_steps)
        c5_ms, c5_s = allmax([c5_ms, c5_s * 1000.0])
        also["config5"] = {"workload": "config5: ONE 256-frame-set batch of the default workload sharded over the ranks (strong scaling)",
                           "frame_sets_per_batch": c5_per * world, "frame_sets_per_rank": c5_per, "steps": c5_steps, "scaling": "strong",
                           "value": c5_per * world * c5_steps / (c5_ms / 1000.0), "unit": UNIT, "ms_per_batch": c5_ms / c5_steps,
                           "host_streamed": {"value": c5_per * world * c5_steps / (c5_s / 1000.0), "unit": UNIT,
                                             "ms_per_batch": c5_s / c5_steps, "api": "pano_process_batch"}}
        if args.workload != "config1":
            del host_in, host_out
            b1 = build("config1", local_rank, dev, B, args.max_batch, 1234 + rank)
            for _ in range(3):
                b1["st"].process_device(b1["frames"], b1["out"], stream.cuda_stream)
            l1 = b1["st"].last_launch_count()
            ms1, _ = time_device(b1, max(2, args.steps // 4), WAVES, stream, barrier)
            acc1 = profile_kernels(b1, stream)
            hi = torch.empty(b1["frames"].shape, dtype=torch.uint8).pin_memory(); hi.copy_(b1["frames"])
            ho = torch.empty(b1["out"].shape, dtype=torch.uint8).pin_memory()
            e1_steps = max(2, e2e_steps // 4)
            e1s = time_e2e(b1, hi, ho, e1_steps, barrier)
            ms1m, e1m = allmax([ms1, e1s * 1000.0])
            lat1 = time_latency(b1, 100)[0] if rank == 0 else None
            barrier()
            # config 5 as SURVEY 8(d) words it: config-1 tables, one 256-frame-set batch over the ranks
            k5_ms, k5_s, k5_per = time_config5(b1, hi, ho, world, stream, barrier, c5_steps)
            k5_ms, k5_s = allmax([k5_ms, k5_s * 1000.0])
            also["config5"]["config1_tables"] = {"value": k5_per * world * c5_steps / (k5_ms / 1000.0), "unit": UNIT,
                                                 "ms_per_batch": k5_ms / c5_steps,
                                                 "host_streamed": {"value": k5_per * world * c5_steps / (k5_s / 1000.0), "unit": UNIT}}
            peak, peak_src = peaks()
            r1 = roofline_record(acc1, min(B, args.max_batch), peak, peak_src)
            also["config1"] = {"workload": WORKLOADS["config1"], "value": world * B * WAVES * max(2, args.steps // 4) / (ms1m / 1000.0),
                               "unit": UNIT, "ms_per_wave": ms1m / (WAVES * max(2, args.steps // 4)), "gpu_launches_per_wave": l1,
                               "e2e": {"value": world * B * e1_steps / (e1m / 1000.0), "unit": UNIT},
                               "latency": lat1,
                               "roofline": {k: r1[k] for k in ("kernel", "frac", "frac_dram", "achieved", "ms_per_launch", "whole_path_GBps")},
                               "kernels": {k: {"ms_per_launch": v["ms_per_launch"], "frac": v["frac"], "frac_dram": v.get("frac_dram")}
                                           for k, v in r1["kernels"].items()}}
            b1["st"].close()
            del b1, hi, ho
        if args.workload != "config3":
            # BASELINE config 3 (imx424 rig, block gains + FeatherBlender) in the same process, device-resident
            b3, err3 = None, None
            try:
                b3 = build("config3", local_rank, dev, B, args.max_batch, 1234 + rank)
                for _ in range(3):
                    b3["st"].process_device(b3["frames"], b3["out"], stream.cuda_stream)
                torch.cuda.synchronize()
            except (Exception, SystemExit) as e:     # a side record must not take the headline down
                err3 = repr(e)[:300]
            # the timing below is collective (barriers): every rank takes part or none does
            if allmax([0.0 if err3 is None else 1.0])[0] == 0.0:
                s3 = max(2, args.steps // 4)
                ms3, _ = time_device(b3, s3, WAVES, stream, barrier)
                acc3 = profile_kernels(b3, stream)
                ms3m, = allmax([ms3])
                also["config3"] = {"workload": WORKLOADS["config3"], "value": world * B * WAVES * s3 / (ms3m / 1000.0), "unit": UNIT,
                                   "ms_per_wave": ms3m / (WAVES * s3), "gpu_launches_per_wave": b3["st"].last_launch_count(),
                                   "kernels_ms_per_launch": {k: v["ms"] / v["launches"] for k, v in acc3.items()}}
            else:
                also["config3"] = {"error": err3 or "another rank failed to initialise config 3"}
            if b3 is not None:
                b3["st"].close()
            del b3
        if world > 1:
            from panob200 import pkg
            try:
                strip = pkg.strips.bench_config4(rank, world, local_rank, small=False, steps=10, warmup=3)
            except Exception as e:                   # a side record must not take the headline down
                strip = {"error": repr(e)[:300]}

    if rank == 0:
        peak, peak_src = peaks()
        total_sets = world * B * WAVES * args.steps
        value = total_sets / (ms_max / 1000.0)
        e2e_value = world * B * e2e_steps / (e2e_ms_max / 1000.0)
        roof = roofline_record(prof_acc, min(B, args.max_batch), peak, peak_src)
        cpu, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            sets = [[frames[k, i].cpu().numpy() for i in range(NCAM)] for k in range(2)]
            kind, t = cpu_reference_setup([f[:, :, :3] for f in sets[0]] if b["front"] is None else b["set0"], args.workload)
            cpu, cpu_panos = cpu_baseline_record(kind, t, sets, b["front"] is not None, args.workload,
                                                 20 if kind == "cv2" else 4, keep=2)
            if args.workload != "config2-fused" and len(cpu_panos) == 2 and cpu_panos[0].shape == gpu_first[0].shape:
                parity = parity_record(gpu_first, np.stack(cpu_panos),
                                       "the CPU reference arm's panoramas of the same two frame-sets, computed in this run (%s)" % kind)
        in_bytes, out_bytes = int(frames[0].numel()), int(out[0].numel())
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/s16 fixed-point + f32 weights", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload], "frame_sets_per_gpu_per_step": B * WAVES, "waves_per_step": WAVES,
                           "frame_sets_per_wave": min(B, args.max_batch), "distinct_frame_sets_resident": B, "masks": b["masks_how"],
                           "timed_region_s": ms_max / 1000.0,
                           "l2": "one wave reads %.1f GB of frames, larger than L2; no flush" % (B * in_bytes / 1e9)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * in_bytes), "d2h_bytes_per_step": int(B * out_bytes),
                        "steps": e2e_steps, "timed_region_s": e2e_ms_max / 1000.0, "matches_device_path": same,
                        "api": "pano_process_batch (pinned host frames in, panoramas back to pinned host memory)",
                        "pcie_GBps_used": {"h2d": e2e_value * in_bytes / 1e9, "d2h": e2e_value * out_bytes / 1e9},
                        "pcie_ceiling_GBps": {"what": "raw cudaMemcpyAsync of the same pinned buffers, both directions busy, all ranks at once",
                                              "h2d_aggregate": pcie_sum[0], "d2h_aggregate": pcie_sum[1],
                                              "h2d_min_rank": pcie_min[0], "d2h_min_rank": pcie_min[1]},
                        "pcie_ceiling_panoramas_per_s": min(pcie_sum[0] * 1e9 / in_bytes, pcie_sum[1] * 1e9 / out_bytes),
                        "numa": numa},
                "gpu_launches": launches_per_wave * WAVES * args.steps, "gpu_launches_per_wave": launches_per_wave,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity, "latency": latency, "also": also or None}
        if strip is not None:
            line["strip_split"] = strip
        if variant is not None:
            line["variant"] = variant
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
