#!/usr/bin/env python
"""panoramas/sec of the per-frame compose path (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU OpenCV path

A "step" = one pass of the hot path over one batch of `--batch` synthetic frame-sets per GPU
(4 x 1920x1080 BGR each; BASELINE config 1: spherical warp + 5-band MultiBandBlender, cut
5336x896).  Frame-sets are sharded across ranks with no data-path collective (weak scaling:
per-GPU work is fixed).  `value` = whole-job frame-sets/s with inputs resident in HBM, timed on
the device (CUDA events, barrier + synchronize on both sides, max over ranks).  `e2e` = the same
metric through the host-buffer API (pano_process_batch: pinned host frames in, panoramas back to
pinned host memory, H2D/D2H inside the timed region).  One batch is 1.6 GB of frames, far larger
than the 126 MB L2, so no L2 flush is needed between iterations.
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
WORKLOAD = WORKLOADS["config1"]
# what ncu shows as the on-chip limiter of each hot kernel (profiles/r1p_ncu_summary.txt): the HBM
# roofline is the denominator the task asks for, but none of the exact-arithmetic gathers is DRAM-limited
LIMITERS = {
    "fe_cubic_undistort": "L1 data pipes: LSU wavefronts 84 % + TEX 81 % (16 smem taps + 32-byte weight entry per pixel); DRAM traffic 24 % of the copy rate",
    "fe_resize": "issue slots 76 % (exact cv::resize fixed point, ~60 instructions per pixel); DRAM traffic 40 % of the copy rate",
    "warp": "L1 LSU wavefronts 80 % + issue 72 %; DRAM traffic 56 % of the copy rate",
    "pyrdown_l0": "DRAM traffic 72 % of the copy rate at 28 % occupancy (94 registers)",
    "collapse_l0": "issue slots 61 %; DRAM traffic 48 % of the copy rate",
    "fe_yuyv_to_bgra": "HBM (97 % of the measured copy rate)",
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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


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


def cpu_reference_time(kind, t, frame_sets, repeats, front=False, workload="config1"):
    """Times the reference's per-frame path (ocvStitcher::process restated call for call; with
    front=True preceded by nvCam's resize/undistort/crop/resize per camera) on the host cores.
    -> (panoramas/s, cores used, description)."""
    n = 0
    if kind == "cv2":
        from oracle import cv2_reference as ref
        fe = None
        if front:
            cam = camera_entry()
            _, mx, my = ref.undistort_tables(cam["K"], cam["distorParams"], (W, H))
            fe = lambda a: ref.front_end(a, (W, H), mx, my, cam["rect"], (W, H))     # noqa: E731

        def to_bgra(f):
            if f.shape[2] == 2:        # YUYVCAM ingest (include/nvcam.hpp:880-886)
                return __import__("cv2").cvtColor(np.ascontiguousarray(f), __import__("cv2").COLOR_YUV2BGRA_YUYV)
            return np.dstack([f, np.full(f.shape[:2], 255, np.uint8)]) if f.shape[2] == 3 else f

        def one(fs):
            if fe is not None:
                fs = [fe(to_bgra(f)) for f in fs]
            if workload == "config3":    # warp -> compensator->apply -> FeatherBlender (src/stitching_detailed.cpp:829-871)
                return ref.process(t, fs, "feather", sharpness=config3_sharpness(t.dst_roi), apply_gain=True)
            return ref.process(t, fs, "multiband", NBANDS, cut=CUT)   # 'faithful': maps rebuilt per call (:1171)
        one(frame_sets[0])          # warm-up
        t0 = time.perf_counter()
        for r in range(repeats):
            one(frame_sets[r % len(frame_sets)])
            n += 1
        dt = time.perf_counter() - t0
        return n / dt, os.cpu_count() or 1, ("%d panoramas, cv2 %s cv::detail classes driven by the restated ocvStitcher::process "
                                             "call sequence (warper->warp rebuilds maps per call as the reference does), "
                                             "cv2.setNumThreads(%d)" % (n, __import__("cv2").__version__, os.cpu_count() or 1))
    from oracle import compose
    t0 = time.perf_counter()
    for r in range(repeats):
        compose.process(t, frame_sets[r % len(frame_sets)], "multiband", NBANDS, cut=CUT)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, 1, "%d panoramas, scalar C port (oracle/pano_oracle.c), cached maps, 1 thread" % n


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    per_step = 2
    sets = [[np.ascontiguousarray(f) for f in synth_numpy(1000 + s)] for s in range(2)]
    kind, t = cpu_reference_setup(sets[0], args.workload)
    if args.workload == "config2-yuyv":
        def yuyv(f):
            y = np.empty(f.shape[:2] + (2,), np.uint8)
            y[:, :, 0] = f[:, :, 0]; y[:, 0::2, 1] = f[:, 0::2, 1]; y[:, 1::2, 1] = f[:, 0::2, 2]
            return y
        sets = [[yuyv(f) for f in s] for s in sets]
    front = args.workload.startswith("config2")
    cpu_reference_time(kind, t, sets, max(1, args.warmup), front, args.workload)
    t0 = time.perf_counter()
    v, cores, desc = cpu_reference_time(kind, t, sets, per_step * args.steps, front, args.workload)
    dt = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/s16 fixed-point + f32 weights", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "frame_sets_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def synth_numpy(seed):
    import util
    return util.synth_set(NCAM, H, W, seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frame-sets per GPU per step")
    ap.add_argument("--max-batch", type=int, default=64, help="frame-sets per launch wave (workspace: 73 MB per slot)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    import panob200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- init (one-time, host): calibration -> initSeam -> tables uploaded once ----
    Ks, Rs, scale = calibration(args.workload)
    B = args.batch
    frames = synth_batch_torch(B, 1234 + rank, dev)
    front = None
    if args.workload.startswith("config2"):
        # camera frames are 8UC4 (the VIC's ARGB output, include/nvcam.hpp:889-893): BGR + alpha 255
        if args.workload == "config2-yuyv":
            # valid YUYV bytes from the synthetic frames: Y = channel 0, U / V = channels 1 / 2 of the even pixel
            yuyv = torch.empty(frames.shape[:-1] + (2,), dtype=torch.uint8, device=dev)
            yuyv[..., 0] = frames[..., 0]
            yuyv[..., 0::2, 1] = frames[..., 0::2, 1]
            yuyv[..., 1::2, 1] = frames[..., 0::2, 2]
            frames = yuyv.contiguous()
            del yuyv
        else:
            frames = torch.cat([frames, torch.full(frames.shape[:-1] + (1,), 255, dtype=torch.uint8, device=dev)], dim=-1).contiguous()
        front = make_front_end(local_rank, args.max_batch * NCAM, "yuyv" if args.workload == "config2-yuyv" else "bgra")
    if front is None:
        set0 = [frames[0, i].cpu().numpy() for i in range(NCAM)]
    else:   # the stitcher calibrates on what the front end delivers
        set0 = [front.getFrame(frames[0, i].cpu().numpy()) for i in range(NCAM)]
    masks_how = "GraphCut seam masks (cv2, host init)"
    if args.workload == "config3":
        # init like src/stitching_detailed.cpp: seam masks + block gains from frame-set 0 (host, cv2), uploaded once
        from oracle import cv2_reference as ref
        t3 = ref.init_seam(set0, Ks, Rs, scale, warp="spherical", seam="gc_color", want_gains=True)
        sharp = config3_sharpness(t3.dst_roi)
        cfg = panob200.StitcherConfig(width=W, height=H, num_images=NCAM, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                      blender="feather", num_bands=0, sharpness=sharp, device=local_rank, max_batch=args.max_batch)
        st = panob200.ocvStitcher(cfg)
        rc = st.initTables(t3.blend_masks, None, ref.feather_weights(t3, sharp))
        if rc == 0:
            st.set_gain_maps(ref.full_res_gain_maps(t3))
    else:
        cfg = panob200.StitcherConfig(width=W, height=H, num_images=NCAM, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                      blender="multiband", num_bands=NBANDS, cut=CUT, device=local_rank,
                                      max_batch=args.max_batch, initMode=2)
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
    out = torch.empty((B, oh, ow, 3), dtype=torch.uint8, device=dev)
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

    def barrier():
        if world > 1:
            dist.barrier()

    def step():
        st.process_device(frames, out, stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    launches_per_step = st.last_launch_count()
    torch.cuda.synchronize()
    barrier()
    # clocks / throttle reasons are sampled under the same load: the sampler runs while the step loop
    # keeps the GPU busy (a pre-roll of identical steps, then the K timed steps, then a post-roll so
    # that at least a few 150 ms nvidia-smi samples land inside the loaded interval)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.4:
        step()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    t_post = time.perf_counter()
    while time.perf_counter() - t_post < 0.4:
        step()
        torch.cuda.synchronize()
    clocks = sampler.summary()

    # ---- per-kernel device time (CUDA events on the launching stream) for the roofline ----
    st.enable_profile(True)
    prof_acc = {}
    for _ in range(3):
        step()
        torch.cuda.synchronize()
        for p in st.read_profile():
            a = prof_acc.setdefault(p["name"], dict(ms=0.0, launches=0, bytes=0.0))
            a["ms"] += p["ms"]; a["launches"] += p["launches"]; a["bytes"] += p["alg_bytes"]
    st.enable_profile(False)

    # ---- end to end through the host-buffer API ----
    host_in = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
    host_in.copy_(frames)
    host_out = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 5))
    st.process_batch(host_in, host_out)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        st.process_batch(host_in, host_out)     # synchronous: returns when the panoramas are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    same = bool(torch.equal(host_out[:2].to(dev), out[:2]))

    tms = torch.tensor([ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(tms[0]), float(tms[1])

    if rank == 0:
        peak, peak_src = peaks()
        value = world * B * args.steps / (ms_max / 1000.0)
        e2e_value = world * B * e2e_steps / (e2e_ms_max / 1000.0)
        total_prof = sum(a["ms"] for a in prof_acc.values()) or 1.0
        kernels = {k: {"ms_per_launch": a["ms"] / a["launches"], "share": a["ms"] / total_prof,
                       "GBps": a["bytes"] / a["ms"] / 1e6, "frac": a["bytes"] / a["ms"] / 1e6 / peak}
                   for k, a in prof_acc.items()}
        dom = max(prof_acc, key=lambda k: prof_acc[k]["ms"])
        d = prof_acc[dom]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                per_set = json.load(open(tp)).get(dom)      # ncu DRAM bytes per frame-set
                traffic = per_set * min(B, args.max_batch) if per_set is not None else None
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": d["bytes"] / d["ms"] / 1e6, "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": d["bytes"] / d["ms"] / 1e6 / peak, "traffic": traffic,
                "limiter": LIMITERS.get(dom), "alg_bytes_per_launch": d["bytes"] / d["launches"], "ms_per_launch": d["ms"] / d["launches"],
                "whole_path_GBps": sum(a["bytes"] for a in prof_acc.values()) / total_prof / 1e6, "kernels": kernels}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sets = [[frames[b, i].cpu().numpy() for i in range(NCAM)] for b in range(2)]
            kind, t = cpu_reference_setup([f[:, :, :3] for f in sets[0]] if front is None else set0, args.workload)
            v, cores, desc = cpu_reference_time(kind, t, sets, 24 if kind == "cv2" else 4, front is not None, args.workload)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/s16 fixed-point + f32 weights", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload], "frame_sets_per_gpu_per_step": B, "frame_sets_per_wave": min(B, args.max_batch),
                           "masks": masks_how, "l2": "inputs (1.6 GB/step) larger than L2; no flush"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_in.numel()),
                        "d2h_bytes_per_step": int(host_out.numel()), "steps": e2e_steps, "matches_device_path": same},
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
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
