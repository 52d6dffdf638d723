"""Pins the FULL-SIZE parity cases on the reference's own frames (2222/1-4.png, 1920x1080):

  config 1  2222 calibration x4, spherical, GraphCut seam masks, 5-band MultiBandBlender, cut [0,64,5336,896]
            (include/ocvstitcher.hpp:975-1216; BASELINE.json configs[0])
  config 2  the same behind nvCam's undistort + crop + resize front end (include/nvcam.hpp:898-929; configs[1])
  config 3  cfg/424camcfg/cameraparaout_1.txt x3, BlocksGainCompensator gains from frame-set 0, FeatherBlender
            (src/stitching_detailed.cpp:829-871; configs[2])

Run in the build container (needs cv2 and /root/reference):   python tests/golden/make_fullsize.py
Writes tests/golden/frames2222/{1..4}.png (byte copies of the reference's sample frames -- data fixtures, the GPU box
has no /root/reference) and tests/golden/fullsize.npz:
  sha256 of cv2's panoramas (the pin that survives a cv2-less box together with the oracle's sha below), the seam
  finder's LOW-RESOLUTION masks (a few KB; the library rebuilds m_blenderMask from them bit-exactly on the device),
  the block gain maps of config 3, and for the record the oracle-vs-cv2 difference at full size.
"""
import hashlib
import os
import shutil
import sys
import time

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import compose, cv2_reference as ref, oracle as orc  # noqa: E402
from golden import calib  # noqa: E402
import util  # noqa: E402

REF_FRAMES = "/root/reference/2222"
W, H, NB, CUT = 1920, 1080, 5, (0, 64, 5336, 896)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def low_res_seam_masks(imgs, Ks, Rs, scale, warp="spherical"):
    """masks_warped after seam_finder->find (include/ocvstitcher.hpp:1033-1035): what pano_set_seam_mask takes."""
    import math
    n = len(imgs)
    swa = min(1.0, math.sqrt(1e5 / (H * W)))
    sw = cv2.PyRotationWarper(warp, np.float32(scale * swa))
    corners, iw, mw = [], [], []
    for i in range(n):
        K = Ks[i].astype(np.float32).copy()
        f = np.float32(swa)
        K[0, 0] *= f; K[0, 2] *= f; K[1, 1] *= f; K[1, 2] *= f
        small = cv2.resize(imgs[i], None, fx=swa, fy=swa, interpolation=cv2.INTER_LINEAR_EXACT)
        c, w = sw.warp(small, K, Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, m = sw.warp(np.full(small.shape[:2], 255, np.uint8), K, Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners.append(c); iw.append(w); mw.append(m)
    finder = cv2.detail_GraphCutSeamFinder("COST_COLOR")
    um = finder.find([a.astype(np.float32) for a in iw], corners, [cv2.UMat(m) for m in mw])
    return [m.get() for m in um]


def oracle_tables(t_cv, Ks, Rs, scale):
    """The C oracle's tables carrying cv2's seam masks (geometry must agree with cv2's)."""
    t = compose.build_tables(Ks, Rs, scale, (W, H), "spherical")
    assert [tuple(c) for c in t.corners] == [tuple(c) for c in t_cv.corners] and [tuple(s) for s in t.sizes] == [tuple(s) for s in t_cv.sizes]
    t.blend_masks = [m.copy() for m in t_cv.blend_masks]
    return t


def diff(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int(d.max()), int(np.count_nonzero(d))


def main():
    os.makedirs(os.path.join(HERE, "frames2222"), exist_ok=True)
    frames = []
    for i in range(1, 5):
        dst = os.path.join(HERE, "frames2222", "%d.png" % i)
        if not os.path.exists(dst):
            shutil.copyfile(os.path.join(REF_FRAMES, "%d.png" % i), dst)
        frames.append(cv2.imread(dst, cv2.IMREAD_COLOR))
        assert frames[-1].shape == (H, W, 3)
    out = {"frames_sha": np.array([sha(f) for f in frames])}

    # ---- config 1
    Ks, Rs, scale = calib.rig("2222", W)
    t0 = time.time()
    t1 = ref.init_seam(frames, Ks, Rs, scale, warp="spherical", seam="gc_color")
    p1 = ref.process(t1, frames, "multiband", NB, cut=CUT)
    for i, m in enumerate(low_res_seam_masks(frames, Ks, Rs, scale)):
        out["c1_seam%d" % i] = m
    for i, m in enumerate(t1.blend_masks):
        out["c1_mask_sha%d" % i] = np.array(sha(m))
    out["c1_sha_cv2"] = np.array(sha(p1))
    to = oracle_tables(t1, Ks, Rs, scale)
    o1 = compose.process(to, frames, "multiband", NB, cut=CUT)        # library-built (host order) float weight pyramids
    out["c1_sha_oracle"] = np.array(sha(o1))
    out["c1_oracle_vs_cv2"] = np.array(diff(o1, p1))
    print("config1", p1.shape, "cv2", sha(p1)[:12], "oracle", sha(o1)[:12], "oracle-vs-cv2 (max, count)", diff(o1, p1), "%.1fs" % (time.time() - t0))

    # ---- config 2: the front end in front (imx390 lijing fov60 1920 entry, cfg/cameras.yaml:80-88)
    cam = calib.CAM_LIJING_390_FOV60_1920
    newK, mx, my = ref.undistort_tables(cam["K"], cam["distorParams"], (W, H))
    out["c2_newK"] = np.asarray(newK, np.float64)
    bgra = [np.dstack([f, np.full((H, W), 255, np.uint8)]) for f in frames]
    fe = [ref.front_end(a, (W, H), mx, my, cam["rect"], (W, H)) for a in bgra]
    out["c2_fe_sha"] = np.array([sha(f) for f in fe])
    t2 = ref.init_seam(fe, Ks, Rs, scale, warp="spherical", seam="gc_color")
    p2 = ref.process(t2, fe, "multiband", NB, cut=CUT)
    for i, m in enumerate(low_res_seam_masks(fe, Ks, Rs, scale)):
        out["c2_seam%d" % i] = m
    out["c2_sha_cv2"] = np.array(sha(p2))
    to2 = oracle_tables(t2, Ks, Rs, scale)
    o2 = compose.process(to2, fe, "multiband", NB, cut=CUT)
    out["c2_sha_oracle"] = np.array(sha(o2))
    out["c2_oracle_vs_cv2"] = np.array(diff(o2, p2))
    print("config2", p2.shape, "cv2", sha(p2)[:12], "oracle", sha(o2)[:12], "oracle-vs-cv2", diff(o2, p2))

    # ---- config 3: imx424 rig x3, block gains + feather, whole dst roi
    Ks3, Rs3, scale3 = calib.rig("424", W)
    t3 = ref.init_seam(frames, Ks3, Rs3, scale3, warp="spherical", seam="gc_color", want_gains=True)
    bw = np.float32(np.sqrt(np.float32(t3.dst_roi[2] * t3.dst_roi[3]))) * np.float32(5.0) / np.float32(100.0)
    sharp = float(np.float32(1.0) / bw)
    p3 = ref.process(t3, frames, "feather", sharpness=sharp, apply_gain=True)
    for i, m in enumerate(low_res_seam_masks(frames, Ks3, Rs3, scale3)):
        out["c3_seam%d" % i] = m
    for i, g in enumerate(t3.gains):
        out["c3_gain%d" % i] = np.asarray(g, np.float32)          # block gain maps (35 x 64-ish), resized at init
    out["c3_sharpness"] = np.float32(sharp)
    out["c3_sha_cv2"] = np.array(sha(p3))
    out["c3_dst_roi"] = np.array(t3.dst_roi, np.int32)
    print("config3", p3.shape, "cv2", sha(p3)[:12], "dst", t3.dst_roi)

    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "fullsize.npz"), **out)
    print("wrote fullsize.npz", os.path.getsize(os.path.join(HERE, "fullsize.npz")), "bytes")


if __name__ == "__main__":
    main()
