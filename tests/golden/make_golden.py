"""Generates tests/golden/*.npz from cv2 4.13 (the reference's own cv::detail classes driven by
oracle/cv2_reference.py, which replays ocvStitcher::initSeam/process call for call).

Run in the build container:  python tests/golden/make_golden.py
Inputs are synthetic (tests/util.synth_frame, integer-only) so the GPU box can regenerate them
without cv2; only OUTPUTS of the cv2 path and the init-time tables (masks, weight pyramids) are
stored.  The reference itself holds no golden vectors.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cv2_reference as ref  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from golden import calib  # noqa: E402
import util  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def weight_pyramids(t, num_bands):
    nb, pwh = orc.mb_prepare(t.dst_roi, num_bands)
    out = []
    for i, m in enumerate(t.blend_masks):
        _, bd = orc.mb_feed_rect(t.dst_roi, pwh, nb, t.corners[i], t.sizes[i])
        w = cv2.copyMakeBorder(m.astype(np.float32) * np.float32(1.0 / 255.0), bd[0], bd[1], bd[2], bd[3], cv2.BORDER_CONSTANT)
        lv = [w]
        for _ in range(nb):
            lv.append(cv2.pyrDown(lv[-1]))
        out.append(lv)
    return out


def compose_case(name, rig_name, width, height, warp, num_bands, seed, extras=True):
    Ks, Rs, scale = calib.rig(rig_name, width)
    imgs = util.synth_set(len(Rs), height, width, seed)
    t = ref.init_seam(imgs, Ks, Rs, scale, warp=warp, seam="gc_color", want_gains=True)
    cut = ref.default_cut(t.dst_roi, (t.dst_roi[3] * 7 // 8) // 2 * 2)
    d = dict(width=width, height=height, seed=seed, num_bands=num_bands, scale=np.float32(scale),
             corners=np.array(t.corners, np.int32), sizes=np.array(t.sizes, np.int32),
             dst_roi=np.array(t.dst_roi, np.int32), cut=np.array(cut, np.int32))
    for i, m in enumerate(t.blend_masks):
        d["mask%d" % i] = m
    wp = weight_pyramids(t, num_bands)
    for i, lv in enumerate(wp):
        for l in range(1, len(lv)):
            d["w%d_%d" % (i, l)] = lv[l]
    d["pano_multiband"] = ref.process(t, imgs, "multiband", num_bands=num_bands, cut=cut)
    if not extras:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "dst", t.dst_roi, "cut", cut, sha(d["pano_multiband"])[:12])
        return
    sharp = np.float32(1.0) / np.float32(np.sqrt(np.float32(t.dst_roi[2] * t.dst_roi[3])) * np.float32(5.0) / np.float32(100.0))
    d["sharpness"] = np.float32(sharp)
    fw = ref.feather_weights(t, float(sharp))
    for i, w in enumerate(fw):
        d["fw%d" % i] = w
    d["pano_feather"] = ref.process(t, imgs, "feather", sharpness=float(sharp), cut=cut)
    d["pano_no"] = ref.process(t, imgs, "no", cut=cut)
    # gain apply + feather (BASELINE config 3 order: warp -> apply gain on 8U -> 16S -> feed)
    gm = ref.full_res_gain_maps(t)
    for i, g in enumerate(gm):
        d["gain%d" % i] = g
    d["pano_gain_feather"] = ref.process(t, imgs, "feather", sharpness=float(sharp), cut=cut, apply_gain=True)
    d["pano_gain_multiband"] = ref.process(t, imgs, "multiband", num_bands=num_bands, cut=cut, apply_gain=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "dst", t.dst_roi, "cut", cut, {k: sha(v)[:12] for k, v in d.items() if k.startswith("pano")})


def frontend_case():
    cam = calib.CAM_LIJING_390_FOV60_1920
    # quarter-size camera: scale K and the rect by 1/4 (480x270); input frame is 8UC4
    s = 0.25
    K = np.array(cam["K"], np.float64).reshape(3, 3).copy()
    K[0, 0] *= s; K[0, 2] *= s; K[1, 1] *= s; K[1, 2] *= s
    D = np.array(cam["distorParams"], np.float64)
    size = (480, 270)
    rect = [18, 26, 444, 222]
    newK, mx, my = ref.undistort_tables(K, D, size)
    argb = util.synth_frame(270, 480, 77, channels=4)
    out_same = ref.front_end(argb, size, mx, my, rect, size)
    argb_big = util.synth_frame(540, 960, 78, channels=4)
    out_down = ref.front_end(argb_big, size, mx, my, rect, (360, 203))
    out_noud = ref.front_end(argb_big, size, mx, my, rect, (360, 203), undistort=False)
    ixy, frac = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    np.savez_compressed(os.path.join(HERE, "frontend_small.npz"), K=K, D=D, newK=newK, rect=np.array(rect, np.int32),
                        map_ixy_sha=np.array(sha(ixy)), map_frac_sha=np.array(sha(frac)),
                        out_same=out_same, out_down=out_down, out_noud=out_noud)
    print("frontend", sha(out_same)[:12], sha(out_down)[:12], sha(out_noud)[:12])


def primitive_hashes():
    """Known-answer hashes of cv2 primitives on integer-synthetic inputs."""
    rng = np.random.default_rng(5)
    src = util.synth_frame(97, 131, 9)
    xm = (rng.integers(-800, 5000, (80, 120)) / 32.0).astype(np.float32)
    ym = (rng.integers(-800, 3800, (80, 120)) / 32.0).astype(np.float32)
    xm[0, :6] = -1; ym[0, :6] = -1
    s16 = rng.integers(-700, 700, (66, 96, 3)).astype(np.int16)
    h = dict(
        remap_linear_reflect=sha(cv2.remap(src, xm, ym, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)),
        remap_linear_const=sha(cv2.remap(src, xm, ym, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)),
        remap_nearest=sha(cv2.remap(src[:, :, 0].copy(), xm, ym, cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT)),
        remap_cubic=sha(cv2.remap(src, xm, ym, cv2.INTER_CUBIC)),
        resize_down=sha(cv2.resize(src, (64, 40))),
        resize_up=sha(cv2.resize(src, (300, 211))),
        pyrdown=sha(cv2.pyrDown(s16)),
        pyrup=sha(cv2.pyrUp(s16)),
        pyrdown_odd=sha(cv2.pyrDown(s16[:65, :95])),
    )
    with open(os.path.join(HERE, "primitive_hashes.txt"), "w") as f:
        for k in sorted(h):
            f.write("%s %s\n" % (k, h[k]))
    print("primitives", len(h))


def widened_case():
    """cv2 outputs for the SURVEY 8(f) rows: YUYV ingest (include/nvcam.hpp:880-886), the updateMask tail
    (include/ocvstitcher.hpp:1251-1257) and the two-ring epilogue (src/master.cpp:321-326, src/panocamimpl.cpp:354-360)."""
    rng = np.random.default_rng(77)
    d = {}
    yuyv = rng.integers(0, 256, (36, 64, 2), np.uint8)
    yuyv[0, :6] = [[0, 0], [0, 0], [255, 255], [255, 255], [0, 255], [255, 0]]
    d["yuyv"] = yuyv
    d["yuyv_bgra"] = cv2.cvtColor(yuyv, cv2.COLOR_YUV2BGRA_YUYV)
    low = (rng.integers(0, 2, (23, 41)) * 255).astype(np.uint8)
    full = ((rng.integers(0, 9, (131, 240)) > 0) * 255).astype(np.uint8)
    d["seam_low"], d["seam_full"] = low, full
    d["seam_mask"] = cv2.bitwise_and(cv2.resize(cv2.dilate(low, None), (240, 131), interpolation=cv2.INTER_LINEAR_EXACT), full)
    up = rng.integers(0, 256, (57, 333, 3), np.uint8)
    down = rng.integers(0, 256, (64, 301, 3), np.uint8)
    d["ring_up"], d["ring_down"] = up, down
    ret = cv2.vconcat([cv2.resize(up, (301, 64)), down])
    cv2.rectangle(ret, (0, ret.shape[0] // 2 - 5, ret.shape[1], 10), (0, 0, 0), -1)
    d["ring_resize"] = ret
    fc, width, height = 3, 301, 57 - 6
    ret = cv2.vconcat([np.ascontiguousarray(up[fc:fc + height, :width]), np.ascontiguousarray(down[fc:fc + height, :width])])
    cv2.rectangle(ret, (0, height - 2, width, 4), (0, 0, 0), -1, 1, 0)
    d["ring_crop"] = ret
    np.savez_compressed(os.path.join(HERE, "widened_small.npz"), **d)
    print("widened", {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    cv2.setNumThreads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "widened":
        widened_case()
        sys.exit(0)
    compose_case("cfg1_small", "2222", 240, 135, "spherical", 5, seed=1)
    compose_case("cfg1_cyl_small", "2222", 240, 135, "cylindrical", 3, seed=2, extras=False)
    frontend_case()
    primitive_hashes()
    widened_case()
