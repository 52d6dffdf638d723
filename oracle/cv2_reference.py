"""cv2-driven restatement of the reference's call sequence -- TEST INFRASTRUCTURE ONLY.

The reference (header-only C++ on OpenCV 3 + Jetson libraries) cannot be compiled here
(SURVEY.md 8c), but `cv2` 4.13 binds the very cv::detail classes it calls.  This module
replays the reference's call order through those classes:

  init_seam()  <- ocvStitcher::initSeam      include/ocvstitcher.hpp:975-1139
  process()    <- ocvStitcher::process       include/ocvstitcher.hpp:1141-1216
  feather / gain-apply order                 src/stitching_detailed.cpp:829-871
  front_end()  <- nvCam::read_frame+getFrame include/nvcam.hpp:898-921,1092-1094
  undistort maps <- prepareUndistorMap       include/nvcam.hpp:823-833

It is used (a) to pin oracle/pano_oracle.c, (b) to generate tests/golden/, (c) as the
CPU baseline (`bench.py --impl reference` and the `cpu_baseline` leg).  The product never
imports it.
"""
import math

import cv2
import numpy as np


def parse_calibration_old(path, block=-1, scale_to_width=None):
    """2222/cameraparaout_{1,2}.txt layout: 'ts:' / K (9) / N x R (9) / scale."""
    lines = [l.strip() for l in open(path).read().splitlines() if l.strip()]
    starts = [i for i, l in enumerate(lines) if l.endswith(":")]
    s = starts[block]
    e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(lines)
    body = lines[s + 1:e]
    K = np.array([float(v) for v in body[0].split(",") if v], np.float32).reshape(3, 3)
    Rs = [np.array([float(v) for v in l.split(",") if v], np.float32).reshape(3, 3) for l in body[1:-1]]
    scale = np.float32(float(body[-1]))
    if scale_to_width is not None:
        f = np.float32(scale_to_width / (2.0 * K[0, 2]))
        K = K.copy()
        K[0, 0] *= f; K[0, 2] *= f; K[1, 1] *= f; K[1, 2] *= f
        scale = np.float32(scale * f)
    return [K.copy() for _ in Rs], Rs, float(scale)


def parse_calibration_new(path, block=-1):
    """cfg/*camcfg/cameraparaout_*.txt layout: 'ts:' / N lines of 18 floats / scale
    (written by saveCameraParams, include/ocvstitcher.hpp:537-559)."""
    lines = [l.strip() for l in open(path).read().splitlines() if l.strip()]
    starts = [i for i, l in enumerate(lines) if l.endswith(":")]
    s = starts[block]
    e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(lines)
    body = lines[s + 1:e]
    Ks, Rs = [], []
    for l in body[:-1]:
        v = [float(t) for t in l.split(",") if t]
        Ks.append(np.array(v[:9], np.float32).reshape(3, 3))
        Rs.append(np.array(v[9:18], np.float32).reshape(3, 3))
    return Ks, Rs, float(np.float32(float(body[-1])))


def num_bands_from_strength(dst_w, dst_h, blend_strength):
    """include/ocvstitcher.hpp:1188-1195.  Returns None when Blender::NO is selected."""
    blend_width = np.float32(math.sqrt(np.float32(dst_w * dst_h))) * np.float32(blend_strength) / np.float32(100.0)
    if blend_width < 1.0:
        return None
    return int(math.ceil(math.log(float(blend_width)) / math.log(2.0)) - 1.0)


class StitchTables:
    """Static products of init: corners, sizes, dst roi, blend masks, ..."""
    pass


def init_seam(imgs, Ks, Rs, scale, warp="spherical", seam="gc_color", want_gains=False):
    """ocvStitcher::initSeam restated (include/ocvstitcher.hpp:975-1139)."""
    n = len(imgs)
    H, W = imgs[0].shape[:2]
    swa = min(1.0, math.sqrt(1e5 / (H * W)))
    seam_imgs = [cv2.resize(im, None, fx=swa, fy=swa, interpolation=cv2.INTER_LINEAR_EXACT) for im in imgs]
    sw = cv2.PyRotationWarper(warp, np.float32(scale * swa))
    corners, images_warped, masks_warped = [], [], []
    for i in range(n):
        K = Ks[i].astype(np.float32).copy()
        f = np.float32(swa)
        K[0, 0] *= f; K[0, 2] *= f; K[1, 1] *= f; K[1, 2] *= f
        c, iw = sw.warp(seam_imgs[i], K, Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        corners.append(c)
        images_warped.append(iw)
        m = np.full(seam_imgs[i].shape[:2], 255, np.uint8)
        _, mw = sw.warp(m, K, Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        masks_warped.append(mw)
    t = StitchTables()
    t.gains = None
    if want_gains:
        comp = cv2.detail_BlocksGainCompensator(32, 32, 1)
        comp.feed(corners=corners, images=[cv2.UMat(a) for a in images_warped],
                  masks=[cv2.UMat(a) for a in masks_warped])
        t.gains = [np.asarray(g.get() if hasattr(g, "get") else g) for g in comp.getMatGains()]
    if seam == "gc_color":
        finder = cv2.detail_GraphCutSeamFinder("COST_COLOR")
        um = [cv2.UMat(m) for m in masks_warped]
        um = finder.find([a.astype(np.float32) for a in images_warped], corners, um)
        masks_warped = [m.get() for m in um]
    elif seam != "no":
        raise ValueError(seam)
    bw = cv2.PyRotationWarper(warp, np.float32(scale))
    t.corners, t.sizes, t.blend_masks, t.warped_masks = [], [], [], []
    for i in range(n):
        roi = bw.warpRoi((W, H), Ks[i], Rs[i])
        t.corners.append((roi[0], roi[1]))
        t.sizes.append((roi[2], roi[3]))
    for i in range(n):
        full = np.full((H, W), 255, np.uint8)
        _, mw = bw.warp(full, Ks[i], Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        dil = cv2.dilate(masks_warped[i], None)
        seam_mask = cv2.resize(dil, (mw.shape[1], mw.shape[0]), interpolation=cv2.INTER_LINEAR_EXACT)
        t.warped_masks.append(mw)
        t.blend_masks.append(cv2.bitwise_and(seam_mask, mw))
    t.dst_roi = cv2.detail.resultRoi(corners=t.corners, sizes=t.sizes)
    t.warp, t.scale, t.Ks, t.Rs, t.src_size = warp, float(scale), Ks, Rs, (W, H)
    return t


def default_cut(dst_roi, cut_h):
    """Non-default-init crop rule, include/ocvstitcher.hpp:959-964."""
    return (0, (dst_roi[3] - cut_h) // 2, dst_roi[2], cut_h)


def build_warp_maps(t):
    """Maps as RotationWarperBase::buildMaps builds them (cached variant)."""
    bw = cv2.PyRotationWarper(t.warp, np.float32(t.scale))
    return [bw.buildMaps(t.src_size, t.Ks[i], t.Rs[i])[1:] for i in range(len(t.Ks))]


def full_res_gain_maps(t):
    """BlocksGainCompensator::apply resizes the block gain map to the image size
    (INTER_LINEAR) on every call; it is static, so build it once."""
    return [cv2.resize(g, t.sizes[i], interpolation=cv2.INTER_LINEAR) for i, g in enumerate(t.gains)]


def feather_weights(t, sharpness):
    """FeatherBlender::feed weight (static): min(distanceTransform(mask, L1, 3)*sharpness, 1)."""
    return [_feather_weight(m, sharpness) for m in t.blend_masks]


def _feather_weight(mask, sharpness):
    w = cv2.distanceTransform(mask, cv2.DIST_L1, 3)
    _, w = cv2.threshold(w * np.float32(sharpness), 1.0, 1.0, cv2.THRESH_TRUNC)
    return w


def process(t, imgs, blender="multiband", num_bands=5, sharpness=0.02, cut=None, maps=None,
            apply_gain=False, return_s16=False):
    """ocvStitcher::process restated (include/ocvstitcher.hpp:1141-1216); with
    apply_gain / feather it follows src/stitching_detailed.cpp:829-871.
    maps=None -> 'faithful' (PyRotationWarper.warp rebuilds maps per call like :1171);
    maps=build_warp_maps(t) -> 'cached-maps' variant (identical pixels)."""
    n = len(imgs)
    if blender == "multiband":
        bl = cv2.detail_MultiBandBlender(0, num_bands)
    elif blender == "feather":
        bl = cv2.detail_FeatherBlender(sharpness)
    else:
        bl = cv2.detail.Blender_createDefault(cv2.detail.Blender_NO)
    bl.prepare(t.dst_roi)
    bw = None if maps is not None else cv2.PyRotationWarper(t.warp, np.float32(t.scale))
    comp = None
    if apply_gain:
        comp = cv2.detail_BlocksGainCompensator(32, 32, 1)
        comp.setMatGains(t.gains)
    for i in range(n):
        if maps is not None:
            iw = cv2.remap(imgs[i], maps[i][0], maps[i][1], cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        else:
            _, iw = bw.warp(imgs[i], t.Ks[i], t.Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        if comp is not None:
            iw = comp.apply(i, t.corners[i], iw, t.warped_masks[i])
        bl.feed(iw.astype(np.int16), t.blend_masks[i], t.corners[i])
    res, res_mask = bl.blend(None, None)
    if return_s16:
        return res, res_mask
    out = np.clip(res, 0, 255).astype(np.uint8)   # == convertTo(CV_8U) on CV_16S
    if cut is not None:
        out = out[cut[1]:cut[1] + cut[3], cut[0]:cut[0] + cut[2]]
    return out


# ----------------------------------------------------------------- front end (nvCam)

def undistort_tables(K, D, size):
    """prepareUndistorMap, include/nvcam.hpp:823-833."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    D = np.asarray(D, np.float64).reshape(-1)
    newK, _ = cv2.getOptimalNewCameraMatrix(K, D, size, 1, size, 0)
    mx, my = cv2.initUndistortRectifyMap(K, D, np.eye(3, dtype=np.float32), newK, size, cv2.CV_32FC1)
    return newK, mx, my


def front_end(argb, undist_size, mapx, mapy, rect, out_size, undistort=True):
    """nvCam::read_frame pixel pipeline (include/nvcam.hpp:898-929) followed by the
    getFrame(.., src=false) resize (:1092-1094).  argb: HxWx4 u8."""
    if undistort:
        tmp = cv2.resize(argb, undist_size)
        tmp = cv2.cvtColor(tmp, cv2.COLOR_RGBA2RGB)
        und = cv2.remap(tmp, mapx, mapy, cv2.INTER_CUBIC)
        und = und[rect[1]:rect[1] + rect[3], rect[0]:rect[0] + rect[2]]
        ret = cv2.resize(und, undist_size)
    else:
        tmp = cv2.cvtColor(argb, cv2.COLOR_RGBA2RGB)
        ret = cv2.resize(tmp, undist_size)
    return cv2.resize(ret, out_size)
