"""Whole-path oracle: ocvStitcher::process restated on the scalar C primitives
(oracle/pano_oracle.c) -- TEST INFRASTRUCTURE ONLY, never imported by the product.

Follows include/ocvstitcher.hpp:1141-1216 (process) with the static tables of
initSeam (:1054-1110); gain apply / feather order per src/stitching_detailed.cpp:829-871.
No cv2 needed: masks and (optionally) weight pyramids are inputs.
"""
import math

import numpy as np

from . import oracle as orc

KIND = {"spherical": orc.SPHERICAL, "cylindrical": orc.CYLINDRICAL}


class Tables:
    pass


def build_tables(Ks, Rs, scale, src_size, warp="spherical"):
    """Geometry half of initSeam (:1054-1063, :1110): corners, sizes, dst roi, float maps,
    and the warped all-255 mask (:1085)."""
    t = Tables()
    W, H = src_size
    t.warp, t.scale, t.src_size = warp, float(scale), (W, H)
    t.Ks = [np.asarray(k, np.float32).reshape(3, 3) for k in Ks]
    t.Rs = [np.asarray(r, np.float32).reshape(3, 3) for r in Rs]
    t.corners, t.sizes, t.maps, t.warped_masks = [], [], [], []
    full = np.full((H, W), 255, np.uint8)
    for K, R in zip(t.Ks, t.Rs):
        roi, xm, ym = orc.build_maps(KIND[warp], np.float32(scale), K, R, W, H)
        t.corners.append((roi[0], roi[1]))
        t.sizes.append((roi[2], roi[3]))
        t.maps.append((xm, ym))
        t.warped_masks.append(orc.remap_nearest_u8(full, xm, ym))
    t.dst_roi = orc.result_roi(t.corners, t.sizes)
    t.blend_masks = [m.copy() for m in t.warped_masks]
    t.gain_maps = None
    return t


def blend_width(dst_roi, strength):
    return float(np.float32(math.sqrt(np.float32(dst_roi[2] * dst_roi[3]))) * np.float32(strength) / np.float32(100.0))


def process(t, imgs, blender="multiband", num_bands=5, feather_weights=None, cut=None,
            ext_weights=None, return_s16=False):
    """-> 8-bit panorama (cropped to `cut` = x,y,w,h if given)."""
    warped = []
    for i, im in enumerate(imgs):
        xm, ym = t.maps[i]
        w8 = orc.remap_bilinear_u8(np.ascontiguousarray(im, np.uint8), xm, ym, "reflect")   # :1171
        if t.gain_maps is not None and t.gain_maps[i] is not None:                           # stitching_detailed.cpp:841
            g = t.gain_maps[i]
            w8 = orc.gain_apply_u8(w8, None, float(g)) if np.ndim(g) == 0 else orc.gain_apply_u8(w8, g)
        warped.append(w8.astype(np.int16))                                                  # :1180
    if blender == "multiband":
        res, mask = orc.multiband_blend(warped, t.blend_masks, t.corners, t.sizes, num_bands, ext_weights)
    elif blender == "feather":
        res, mask = orc.feather_blend(warped, feather_weights, t.corners, t.sizes)
    else:
        res, mask = orc.no_blend(warped, t.blend_masks, t.corners, t.sizes)
    if return_s16:
        return res, mask
    if cut is None:
        cut = (0, 0, res.shape[1], res.shape[0])
    return orc.s16_to_u8_crop(res, cut)                                                     # :1208-1210


def front_end(argb, undist_size, mapx, mapy, rect, out_size, undistort=True):
    """nvCam::read_frame pixel pipeline + getFrame resize (include/nvcam.hpp:898-929,1092-1094)."""
    argb = np.ascontiguousarray(argb, np.uint8)
    if undistort:
        tmp = orc.resize_bilinear_u8(argb, undist_size)[:, :, :3]       # resize + cvtColor RGBA2RGB
        und = orc.remap_cubic_u8(np.ascontiguousarray(tmp), mapx, mapy)
        und = np.ascontiguousarray(und[rect[1]:rect[1] + rect[3], rect[0]:rect[0] + rect[2]])
        ret = orc.resize_bilinear_u8(und, undist_size)
    else:
        ret = orc.resize_bilinear_u8(np.ascontiguousarray(argb[:, :, :3]), undist_size)
    return orc.resize_bilinear_u8(ret, out_size)


def fused_front_end_maps(xm, ym, out_size, undist_size, mapx, mapy, rect, cam_size, undistort=True):
    """Composed backward map of the single-gather ("fused map") variant: rotation-warp map (into the
    stitcher input, `out_size`) -> inverse of every front_end() stage -> position in the camera frame.
    The reference has no such variant (it always resamples sequentially); this restates the
    definition given in include/panob200.h (PANO_FRONTEND_FUSED) in float64 numpy so the tests can
    check the library's composed maps independently.  Returns float64 (x, y)."""
    def foldc(c, n):      # BORDER_REFLECT of the warp in continuous form, clamped to the pixel centres
        t = np.mod(c + 0.5, 2.0 * n)
        t = np.where(t >= n, 2.0 * n - t, t)
        return np.clip(t - 0.5, 0.0, n - 1.0)

    def inv(d, ssize, dsize):   # cv::resize INTER_LINEAR source coordinate, clamped like its tables
        return np.clip((d + 0.5) * (1.0 / (dsize / ssize)) - 0.5, 0.0, ssize - 1.0)

    x = foldc(np.asarray(xm, np.float64), out_size[0])
    y = foldc(np.asarray(ym, np.float64), out_size[1])
    uw, uh = undist_size
    if tuple(out_size) != (uw, uh):
        x, y = inv(x, uw, out_size[0]), inv(y, uh, out_size[1])
    if undistort:
        if (rect[2], rect[3]) != (uw, uh):
            x, y = inv(x, rect[2], uw), inv(y, rect[3], uh)
        x = np.clip(x + rect[0], 0.0, uw - 1.0)
        y = np.clip(y + rect[1], 0.0, uh - 1.0)
        x0 = np.minimum(uw - 2, x.astype(np.int64)); y0 = np.minimum(uh - 2, y.astype(np.int64))
        fx, fy = x - x0, y - y0
        def lerp(m):
            m = np.asarray(m, np.float64)
            return (1 - fy) * ((1 - fx) * m[y0, x0] + fx * m[y0, x0 + 1]) + fy * ((1 - fx) * m[y0 + 1, x0] + fx * m[y0 + 1, x0 + 1])
        x, y = lerp(mapx), lerp(mapy)
    if tuple(cam_size) != (uw, uh):
        x, y = inv(x, cam_size[0], uw), inv(y, cam_size[1], uh)
    return np.clip(x, 0.0, cam_size[0] - 1.0), np.clip(y, 0.0, cam_size[1] - 1.0)
