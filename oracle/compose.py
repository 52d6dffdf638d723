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
