"""ctypes front for oracle/pano_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

Every function is a thin numpy wrapper over the scalar C restatement; see the C file
for the reference file:line each one follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "pano_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


SPHERICAL, CYLINDRICAL = 0, 1


def warp_roi(kind, scale, K, R, W, H):
    """-> (x, y, w, h) == cv2.PyRotationWarper.warpRoi"""
    K, R = _f32(K), _f32(R)
    roi = (C.c_int * 4)()
    lib().orc_warp_roi(C.c_int(kind), C.c_float(scale), _p(K), _p(R), C.c_int(W), C.c_int(H), roi)
    return tuple(roi)


def build_maps(kind, scale, K, R, W, H):
    """-> (roi, xmap, ymap) float32 maps of shape (roi.h, roi.w)"""
    K, R = _f32(K), _f32(R)
    roi = warp_roi(kind, scale, K, R, W, H)
    xm = np.empty((roi[3], roi[2]), np.float32)
    ym = np.empty((roi[3], roi[2]), np.float32)
    lib().orc_build_maps(C.c_int(kind), C.c_float(scale), _p(K), _p(R), C.c_int(W), C.c_int(H), _p(xm), _p(ym))
    return roi, xm, ym


def convert_maps(xmap, ymap):
    xmap, ymap = _f32(xmap), _f32(ymap)
    ixy = np.empty(xmap.shape + (2,), np.int16)
    frac = np.empty(xmap.shape, np.uint16)
    lib().orc_convert_maps(_p(xmap), _p(ymap), C.c_size_t(xmap.size), _p(ixy), _p(frac))
    return ixy, frac


def _img3(src):
    src = np.ascontiguousarray(src)
    if src.ndim == 2:
        return src, 1
    return src, src.shape[2]


def remap_bilinear_u8(src, xmap, ymap, border="reflect"):
    src, ch = _img3(src)
    assert src.dtype == np.uint8
    xmap, ymap = _f32(xmap), _f32(ymap)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) + ((ch,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_remap_bilinear_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(ch),
                                C.c_size_t(src.strides[0]), _p(xmap), _p(ymap), C.c_int(dw), C.c_int(dh),
                                C.c_int(1 if border == "reflect" else 0), _p(dst))
    return dst


def remap_nearest_u8(src, xmap, ymap):
    src, ch = _img3(src)
    xmap, ymap = _f32(xmap), _f32(ymap)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) + ((ch,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_remap_nearest_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(ch),
                               C.c_size_t(src.strides[0]), _p(xmap), _p(ymap), C.c_int(dw), C.c_int(dh), _p(dst))
    return dst


def remap_cubic_u8(src, xmap, ymap):
    src, ch = _img3(src)
    xmap, ymap = _f32(xmap), _f32(ymap)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) + ((ch,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_remap_cubic_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(ch),
                             C.c_size_t(src.strides[0]), _p(xmap), _p(ymap), C.c_int(dw), C.c_int(dh), _p(dst))
    return dst


def cubic_table():
    t = np.empty((1024, 16), np.int16)
    lib().orc_cubic_table(_p(t))
    return t


def resize_bilinear_u8(src, dsize):
    """dsize = (w, h) like cv2.resize"""
    src, ch = _img3(src)
    dw, dh = dsize
    dst = np.empty((dh, dw) + ((ch,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_resize_bilinear_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(ch),
                                 C.c_size_t(src.strides[0]), _p(dst), C.c_int(dw), C.c_int(dh))
    return dst


def resize_bilinear_scaled_u8(src, fx, fy):
    """cv2.resize(src, None, fx=fx, fy=fy) [INTER_LINEAR]"""
    src, ch = _img3(src)
    dw, dh = int(np.rint(src.shape[1] * fx)), int(np.rint(src.shape[0] * fy))
    dst = np.empty((dh, dw) + ((ch,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_resize_bilinear_scaled_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(ch),
                                        C.c_size_t(src.strides[0]), _p(dst), C.c_int(dw), C.c_int(dh), C.c_double(fx), C.c_double(fy))
    return dst


def init_undistort_map(K, D, newK, w, h):
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    D = np.ascontiguousarray(list(D)[:4], np.float64)
    newK = np.ascontiguousarray(newK, np.float64).reshape(9)
    mx = np.empty((h, w), np.float32)
    my = np.empty((h, w), np.float32)
    lib().orc_init_undistort_map(_p(K), _p(D), _p(newK), C.c_int(w), C.c_int(h), _p(mx), _p(my))
    return mx, my


def pyrdown_s16(src):
    src = np.ascontiguousarray(src, np.int16)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dst = np.empty(((h + 1) // 2, (w + 1) // 2) + ((ch,) if src.ndim == 3 else ()), np.int16)
    lib().orc_pyrdown_s16(_p(src), C.c_int(w), C.c_int(h), C.c_int(ch), _p(dst))
    return dst


def pyrdown_f32(src):
    src = _f32(src)
    h, w = src.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyrdown_f32(_p(src), C.c_int(w), C.c_int(h), _p(dst))
    return dst


def pyrup_s16(src, dsize):
    src = np.ascontiguousarray(src, np.int16)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dw, dh = dsize
    dst = np.empty((dh, dw) + ((ch,) if src.ndim == 3 else ()), np.int16)
    lib().orc_pyrup_s16(_p(src), C.c_int(w), C.c_int(h), C.c_int(ch), _p(dst), C.c_int(dw), C.c_int(dh))
    return dst


def result_roi(corners, sizes):
    n = len(corners)
    c = np.ascontiguousarray(corners, np.int32).reshape(n, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(n, 2)
    roi = (C.c_int * 4)()
    lib().orc_result_roi(C.c_int(n), _p(c), _p(s), roi)
    return tuple(roi)


def mb_prepare(roi, num_bands):
    r = (C.c_int * 4)(*roi)
    pwh = (C.c_int * 2)()
    nb = lib().orc_mb_prepare(r, C.c_int(num_bands), pwh)
    return nb, (pwh[0], pwh[1])


def mb_feed_rect(roi, padded_wh, nb, corner, size):
    """-> (rect (x,y,w,h) relative to the padded dst roi, borders (top,bottom,left,right))"""
    r = (C.c_int * 4)(*roi)
    p = (C.c_int * 2)(*padded_wh)
    c = (C.c_int * 2)(*corner)
    s = (C.c_int * 2)(*size)
    rect = (C.c_int * 4)()
    bd = (C.c_int * 4)()
    lib().orc_mb_feed_rect(r, p, C.c_int(nb), c, s, rect, bd)
    return tuple(rect), tuple(bd)


def _ptr_array(arrs, ctype=C.c_void_p):
    arr = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        arr[i] = a.ctypes.data
    return arr


def multiband_blend(imgs, masks, corners, sizes, num_bands, ext_weights=None):
    """imgs: list of int16 HxWx3; masks: list of uint8 HxW; corners/sizes: lists of (x,y)/(w,h).
    ext_weights: optional list (per image) of lists (per level 0..nb) of float32 weight maps.
    -> (result int16 roi_h x roi_w x 3, mask uint8)"""
    n = len(imgs)
    imgs = [np.ascontiguousarray(a, np.int16) for a in imgs]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    c = np.ascontiguousarray(corners, np.int32).reshape(n, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(n, 2)
    roi = result_roi(corners, sizes)
    out = np.empty((roi[3], roi[2], 3), np.int16)
    om = np.empty((roi[3], roi[2]), np.uint8)
    ew = None
    keep = []
    if ext_weights is not None:
        for per_img in ext_weights:
            for w in per_img:
                keep.append(_f32(w))
        ew = _ptr_array(keep)
    lib().orc_multiband_blend(C.c_int(n), _ptr_array(imgs), _ptr_array(masks), _p(c), _p(s),
                              C.c_int(num_bands), ew, _p(out), _p(om))
    return out, om


def feather_blend(imgs, weights, corners, sizes):
    n = len(imgs)
    imgs = [np.ascontiguousarray(a, np.int16) for a in imgs]
    weights = [_f32(a) for a in weights]
    c = np.ascontiguousarray(corners, np.int32).reshape(n, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(n, 2)
    roi = result_roi(corners, sizes)
    out = np.empty((roi[3], roi[2], 3), np.int16)
    om = np.empty((roi[3], roi[2]), np.uint8)
    lib().orc_feather_blend(C.c_int(n), _ptr_array(imgs), _ptr_array(weights), _p(c), _p(s), _p(out), _p(om))
    return out, om


def no_blend(imgs, masks, corners, sizes):
    n = len(imgs)
    imgs = [np.ascontiguousarray(a, np.int16) for a in imgs]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    c = np.ascontiguousarray(corners, np.int32).reshape(n, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(n, 2)
    roi = result_roi(corners, sizes)
    out = np.empty((roi[3], roi[2], 3), np.int16)
    om = np.empty((roi[3], roi[2]), np.uint8)
    lib().orc_no_blend(C.c_int(n), _ptr_array(imgs), _ptr_array(masks), _p(c), _p(s), _p(out), _p(om))
    return out, om


def gain_apply_u8(img, gain=None, g=1.0):
    img = np.array(img, np.uint8, copy=True, order="C")
    ch = 1 if img.ndim == 2 else img.shape[2]
    gm = _f32(gain) if gain is not None else None
    lib().orc_gain_apply_u8(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(ch), _p(gm), C.c_double(g))
    return img


def s16_to_u8_crop(src, cut):
    src = np.ascontiguousarray(src, np.int16)
    ch = src.shape[2]
    cutc = (C.c_int * 4)(*cut)
    dst = np.empty((cut[3], cut[2], ch), np.uint8)
    lib().orc_s16_to_u8_crop(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(ch), cutc, _p(dst))
    return dst
