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


def ring_epilogue(up, down, mode="resize", finalcut=0, bar=None):
    """Caller step after the two process calls: src/master.cpp:321-326 (mode 'resize': cv::resize(up, down.size()),
    vconcat, rectangle(Rect(0, rows/2 - 5, cols, 10), 0, -1)) or src/panocamimpl.cpp:354-360 (mode 'crop': both
    cropped to Rect(0, finalcut, min w, min h - 2*finalcut), vconcat, rectangle(Rect(0, height - 2, width, 4)))."""
    up = np.ascontiguousarray(up, np.uint8); down = np.ascontiguousarray(down, np.uint8)
    if mode == "resize":
        bar = 10 if bar is None else bar
        if up.shape != down.shape:
            up = orc.resize_bilinear_u8(up, (down.shape[1], down.shape[0]))
        ret = np.vstack([up, down])
        y0 = ret.shape[0] // 2 - bar // 2
    else:
        bar = 4 if bar is None else bar
        width = min(up.shape[1], down.shape[1])
        height = min(up.shape[0], down.shape[0]) - 2 * finalcut
        ret = np.vstack([up[finalcut:finalcut + height, :width], down[finalcut:finalcut + height, :width]])
        y0 = height - bar // 2
    ret = ret.copy()
    ret[max(0, y0):max(0, min(ret.shape[0], y0 + bar))] = 0          # filled cv::rectangle covers exactly `bar` rows
    return ret


def fit2final(frame, canvas_size=(1920, 1080)):
    """nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189): copy when the sizes agree, else scale by
    fitscale = canvas_w / cols (only when wider than the canvas) with cv::resize(.., Size(), fitscale, fitscale) and paste
    centred on the zeroed canvas (:11).  Exactly 1/2 takes cv::resize's 2x2-average branch."""
    frame = np.ascontiguousarray(frame, np.uint8)
    cw, ch = canvas_size
    if frame.shape[1] == cw and frame.shape[0] == ch:
        return frame.copy()
    fitscale = cw * 1.0 / frame.shape[1] if frame.shape[1] > cw else 1.0
    if fitscale == 0.5 and frame.shape[1] % 2 == 0 and frame.shape[0] % 2 == 0:
        f = frame.astype(np.int32)
        tmp = ((f[0::2, 0::2] + f[0::2, 1::2] + f[1::2, 0::2] + f[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    else:
        tmp = orc.resize_bilinear_scaled_u8(frame, fitscale, fitscale) if fitscale != 1.0 else frame
    h, w = tmp.shape[:2]
    ox, oy = (cw - w) // 2, (ch - h) // 2
    canvas = np.zeros((ch, cw, 3), np.uint8)
    canvas[oy:oy + h, ox:ox + w] = tmp
    return canvas


def dilate3x3_u8(m):
    """cv::dilate(src, dst, Mat()) -- 3x3 rectangle, anchor at the centre, border pixels ignored
    (include/ocvstitcher.hpp:1095, 1251)."""
    m = np.asarray(m, np.uint8)
    p = np.zeros((m.shape[0] + 2, m.shape[1] + 2), np.uint8)
    p[1:-1, 1:-1] = m
    out = np.zeros_like(m)
    for dy in range(3):
        for dx in range(3):
            out = np.maximum(out, p[dy:dy + m.shape[0], dx:dx + m.shape[1]])
    return out


def linear_exact_axis(ssize, dsize):
    """Offsets and 8.8 fixed-point coefficients of cv::resize(INTER_LINEAR_EXACT) along one axis
    (OpenCV resize.cpp interpolationLinear::getCoeffs on softdouble == IEEE double, ufixedpoint16 = round(x * 256)).
    Positions before the first / after the last source sample replicate the edge (coefficients 256, 0)."""
    scale = 1.0 / (float(dsize) / float(ssize))
    ofs = np.zeros(dsize, np.int32); c1 = np.zeros(dsize, np.int32)
    for d in range(dsize):
        fval = scale * (d + 0.5) - 0.5
        ival = int(math.floor(fval))
        if ival >= 0 and ssize > 1:
            if ival < ssize - 1:
                ofs[d] = ival
                c1[d] = int(np.rint((fval - ival) * 256.0))
            else:
                ofs[d] = ssize - 1
        # else: replicate sample 0
    return ofs, c1


def resize_linear_exact_u8(src, dsize):
    """cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR_EXACT) for 8UC1 (include/ocvstitcher.hpp:1096, 1253):
    horizontal pass to 8.8 fixed point, vertical pass to 16.16, rounded half up."""
    src = np.asarray(src, np.uint8)
    dw, dh = dsize
    xo, xc = linear_exact_axis(src.shape[1], dw)
    yo, yc = linear_exact_axis(src.shape[0], dh)
    x1 = np.minimum(xo + 1, src.shape[1] - 1); y1 = np.minimum(yo + 1, src.shape[0] - 1)
    s = src.astype(np.int64)
    hrow = s[:, xo] * (256 - xc)[None, :] + s[:, x1] * xc[None, :]                # 8.8, <= 65280
    v = hrow[yo, :] * (256 - yc)[:, None] + hrow[y1, :] * yc[:, None]             # 16.16
    return ((v + 32768) >> 16).astype(np.uint8)


def seam_mask_tail(seam_lowres, full_mask):
    """m_blenderMask[i] from the seam finder's low-resolution mask: dilate -> INTER_LINEAR_EXACT up-scale to the
    warped size -> AND with the warped full mask (include/ocvstitcher.hpp:1095-1101, 1251-1257)."""
    up = resize_linear_exact_u8(dilate3x3_u8(seam_lowres), (full_mask.shape[1], full_mask.shape[0]))
    return up & np.asarray(full_mask, np.uint8)


def yuyv_to_bgra(yuyv):
    """cv::cvtColor(mtt, m_argb, cv::COLOR_YUV2BGRA_YUYV) of the YUYVCAM ingest (include/nvcam.hpp:880-886):
    OpenCV's ITU-R BT.601 conversion in 20-bit fixed point (color_yuv: CY 1220542, CUB 2116026, CUG -409993,
    CVG -852492, CVR 1673527), alpha 255.  yuyv: [H][W][2] uint8 (Y0 U | Y1 V per pixel pair)."""
    yuyv = np.asarray(yuyv, np.uint8)
    Y = yuyv[:, :, 0].astype(np.int64)
    U = np.repeat(yuyv[:, 0::2, 1], 2, axis=1).astype(np.int64) - 128
    V = np.repeat(yuyv[:, 1::2, 1], 2, axis=1).astype(np.int64) - 128
    half = 1 << 19
    ruv = half + 1673527 * V
    guv = half - 852492 * V - 409993 * U
    buv = half + 2116026 * U
    y = np.maximum(0, Y - 16) * 1220542
    out = np.empty(yuyv.shape[:2] + (4,), np.uint8)
    out[..., 0] = np.clip((y + buv) >> 20, 0, 255)
    out[..., 1] = np.clip((y + guv) >> 20, 0, 255)
    out[..., 2] = np.clip((y + ruv) >> 20, 0, 255)
    out[..., 3] = 255
    return out
