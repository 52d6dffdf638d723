"""`ocvStitcher`-shaped host mirror over the C ABI.

Mirrors the reference class surface (include/ocvstitcher.hpp:254-1306): `init(cfgPath)`,
`calibration(imgs)`, `process(imgs, ret)`, `updateMask(imgs)`, returning RET_OK / RET_ERR
(include/stitcherglobal.h:13-14).  What differs by design:

* per-frame work (`process`) is entirely CUDA behind `pano_process*`; nothing here touches pixels;
* one-time host init stays on the host exactly like the reference: the seam search
  (GraphCutSeamFinder on low-res warps, dilate, INTER_LINEAR_EXACT upsample, AND;
  ocvstitcher.hpp:975-1101 / :1218-1261) runs through OpenCV (`cv2`) and only produces the
  static masks uploaded with `pano_set_mask`; warp tables are built inside the library;
* `initAll` (SURF + matcher + bundle adjustment, :654-974) is feature-matching init and is out
  of scope: `calibration` with initMode 1 returns RET_ERR.
"""
import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import capi

RET_OK, RET_ERR = 0, -1
enInitALL, enInitByDefault, enInitByCfg = 1, 2, 3


@dataclass
class StitcherConfig:
    """stStitcherCfg (include/stitcherglobal.h:68-81) + the calibration members the reference
    keeps on the object (camK, cameraR, warped_image_scale, m_cutParams)."""
    width: int = 0
    height: int = 0
    id: int = 0
    num_images: int = 0
    blendStrength: float = 1.0
    initMode: int = enInitByDefault
    cfgPath: str = ""
    Ks: List[np.ndarray] = field(default_factory=list)
    Rs: List[np.ndarray] = field(default_factory=list)
    warped_image_scale: float = 0.0
    cut: Optional[Sequence[int]] = None      # m_cutParams; None -> whole dst roi
    cut_height: Optional[int] = None         # rule of :959-964: [0,(rows-h)/2,cols,h]
    warp: str = "spherical"                  # 'spherical' (:1000) | 'cylindrical' (:999)
    blender: Optional[str] = None            # None -> reference rule (:1188-1195); 'multiband'|'feather'|'no'
    num_bands: Optional[int] = None          # pin the band count (BASELINE config 1 pins 5)
    sharpness: Optional[float] = None        # feather; None -> 1/blend_width (stitching_detailed.cpp:868)
    seam: str = "gc_color"                   # 'gc_color' (:1033) | 'no' (:1034)
    exact_weights: bool = False              # True: upload cv2-built float weight tables over the device-built ones (cross-check
                                             # only: the device builders are bit-exact with cv2 4.x, tests/test_gpu_parity.py)
    device: int = 0
    max_batch: int = 1


def parse_camera_params_file(path, num_images=None):
    """Last block of a cameraparaout_<id>.txt (initCamParams, :452-520).  Handles the current
    layout (N lines of 18 floats + scale) and the 2021 layout (shared K, N R lines, scale)."""
    lines = [l.strip() for l in open(path).read().splitlines() if l.strip()]
    starts = [i for i, l in enumerate(lines) if ":" in l]
    if not starts:
        raise ValueError("no parameter block in %s" % path)
    body = lines[starts[-1] + 1:]
    rows = [[float(v) for v in l.split(",") if v.strip()] for l in body]
    scale = rows[-1][0]
    rows = rows[:-1]
    if all(len(r) == 18 for r in rows):
        Ks = [np.array(r[:9], np.float32).reshape(3, 3) for r in rows]
        Rs = [np.array(r[9:], np.float32).reshape(3, 3) for r in rows]
    elif len(rows[0]) == 9 and all(len(r) == 9 for r in rows):
        K = np.array(rows[0], np.float32).reshape(3, 3)
        Rs = [np.array(r, np.float32).reshape(3, 3) for r in rows[1:]]
        Ks = [K.copy() for _ in Rs]
    else:
        raise ValueError("camera parameter block malformed in %s" % path)
    if num_images is not None and len(Ks) != num_images:
        raise ValueError("expected %d cameras, file has %d" % (num_images, len(Ks)))
    return Ks, Rs, float(np.float32(scale))


def scale_intrinsics(Ks, scale, factor):
    """Re-target a calibration to frames `factor` times larger (K rows 0-1 and the warp scale)."""
    f = np.float32(factor)
    out = []
    for K in Ks:
        K = np.array(K, np.float32, copy=True)
        K[0, 0] *= f; K[0, 2] *= f; K[1, 1] *= f; K[1, 2] *= f
        out.append(K)
    return out, float(np.float32(scale) * f)


class ocvStitcher:
    m_num = 0  # construction-order id, like the reference's static counter (:1308)

    def __init__(self, config: Optional[StitcherConfig] = None):
        self.m_cfg = config
        self._h = None
        self._lib = None
        self.m_corners, self.m_sizes, self.dst_roi = [], [], None
        self.m_blenderMask: List[np.ndarray] = []
        self.m_cutParams = None
        self.last_error = ""

    # ------------------------------------------------------------------ init
    def init(self, stitcherCfgPath) -> int:
        """Parse a stitcher yaml (+ the rig table it points to) like init(std::string&) :262-358."""
        try:
            import yaml
            cfg = yaml.safe_load(open(stitcherCfgPath))
            sc = StitcherConfig()
            sc.width = int(cfg["outPutWidth"]); sc.height = int(cfg["outPutHeight"])
            sc.id = ocvStitcher.m_num
            ocvStitcher.m_num += 1
            sc.num_images = int(cfg["num_images"])
            sc.blendStrength = float(cfg["stitcherBlenderStrength"])
            sc.cfgPath = str(cfg["camcfgpath"])
            sc.initMode = int(cfg["initMode"])
            cams_yaml = cfg["cameraparams"]
            if not os.path.isabs(cams_yaml) or not os.path.exists(cams_yaml):
                cams_yaml = os.path.join(os.path.dirname(os.path.abspath(stitcherCfgPath)), os.path.basename(cams_yaml))
            rig = yaml.safe_load(open(cams_yaml))
            found = None
            for s in rig["structures"]:
                if (str(s["vendor"]) == str(cfg["vendor"]) and str(s["sensor"]) == str(cfg["sensor"]) and
                        str(s["sttype"]) == str(cfg["sttype"]) and bool(s["undistor"]) == bool(cfg["undistor"]) and
                        int(s["fov"]) == int(cfg["fov"]) and int(s["inputsz"]) == int(cfg["outPutWidth"])):
                    found = s["params"][sc.id]   # the last match wins in the reference loop too
            if found is None:
                self.last_error = "default structure params invalid, camera and structure can't be matched"
                return RET_ERR
            vals = [float(v) for v in found["cams"]]
            sc.Ks = [np.array(vals[18 * i:18 * i + 9], np.float32).reshape(3, 3) for i in range(sc.num_images)]
            sc.Rs = [np.array(vals[18 * i + 9:18 * i + 18], np.float32).reshape(3, 3) for i in range(sc.num_images)]
            sc.warped_image_scale = float(np.float32(vals[-1]))
            sc.cut = [int(v) for v in found["cut"]]
            self.m_cfg = sc
            return RET_OK
        except Exception as e:  # the reference catches everything and returns RET_ERR (:344-348)
            self.last_error = "stitcher yml parse failed: %r" % (e,)
            return RET_ERR

    def calibration(self, imgs) -> int:
        """calibration(), :592-650.  Fixed-parameter modes only (see module docstring)."""
        c = self.m_cfg
        if c is None:
            self.last_error = "init() first"
            return RET_ERR
        if c.initMode == enInitByCfg:
            path = os.path.join(c.cfgPath, "cameraparaout_%d.txt" % c.id)
            try:
                c.Ks, c.Rs, c.warped_image_scale = parse_camera_params_file(path, c.num_images)
            except Exception as e:
                self.last_error = "no preset parameters (%s); initAll is not part of this library" % (e,)
                return RET_ERR
        elif c.initMode != enInitByDefault:
            self.last_error = "initMode %d (feature-matching calibration) is out of scope" % c.initMode
            return RET_ERR
        return self.initSeam(imgs)

    def initSeam(self, imgs) -> int:
        """initSeam(), :975-1139: geometry + seam masks; then the handle is ready to process."""
        try:
            self._create_handle()
            masks = self._seam_masks(imgs)
            self._upload_masks(masks, on_device=True)
            return RET_OK
        except capi.PanoError as e:
            self.last_error = str(e)
            return RET_ERR

    def initTables(self, masks=None, weight_levels=None, feather_weights=None) -> int:
        """initSeam with caller-supplied static tables (no OpenCV needed): masks = m_blenderMask
        (None -> the warped all-255 masks), weight_levels[cam][level>=1] = float weight pyramid
        levels, feather_weights[cam] = FeatherBlender weight maps."""
        try:
            self._create_handle()
            if masks is not None:
                self.m_blenderMask = [np.ascontiguousarray(m, np.uint8) for m in masks]
                for i, m in enumerate(self.m_blenderMask):
                    capi.check(self._lib.pano_set_mask(self._h, i, capi.ptr(m), m.shape[1], m.shape[0], m.strides[0]), self._h)
            if weight_levels is not None:
                for i, lv in enumerate(weight_levels):
                    for l, w in lv.items() if isinstance(lv, dict) else enumerate(lv):
                        if w is None:
                            continue
                        w = np.ascontiguousarray(w, np.float32)
                        capi.check(self._lib.pano_set_weight_level(self._h, i, int(l), capi.ptr(w), w.shape[1], w.shape[0]), self._h)
            if feather_weights is not None:
                for i, w in enumerate(feather_weights):
                    w = np.ascontiguousarray(w, np.float32)
                    capi.check(self._lib.pano_set_feather_weight(self._h, i, capi.ptr(w), w.shape[1], w.shape[0]), self._h)
            return RET_OK
        except capi.PanoError as e:
            self.last_error = str(e)
            return RET_ERR

    def set_mask(self, cam, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        capi.check(self._lib.pano_set_mask(self._h, cam, capi.ptr(m), m.shape[1], m.shape[0], m.strides[0]), self._h)

    def updateMask(self, imgs) -> int:
        """updateMask(), :1218-1261 -- explicit instead of every 200th process() call."""
        try:
            self._upload_masks(self._seam_masks(imgs), on_device=True)
            return RET_OK
        except capi.PanoError as e:
            self.last_error = str(e)
            return RET_ERR

    # ------------------------------------------------------------------ per frame
    def process(self, imgs, ret: Optional[np.ndarray] = None) -> np.ndarray:
        """process(), :1141-1216: N BGR frames in, cropped 8-bit panorama out."""
        c = self.m_cfg
        if self._h is None:
            raise capi.PanoError("process() before calibration()")
        if len(imgs) != c.num_images:
            raise capi.PanoError("expected %d images" % c.num_images)
        frames = [np.ascontiguousarray(im, np.uint8) for im in imgs]
        front = getattr(self, "_front", None)
        want = (c.height, c.width, 3) if front is None else (front.cfg.camSrcHeight, front.cfg.camSrcWidth,
                                                             2 if front.cfg.srcFormat == "yuyv" else 4)
        for f in frames:
            if f.shape != want:
                raise capi.PanoError("frame must be %dx%dx%d" % want)
        ow, oh = self.out_size
        if ret is None:
            ret = np.empty((oh, ow, 3), np.uint8)
        fp = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        st = (C.c_int * len(frames))(*[f.strides[0] for f in frames])
        capi.check(self._lib.pano_process(self._h, fp, st, capi.ptr(ret), ret.strides[0]), self._h)
        return ret

    def process_device(self, frames, out, stream=None):
        """Batch of frame-sets resident on the device.  frames: uint8 torch tensor
        [B, N, H, W, 3]; out: uint8 [B, cut_h, cut_w, 3].  Asynchronous on `stream`
        (an int cudaStream_t; default: torch's current stream)."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(frames.device).cuda_stream
        capi.check(self._lib.pano_process_device(self._h, capi.ptr(frames), capi.ptr(out), int(frames.shape[0]),
                                                 C.c_void_p(stream)), self._h)
        return out

    def process_batch(self, frames_host, out_host):
        """Batch of frame-sets in (preferably pinned) HOST memory, [B, N, H, W, 3] -> [B, h, w, 3];
        H2D / compose / D2H are pipelined inside the library."""
        capi.check(self._lib.pano_process_batch(self._h, capi.ptr(frames_host), capi.ptr(out_host),
                                                int(frames_host.shape[0])), self._h)
        return out_host

    # ------------------------------------------------------------------ inspection
    @property
    def out_size(self):
        wh = (C.c_int * 2)()
        capi.check(self._lib.pano_get_geometry(self._h, None, None, None, wh), self._h)
        return wh[0], wh[1]

    def warp_maps(self, cam):
        w, h = self.m_sizes[cam]
        xm = np.empty((h, w), np.float32); ym = np.empty((h, w), np.float32)
        capi.check(self._lib.pano_get_warp_maps(self._h, cam, capi.ptr(xm), capi.ptr(ym)), self._h)
        return xm, ym

    def fixed_maps(self, cam):
        w, h = self.m_sizes[cam]
        ixy = np.empty((h, w, 2), np.int16); fr = np.empty((h, w), np.uint16)
        capi.check(self._lib.pano_get_fixed_maps(self._h, cam, capi.ptr(ixy), capi.ptr(fr)), self._h)
        return ixy, fr

    def set_seam_mask(self, cam, seam_lowres):
        """updateMask tail on the device: dilate -> INTER_LINEAR_EXACT up-scale -> AND with the warped full mask
        (include/ocvstitcher.hpp:1251-1257), from the seam finder's low-resolution mask."""
        m = np.ascontiguousarray(seam_lowres, np.uint8)
        capi.check(self._lib.pano_set_seam_mask(self._h, cam, capi.ptr(m), m.shape[1], m.shape[0], m.strides[0]), self._h)

    def get_mask(self, cam):
        w, h = self.m_sizes[cam]
        m = np.empty((h, w), np.uint8)
        capi.check(self._lib.pano_get_mask(self._h, cam, capi.ptr(m), m.strides[0]), self._h)
        return m

    def weight_level(self, cam, level):
        """The float weight level the compose uses (library-built on the device, or the caller's override)."""
        if self.m_cfg.blender == "multiband":
            _, _, rects = self.blend_geometry()
            w, h = int(rects[cam][2]) >> level, int(rects[cam][3]) >> level
        else:
            w, h = self.m_sizes[cam]
        out = np.empty((h, w), np.float32)
        capi.check(self._lib.pano_get_weight_level(self._h, cam, level, capi.ptr(out)), self._h)
        return out

    def blend_geometry(self):
        n = self.m_cfg.num_images
        nb = C.c_int(); pwh = (C.c_int * 2)(); fr = np.empty((n, 4), np.int32)
        capi.check(self._lib.pano_get_blend_geometry(self._h, C.byref(nb), pwh, capi.ptr(fr)), self._h)
        return nb.value, (pwh[0], pwh[1]), [tuple(int(v) for v in r) for r in fr]

    def set_gain_maps(self, gains):
        """compensator->apply tables (stitching_detailed.cpp:841): full-res float maps or None."""
        for i, g in enumerate(gains):
            if g is None:
                capi.check(self._lib.pano_set_gain_map(self._h, i, None, 0, 0), self._h)
            elif np.ndim(g) == 0:
                capi.check(self._lib.pano_set_gain_scalar(self._h, i, float(g)), self._h)
            else:
                g = np.ascontiguousarray(g, np.float32)
                capi.check(self._lib.pano_set_gain_map(self._h, i, capi.ptr(g), g.shape[1], g.shape[0]), self._h)

    def attach_frontend(self, front, cam=-1):
        """Chain an nvCamFrontEnd: process*() then take 8UC4 camera frames (src/master.cpp:300-318)."""
        capi.check(self._lib.pano_attach_frontend(self._h, cam, front._h if front is not None else None), self._h)
        self._front = front

    def set_frontend_mode(self, fused: bool):
        """False: the reference's sequential undistort -> crop -> resize -> warp order (bit-exact, default).
        True: one composed remap table per camera, a single bilinear gather from the 8UC4 camera frame
        (PANO_FRONTEND_FUSED; not bit-exact -- reported separately with its own PSNR)."""
        capi.check(self._lib.pano_set_frontend_mode(self._h, 1 if fused else 0), self._h)

    def enable_profile(self, on=True):
        capi.check(self._lib.pano_profile_enable(self._h, int(on)), self._h)

    def read_profile(self):
        names = (C.c_char_p * 64)(); ms = (C.c_float * 64)(); cnt = (C.c_int * 64)(); by = (C.c_double * 64)()
        k = self._lib.pano_profile_read(self._h, 64, names, ms, cnt, by)
        if k < 0:
            capi.check(k, self._h)
        return [dict(name=names[i].decode(), ms=ms[i], launches=cnt[i], alg_bytes=by[i]) for i in range(k)]

    def last_launch_count(self):
        return self._lib.pano_last_launch_count(self._h)

    def close(self):
        if self._h is not None:
            self._lib.pano_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ internals
    def _blender_choice(self, dst_w, dst_h):
        """Blender selection of :1184-1199 unless the config pins one."""
        c = self.m_cfg
        blend_width = float(np.float32(math.sqrt(np.float32(dst_w * dst_h))) * np.float32(c.blendStrength) / np.float32(100.0))
        kind = c.blender
        if kind is None:
            kind = "no" if blend_width < 1.0 else "multiband"
        nb = c.num_bands
        if nb is None:
            nb = int(math.ceil(math.log(blend_width) / math.log(2.0)) - 1.0) if blend_width >= 1.0 else 0
        sharp = c.sharpness if c.sharpness is not None else (1.0 / blend_width if blend_width > 0 else 0.02)
        return kind, max(nb, 0), float(sharp)

    def _create_handle(self):
        c = self.m_cfg
        self._lib = capi.lib()
        self.close()
        kind = capi.WARP_SPHERICAL if c.warp == "spherical" else capi.WARP_CYLINDRICAL
        K = np.ascontiguousarray(np.stack([np.asarray(k, np.float32).reshape(3, 3) for k in c.Ks]))
        R = np.ascontiguousarray(np.stack([np.asarray(r, np.float32).reshape(3, 3) for r in c.Rs]))
        # dst size decides the blender (the reference does this on the first process() call)
        rois = [capi.host_warp_roi(kind, np.float32(c.warped_image_scale), K[i], R[i], c.width, c.height)
                for i in range(c.num_images)]
        geo = capi.host_blend_geometry([r[:2] for r in rois], [r[2:] for r in rois], 0)
        dst = geo["dst_roi"]
        bkind, nb, sharp = self._blender_choice(dst[2], dst[3])
        cut = c.cut
        if cut is None and c.cut_height is not None:
            cut = [0, (dst[3] - c.cut_height) // 2, dst[2], c.cut_height]
        cfg = capi.pano_config()
        cfg.num_images = c.num_images; cfg.src_width = c.width; cfg.src_height = c.height
        cfg.warp_kind = kind; cfg.warped_image_scale = np.float32(c.warped_image_scale)
        cfg.K = K.ctypes.data_as(C.POINTER(C.c_float)); cfg.R = R.ctypes.data_as(C.POINTER(C.c_float))
        cfg.blender = {"no": capi.BLEND_NO, "feather": capi.BLEND_FEATHER, "multiband": capi.BLEND_MULTIBAND}[bkind]
        cfg.num_bands = nb; cfg.sharpness = sharp
        cfg.cut = (C.c_int * 4)(*(cut if cut is not None else (0, 0, 0, 0)))
        cfg.device = c.device; cfg.max_batch = c.max_batch
        h = C.c_void_p()
        rc = self._lib.pano_create(C.byref(cfg), C.byref(h))
        if rc != capi.PANO_OK:
            raise capi.PanoError(self._lib.pano_last_error(None).decode())
        self._h = h
        self.blender_kind, self.num_bands, self.sharpness = bkind, nb, sharp
        n = c.num_images
        cs = np.empty((n, 2), np.int32); ss = np.empty((n, 2), np.int32); roi = (C.c_int * 4)()
        capi.check(self._lib.pano_get_geometry(h, capi.ptr(cs), capi.ptr(ss), roi, None), h)
        self.m_corners = [tuple(int(v) for v in r) for r in cs]
        self.m_sizes = [tuple(int(v) for v in r) for r in ss]
        self.dst_roi = tuple(roi)
        self.m_cutParams = list(cut) if cut is not None else [0, 0, roi[2], roi[3]]

    def _seam_masks(self, imgs):
        """m_blenderMask as initSeam / updateMask build it.  Host-side, one-time, through OpenCV."""
        import cv2
        c = self.m_cfg
        n = c.num_images
        K = [np.asarray(k, np.float32).reshape(3, 3) for k in c.Ks]
        R = [np.asarray(r, np.float32).reshape(3, 3) for r in c.Rs]
        swa = min(1.0, math.sqrt(1e5 / (c.height * c.width)))
        sw = cv2.PyRotationWarper(c.warp, np.float32(c.warped_image_scale * swa))
        corners, iw, mw = [], [], []
        for i in range(n):
            Ks = K[i].copy()
            f = np.float32(swa)
            Ks[0, 0] *= f; Ks[0, 2] *= f; Ks[1, 1] *= f; Ks[1, 2] *= f
            small = cv2.resize(np.ascontiguousarray(imgs[i]), None, fx=swa, fy=swa, interpolation=cv2.INTER_LINEAR_EXACT)
            cn, w = sw.warp(small, Ks, R[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
            _, m = sw.warp(np.full(small.shape[:2], 255, np.uint8), Ks, R[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
            corners.append(cn); iw.append(w); mw.append(m)
        if c.seam == "gc_color":
            finder = cv2.detail_GraphCutSeamFinder("COST_COLOR")
            um = finder.find([a.astype(np.float32) for a in iw], corners, [cv2.UMat(m) for m in mw])
            mw = [m.get() for m in um]
        # tail (:1095-1101, :1251-1257): dilate -> INTER_LINEAR_EXACT up-scale -> AND with the warped full mask runs on
        # the device (pano_set_seam_mask, bit-exact with the cv2 calls); only the low-resolution masks are uploaded
        masks = []
        for i in range(n):
            self.set_seam_mask(i, mw[i])
            masks.append(self.get_mask(i))
        return masks

    def _upload_masks(self, masks, on_device=False):
        self.m_blenderMask = [np.ascontiguousarray(m, np.uint8) for m in masks]
        for i, m in enumerate(self.m_blenderMask):
            if not on_device:      # pano_set_seam_mask already installed them
                capi.check(self._lib.pano_set_mask(self._h, i, capi.ptr(m), m.shape[1], m.shape[0], m.strides[0]), self._h)
        c = self.m_cfg
        if not c.exact_weights:
            return
        import cv2
        if self.blender_kind == "multiband":
            # MultiBandBlender::feed's weight pyramid, built by OpenCV itself (bit-exact parity)
            nb, _, rects = self.blend_geometry()
            geo = capi.host_blend_geometry(self.m_corners, self.m_sizes, self.num_bands)
            for i, m in enumerate(self.m_blenderMask):
                t, b, l, r = geo["borders"][i]
                w = cv2.copyMakeBorder(m.astype(np.float32) * np.float32(1.0 / 255.0), t, b, l, r, cv2.BORDER_CONSTANT)
                for lvl in range(1, nb + 1):
                    w = cv2.pyrDown(w)
                    w = np.ascontiguousarray(w, np.float32)
                    capi.check(self._lib.pano_set_weight_level(self._h, i, lvl, capi.ptr(w), w.shape[1], w.shape[0]), self._h)
        elif self.blender_kind == "feather":
            for i, m in enumerate(self.m_blenderMask):
                w = cv2.distanceTransform(m, cv2.DIST_L1, 3)
                _, w = cv2.threshold(w * np.float32(self.sharpness), 1.0, 1.0, cv2.THRESH_TRUNC)
                w = np.ascontiguousarray(w, np.float32)
                capi.check(self._lib.pano_set_feather_weight(self._h, i, capi.ptr(w), w.shape[1], w.shape[0]), self._h)
