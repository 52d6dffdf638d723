"""`nvCam` pixel-pipeline mirror over the C ABI (include/nvcam.hpp:674-833, 898-929, 1083-1100).

Only the per-frame pixel work of read_frame/getFrame is in scope: V4L2 /
NvBuffer capture is Jetson I/O and is replaced by frames handed in by the caller.  The camera
entry (K, distortion, crop rect) comes from the `cameras:` table of cfg/cameras.yaml, matched
like nvCam::init does (:716-719).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import capi


@dataclass
class CamConfig:
    """stCamCfg (include/stitcherglobal.h:39-55) + the matched cameras.yaml entry."""
    camSrcWidth: int = 1920
    camSrcHeight: int = 1080
    undistoredWidth: int = 1920
    undistoredHeight: int = 1080
    outPutWidth: int = 1920
    outPutHeight: int = 1080
    undistor: bool = True
    K: Sequence[float] = field(default_factory=lambda: [1, 0, 0, 0, 1, 0, 0, 0, 1])
    distorParams: Sequence[float] = field(default_factory=lambda: [0, 0, 0, 0])
    rect: Sequence[int] = field(default_factory=lambda: [0, 0, 0, 0])
    newK: Optional[Sequence[float]] = None   # getOptimalNewCameraMatrix result; None -> computed via cv2
    device: int = 0
    max_batch: int = 1
    srcFormat: str = "bgra"     # "bgra": the VIC's 8UC4 output (:889-893); "yuyv": 8UC2 as captured, converted on the
                                # device like the YUYVCAM build's cv::cvtColor(COLOR_YUV2BGRA_YUYV) (:880-886)


def match_camera_entry(cameras_yaml, vendor, sensor, fov, srcsz, undistorsz, sttype="default"):
    """The lookup loop of nvCam::init (include/nvcam.hpp:716-719)."""
    import yaml
    table = yaml.safe_load(open(cameras_yaml))["cameras"]
    hit = None
    for e in table:
        if (str(e["vendor"]) == str(vendor) and str(e["sensor"]) == str(sensor) and int(e["fov"]) == int(fov) and
                int(e["srcsz"]) == int(srcsz) and int(e["undistorsz"]) == int(undistorsz) and
                str(e.get("sttype", "default")) == str(sttype)):
            hit = e
    if hit is None:
        raise KeyError("no cameras.yaml entry for %s/%s fov%s %s->%s" % (vendor, sensor, fov, srcsz, undistorsz))
    return dict(K=[float(v) for v in hit["K"]], distorParams=[float(v) for v in hit["distorParams"]],
                rect=[int(v) for v in hit["rect"]])


class nvCamFrontEnd:
    def __init__(self, cfg: CamConfig, maps=None):
        self.cfg = cfg
        self._lib = capi.lib()
        fc = capi.pano_frontend_config()
        fc.cam_src_width, fc.cam_src_height = cfg.camSrcWidth, cfg.camSrcHeight
        fc.undist_width, fc.undist_height = cfg.undistoredWidth, cfg.undistoredHeight
        fc.out_width, fc.out_height = cfg.outPutWidth, cfg.outPutHeight
        fc.undistort = int(cfg.undistor)
        newK = cfg.newK
        if cfg.undistor and newK is None and maps is None:
            # prepareUndistorMap (:831): one-time host init through OpenCV
            import cv2
            size = (cfg.undistoredWidth, cfg.undistoredHeight)
            newK, _ = cv2.getOptimalNewCameraMatrix(np.asarray(cfg.K, np.float64).reshape(3, 3),
                                                    np.asarray(cfg.distorParams, np.float64), size, 1, size, 0)
        self.newK = None if newK is None else np.asarray(newK, np.float64).reshape(3, 3)
        fc.K = (C.c_double * 9)(*[float(v) for v in np.asarray(cfg.K).reshape(-1)])
        fc.D = (C.c_double * 4)(*[float(v) for v in list(cfg.distorParams)[:4]])
        fc.newK = (C.c_double * 9)(*([float(v) for v in self.newK.reshape(-1)] if self.newK is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]))
        fc.rect = (C.c_int * 4)(*[int(v) for v in cfg.rect])
        self._keep = None
        if maps is not None:
            mx = np.ascontiguousarray(maps[0], np.float32); my = np.ascontiguousarray(maps[1], np.float32)
            self._keep = (mx, my)
            fc.mapx = mx.ctypes.data_as(C.POINTER(C.c_float)); fc.mapy = my.ctypes.data_as(C.POINTER(C.c_float))
        fc.device = cfg.device; fc.max_batch = cfg.max_batch
        fc.src_format = 1 if cfg.srcFormat == "yuyv" else 0
        h = C.c_void_p()
        if self._lib.pano_frontend_create(C.byref(fc), C.byref(h)) != capi.PANO_OK:
            raise capi.PanoError(self._lib.pano_frontend_last_error(None).decode())
        self._h = h

    def maps(self):
        w, h = self.cfg.undistoredWidth, self.cfg.undistoredHeight
        mx = np.empty((h, w), np.float32); my = np.empty((h, w), np.float32)
        capi.check(self._lib.pano_frontend_get_maps(self._h, capi.ptr(mx), capi.ptr(my)), self._h, frontend=True)
        return mx, my

    def getFrame(self, argb: np.ndarray) -> np.ndarray:
        """read_frame's pixel pipeline + getFrame(mat, src=false) for one 8UC4 (or 8UC2 YUYV) host frame."""
        argb = np.ascontiguousarray(argb, np.uint8)
        out = np.empty((self.cfg.outPutHeight, self.cfg.outPutWidth, 3), np.uint8)
        capi.check(self._lib.pano_frontend_process(self._h, capi.ptr(argb), argb.strides[0], capi.ptr(out), out.strides[0]),
                   self._h, frontend=True)
        return out

    def process_device(self, argb, out, stream=None):
        """argb: uint8 torch tensor [B, H, W, 4] on the device -> out [B, oh, ow, 3]."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(argb.device).cuda_stream
        capi.check(self._lib.pano_frontend_process_device(self._h, capi.ptr(argb), capi.ptr(out), int(argb.shape[0]),
                                                          C.c_void_p(stream)), self._h, frontend=True)
        return out

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.pano_frontend_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
