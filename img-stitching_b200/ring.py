"""Two-ring epilogue mirror (the caller step right after the two `process` calls): src/master.cpp:321-326
(`resize` + `vconcat` + 10-row separator) and src/panocamimpl.cpp:354-360 (`finalcut` crop + `vconcat` + 4-row
separator), as one kernel over the C ABI (pano_ring_*)."""
import ctypes as C

import numpy as np

from . import capi

RESIZE, CROP = 0, 1


class RingComposer:
    def __init__(self, up_size, down_size, mode="resize", finalcut=0, bar=None, device=0):
        """up_size / down_size = (width, height) of the two panoramas; mode 'resize' (master.cpp) or 'crop' (panocamimpl)."""
        self._lib = capi.lib()
        c = capi.pano_ring_config()
        c.up_width, c.up_height = int(up_size[0]), int(up_size[1])
        c.down_width, c.down_height = int(down_size[0]), int(down_size[1])
        c.mode = RESIZE if mode == "resize" else CROP
        c.finalcut = int(finalcut)
        c.bar = int(bar if bar is not None else (10 if mode == "resize" else 4))
        c.device = device
        self.cfg = c
        h = C.c_void_p()
        rc = self._lib.pano_ring_create(C.byref(c), C.byref(h))
        if rc != capi.PANO_OK:
            raise capi.PanoError((self._lib.pano_ring_last_error(None) or b"pano_ring_create failed").decode())
        self._h = h
        wh = (C.c_int * 2)()
        self._lib.pano_ring_out_size(h, wh)
        self.out_size = (wh[0], wh[1])

    def _check(self, rc):
        if rc != capi.PANO_OK:
            raise capi.PanoError((self._lib.pano_ring_last_error(self._h) or b"pano_ring call failed").decode())

    def compose(self, up: np.ndarray, down: np.ndarray) -> np.ndarray:
        up = np.ascontiguousarray(up, np.uint8); down = np.ascontiguousarray(down, np.uint8)
        out = np.empty((self.out_size[1], self.out_size[0], 3), np.uint8)
        self._check(self._lib.pano_ring_compose(self._h, capi.ptr(up), up.strides[0], capi.ptr(down), down.strides[0],
                                                capi.ptr(out), out.strides[0]))
        return out

    def compose_device(self, up, down, out, stream=None):
        """torch uint8 CUDA tensors [batch, h, w, 3] (contiguous); asynchronous on `stream`."""
        b = up.shape[0]
        self._check(self._lib.pano_ring_compose_device(self._h, capi.ptr(up), up.shape[2] * 3, capi.ptr(down), down.shape[2] * 3,
                                                       capi.ptr(out), out.shape[2] * 3, b, C.c_void_p(stream or 0)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pano_ring_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FitCanvas:
    """nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189) over the C ABI (pano_fit_*): the stacked frame scaled by
    fitscale = min(1, canvas_w / cols) and pasted centred on the black canvas (1920 x 1080 in the reference)."""

    def __init__(self, in_size, canvas_size=(1920, 1080), device=0):
        self._lib = capi.lib()
        c = capi.pano_fit_config()
        c.in_width, c.in_height = int(in_size[0]), int(in_size[1])
        c.canvas_width, c.canvas_height = int(canvas_size[0]), int(canvas_size[1])
        c.device = device
        self.cfg = c
        h = C.c_void_p()
        if self._lib.pano_fit_create(C.byref(c), C.byref(h)) != capi.PANO_OK:
            raise capi.PanoError((self._lib.pano_fit_last_error(None) or b"pano_fit_create failed").decode())
        self._h = h
        rect = (C.c_int * 4)(); fs = C.c_double()
        self._lib.pano_fit_geometry(h, rect, C.byref(fs))
        self.rect, self.fitscale = tuple(rect), fs.value

    def _check(self, rc):
        if rc != capi.PANO_OK:
            raise capi.PanoError((self._lib.pano_fit_last_error(self._h) or b"pano_fit call failed").decode())

    def fit2final(self, frame: np.ndarray) -> np.ndarray:
        frame = np.ascontiguousarray(frame, np.uint8)
        out = np.empty((self.cfg.canvas_height, self.cfg.canvas_width, 3), np.uint8)
        self._check(self._lib.pano_fit_compose(self._h, capi.ptr(frame), frame.strides[0], capi.ptr(out), out.strides[0]))
        return out

    def fit2final_device(self, frames, out, stream=None):
        """torch uint8 CUDA tensors [batch, h, w, 3] -> [batch, canvas_h, canvas_w, 3]; asynchronous on `stream`."""
        self._check(self._lib.pano_fit_compose_device(self._h, capi.ptr(frames), frames.shape[2] * 3, capi.ptr(out), out.shape[2] * 3,
                                                      frames.shape[0], C.c_void_p(stream or 0)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pano_fit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
