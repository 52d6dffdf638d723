"""ctypes binding of include/panob200.h (the C ABI of lib/libpanob200.so).

This is the only route from Python into the product: there is no Python/NumPy compute path.
If the shared library is missing it is built with `make` (nvcc cross-compiles without a GPU);
if that fails the import raises -- nothing falls back to the CPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", "libpanob200.so")

PANO_OK, PANO_ERR = 0, -1
WARP_SPHERICAL, WARP_CYLINDRICAL = 0, 1
BLEND_NO, BLEND_FEATHER, BLEND_MULTIBAND = 0, 1, 2


class PanoError(RuntimeError):
    pass


def library_path():
    return _LIB


def build_library(force=False):
    """Compile csrc/ for sm_100a into lib/libpanob200.so."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    srcs.append(os.path.join(_HERE, "..", "include", "panob200.h"))
    stale = not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", _HERE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise PanoError("building libpanob200.so failed:\n" + r.stdout[-4000:])
    return _LIB


class pano_config(C.Structure):
    _fields_ = [("num_images", C.c_int), ("src_width", C.c_int), ("src_height", C.c_int),
                ("warp_kind", C.c_int), ("warped_image_scale", C.c_float),
                ("K", C.POINTER(C.c_float)), ("R", C.POINTER(C.c_float)),
                ("blender", C.c_int), ("num_bands", C.c_int), ("sharpness", C.c_float),
                ("cut", C.c_int * 4), ("device", C.c_int), ("max_batch", C.c_int)]


class pano_ring_config(C.Structure):
    _fields_ = [("up_width", C.c_int), ("up_height", C.c_int), ("down_width", C.c_int), ("down_height", C.c_int),
                ("mode", C.c_int), ("finalcut", C.c_int), ("bar", C.c_int), ("device", C.c_int)]


class pano_fit_config(C.Structure):
    _fields_ = [("in_width", C.c_int), ("in_height", C.c_int), ("canvas_width", C.c_int), ("canvas_height", C.c_int),
                ("device", C.c_int)]


class pano_frontend_config(C.Structure):
    _fields_ = [("cam_src_width", C.c_int), ("cam_src_height", C.c_int),
                ("undist_width", C.c_int), ("undist_height", C.c_int),
                ("out_width", C.c_int), ("out_height", C.c_int), ("undistort", C.c_int),
                ("K", C.c_double * 9), ("D", C.c_double * 4), ("newK", C.c_double * 9),
                ("rect", C.c_int * 4), ("mapx", C.POINTER(C.c_float)), ("mapy", C.POINTER(C.c_float)),
                ("device", C.c_int), ("max_batch", C.c_int), ("src_format", C.c_int)]


# every symbol include/panob200.h declares (tests check the list against the header)
_SIGS = {
    "pano_version": (C.c_char_p, []),
    "pano_last_error": (C.c_char_p, [C.c_void_p]),
    "pano_create": (C.c_int, [C.POINTER(pano_config), C.POINTER(C.c_void_p)]),
    "pano_destroy": (C.c_int, [C.c_void_p]),
    "pano_get_geometry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_get_blend_geometry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_get_warp_maps": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_get_fixed_maps": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_set_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pano_set_seam_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pano_get_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "pano_set_weight_level": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "pano_get_weight_level": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pano_set_feather_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "pano_set_gain_map": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "pano_set_gain_scalar": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "pano_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "pano_process_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pano_process_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "pano_strip_set_window": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pano_strip_phase_count": (C.c_int, [C.c_void_p]),
    "pano_strip_cameras": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "pano_strip_run_phase": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_strip_halo_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "pano_strip_halo_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_strip_halo_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_strip_set_window_hybrid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pano_strip_run_phases": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_strip_level_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "pano_strip_level_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_strip_level_unpack_all": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pano_strip_p2p_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    "pano_strip_p2p_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pano_strip_p2p_connect_local": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pano_strip_p2p_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pano_strip_p2p_push": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pano_strip_p2p_wait_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pano_strip_p2p_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_strip_run_p2p": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_strip_p2p_check": (C.c_int, [C.c_void_p]),
    "pano_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pano_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_last_launch_count": (C.c_int, [C.c_void_p]),
    "pano_frontend_create": (C.c_int, [C.POINTER(pano_frontend_config), C.POINTER(C.c_void_p)]),
    "pano_frontend_destroy": (C.c_int, [C.c_void_p]),
    "pano_frontend_last_error": (C.c_char_p, [C.c_void_p]),
    "pano_frontend_get_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_frontend_process_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pano_frontend_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "pano_host_seam_scale": (C.c_double, [C.c_int, C.c_int]),
    "pano_host_seam_input": (C.c_int, [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_fit_create": (C.c_int, [C.POINTER(pano_fit_config), C.POINTER(C.c_void_p)]),
    "pano_fit_destroy": (C.c_int, [C.c_void_p]),
    "pano_fit_last_error": (C.c_char_p, [C.c_void_p]),
    "pano_fit_geometry": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "pano_fit_compose_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pano_fit_compose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "pano_attach_frontend": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pano_set_frontend_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "pano_ring_create": (C.c_int, [C.POINTER(pano_ring_config), C.POINTER(C.c_void_p)]),
    "pano_ring_destroy": (C.c_int, [C.c_void_p]),
    "pano_ring_last_error": (C.c_char_p, [C.c_void_p]),
    "pano_ring_out_size": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pano_ring_compose_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pano_ring_compose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "pano_host_warp_roi": (C.c_int, [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pano_host_build_maps": (C.c_int, [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_host_blend_geometry": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_host_fold_reflect": (C.c_uint, [C.c_int, C.c_int, C.c_int]),
    "pano_host_fixed_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "pano_host_pyrdown_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pano_host_feather_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "pano_host_undistort_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pano_host_cubic_table": (C.c_int, [C.c_void_p]),
    "pano_host_resize_axis": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pano_host_linear_exact_axis": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded shared library (built on first use if stale).  Raises if unavailable."""
    global _lib
    if _lib is None:
        path = os.environ.get("PANOB200_LIB")          # A/B runs of two builds of the library; default: the in-tree build
        if not path:
            build_library()
            path = _LIB
        l = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def ptr(a):
    """numpy array / torch tensor / int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def check(rc, handle=None, frontend=False):
    if rc != PANO_OK:
        fn = lib().pano_frontend_last_error if frontend else lib().pano_last_error
        msg = fn(handle)
        raise PanoError(msg.decode() if msg else "panob200 call failed")


# ------------------------------------------------------------------ host-only helpers

def host_warp_roi(kind, scale, K, R, w, h):
    K = np.ascontiguousarray(K, np.float32); R = np.ascontiguousarray(R, np.float32)
    roi = (C.c_int * 4)()
    check(lib().pano_host_warp_roi(kind, scale, ptr(K), ptr(R), w, h, roi))
    return tuple(roi)


def host_seam_input(kind, scale, K, R, frame):
    """One camera's seam-finder input (pano_host_seam_input): -> (roi, image_warped, mask_warped)."""
    K = np.ascontiguousarray(K, np.float32); R = np.ascontiguousarray(R, np.float32)
    frame = np.ascontiguousarray(frame, np.uint8)
    h, w = frame.shape[:2]
    roi = (C.c_int * 4)()
    check(lib().pano_host_seam_input(kind, C.c_float(scale), ptr(K), ptr(R), None, w, h, 0, roi, None, None))
    iw = np.empty((roi[3], roi[2], 3), np.uint8); mw = np.empty((roi[3], roi[2]), np.uint8)
    check(lib().pano_host_seam_input(kind, C.c_float(scale), ptr(K), ptr(R), ptr(frame), w, h, frame.strides[0], roi, ptr(iw), ptr(mw)))
    return tuple(roi), iw, mw


def host_build_maps(kind, scale, K, R, w, h):
    K = np.ascontiguousarray(K, np.float32); R = np.ascontiguousarray(R, np.float32)
    roi = host_warp_roi(kind, scale, K, R, w, h)
    xm = np.empty((roi[3], roi[2]), np.float32); ym = np.empty_like(xm)
    check(lib().pano_host_build_maps(kind, scale, ptr(K), ptr(R), w, h, ptr(xm), ptr(ym)))
    return roi, xm, ym


def host_blend_geometry(corners, sizes, num_bands):
    n = len(corners)
    c = np.ascontiguousarray(corners, np.int32).reshape(n, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(n, 2)
    roi = (C.c_int * 4)(); nb = C.c_int(); pwh = (C.c_int * 2)()
    fr = np.empty((n, 4), np.int32); bd = np.empty((n, 4), np.int32)
    check(lib().pano_host_blend_geometry(n, ptr(c), ptr(s), num_bands, roi, C.byref(nb), pwh, ptr(fr), ptr(bd)))
    return dict(dst_roi=tuple(roi), num_bands=nb.value, padded=(pwh[0], pwh[1]),
                feed_rects=[tuple(int(v) for v in r) for r in fr], borders=[tuple(int(v) for v in r) for r in bd])


def host_fixed_maps(xmap, ymap):
    xmap = np.ascontiguousarray(xmap, np.float32); ymap = np.ascontiguousarray(ymap, np.float32)
    ixy = np.empty(xmap.shape + (2,), np.int16); fr = np.empty(xmap.shape, np.uint16)
    check(lib().pano_host_fixed_maps(ptr(xmap), ptr(ymap), xmap.size, ptr(ixy), ptr(fr)))
    return ixy, fr


def host_pyrdown_f32(a):
    a = np.ascontiguousarray(a, np.float32)
    h, w = a.shape
    o = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    check(lib().pano_host_pyrdown_f32(ptr(a), w, h, ptr(o)))
    return o


def host_feather_weight(mask, sharpness):
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    o = np.empty((h, w), np.float32)
    check(lib().pano_host_feather_weight(ptr(mask), w, h, w, C.c_float(sharpness), ptr(o)))
    return o


def host_undistort_maps(K, D, newK, w, h):
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    D = np.ascontiguousarray(list(np.asarray(D).reshape(-1))[:4], np.float64)
    newK = np.ascontiguousarray(newK, np.float64).reshape(9)
    mx = np.empty((h, w), np.float32); my = np.empty((h, w), np.float32)
    check(lib().pano_host_undistort_maps(ptr(K), ptr(D), ptr(newK), w, h, ptr(mx), ptr(my)))
    return mx, my


def host_cubic_table():
    t = np.empty((1024, 16), np.int16)
    check(lib().pano_host_cubic_table(ptr(t)))
    return t


def host_resize_axis(ssize, dsize, clamp_frac):
    o = np.empty(dsize, np.int32); a0 = np.empty(dsize, np.int16); a1 = np.empty(dsize, np.int16)
    check(lib().pano_host_resize_axis(ssize, dsize, int(clamp_frac), ptr(o), ptr(a0), ptr(a1)))
    return o, a0, a1


def host_linear_exact_axis(ssize, dsize):
    ofs = np.empty(dsize, np.int32); c1 = np.empty(dsize, np.int32)
    check(lib().pano_host_linear_exact_axis(ssize, dsize, ptr(ofs), ptr(c1)))
    return ofs, c1
