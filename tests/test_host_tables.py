"""Product host-side table builders (csrc/geometry.cpp via the C ABI's pano_host_* entry points)
against the oracle.  CPU only: no compute kernels run."""
import numpy as np
import pytest

import panob200
import util
from golden import calib
from oracle import oracle as orc

capi = panob200.capi

try:
    import cv2  # noqa: F401
    HAVE_CV2 = True
except ImportError:
    HAVE_CV2 = False
needs_cv2 = pytest.mark.skipif(not HAVE_CV2, reason="cv2 is the comparison target")


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float32)


CAMS = []
for rig_name, width, h in (("2222", 480, 270), ("424", 640, 360), ("2222", 1920, 1080)):
    Ks, Rs, sc = calib.rig(rig_name, width)
    for K, R in zip(Ks, Rs):
        CAMS.append((K, R, sc, width, h))
K0 = np.array([[300, 0, 160], [0, 300, 120], [0, 0, 1]], np.float32)
CAMS.append((K0, rot_x(1.2), 280.0, 320, 240))       # looks at the sphere's pole
CAMS.append((K0, rot_x(-1.35) @ rot_y(0.4), 280.0, 320, 240))


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("cam", range(len(CAMS)))
def test_warp_tables_bit_exact(kind, cam):
    K, R, sc, W, H = CAMS[cam]
    if kind == 1 and cam >= len(CAMS) - 2:
        pytest.skip("cylindrical warp of a pole-facing camera is degenerate")
    if W == 1920 and cam % 4:
        pytest.skip("one full-size camera is enough")
    roi = capi.host_warp_roi(kind, np.float32(sc), K, R, W, H)
    assert roi == orc.warp_roi(kind, np.float32(sc), K, R, W, H)
    _, xm, ym = capi.host_build_maps(kind, np.float32(sc), K, R, W, H)
    _, oxm, oym = orc.build_maps(kind, np.float32(sc), K, R, W, H)
    assert np.array_equal(xm, oxm) and np.array_equal(ym, oym)
    ixy, fr = capi.host_fixed_maps(xm, ym)
    oixy, ofr = orc.convert_maps(oxm, oym)
    assert np.array_equal(ixy, oixy) and np.array_equal(fr, ofr)


def test_fixed_maps_extremes():
    xm = np.array([[-1, 1e6, -1e7, 5e9, 40000, -40000, 0.015624, 0.015626, 31.984375, np.float32(2047.99)]], np.float32)
    ym = xm[:, ::-1].copy()
    a = capi.host_fixed_maps(xm, ym)
    b = orc.convert_maps(xm, ym)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("n", [1, 2, 7, 64, 1080])
def test_fold_reflect_equals_border_reflect_taps(n):
    """value(reflect(i))*(32-f) + value(reflect(i+1))*f must equal the folded in-range sample."""
    rng = np.random.default_rng(n)
    v = rng.integers(0, 256, n).astype(np.int64)

    def refl(p):
        if n == 1:
            return 0
        while p < 0 or p >= n:
            p = -p - 1 if p < 0 else 2 * n - 1 - p
        return p
    lo = -3 * n - 5
    for i in range(lo, 3 * n + 5):
        for f in (0, 1, 13, 31):
            want = v[refl(i)] * (32 - f) + v[refl(i + 1)] * f
            s = capi.lib().pano_host_fold_reflect(i, f, n)
            j, g = s >> 5, s & 31
            assert 0 <= j < n
            got = v[j] * (32 - g) + v[min(j + 1, n - 1)] * g
            assert got == want, (i, f, n, s)


@pytest.mark.parametrize("num_bands", [0, 1, 3, 5, 7, 9])
def test_blend_geometry(num_bands):
    corner_sets = [
        ([(-840, 1903), (-1986, 1913), (-3265, 1904), (-4486, 1906)], [(1690, 1016), (1688, 1015), (1690, 1016), (1695, 1017)]),
        ([(-100, 20), (60, 33), (230, 25)], [(200, 150), (210, 140), (190, 155)]),
        ([(0, 0)], [(33, 17)]),
        ([(5, -7), (5, 300)], [(1000, 310), (999, 120)]),
    ]
    for corners, sizes in corner_sets:
        g = capi.host_blend_geometry(corners, sizes, num_bands)
        roi = orc.result_roi(corners, sizes)
        nb, pwh = orc.mb_prepare(roi, num_bands)
        assert g["dst_roi"] == roi and g["num_bands"] == nb and g["padded"] == pwh
        for i in range(len(corners)):
            rect, bd = orc.mb_feed_rect(roi, pwh, nb, corners[i], sizes[i])
            assert g["feed_rects"][i] == rect and g["borders"][i] == bd
            assert rect[2] % (1 << nb) == 0 and rect[3] % (1 << nb) == 0 and rect[0] % (1 << nb) == 0
            assert rect[0] + rect[2] <= pwh[0] and rect[1] + rect[3] <= pwh[1]


def test_config1_geometry_matches_survey():
    """corners/sizes/dst roi/feed rects of BASELINE config 1 (from the cv2 oracle)."""
    Ks, Rs, sc = calib.rig("2222", 1920)
    rois = [capi.host_warp_roi(0, np.float32(sc), K, R, 1920, 1080) for K, R in zip(Ks, Rs)]
    assert [r[:2] for r in rois] == [(-840, 1903), (-1986, 1913), (-3265, 1904), (-4486, 1906)]
    assert [r[2:] for r in rois] == [(1690, 1016), (1688, 1015), (1690, 1016), (1695, 1017)]
    g = capi.host_blend_geometry([r[:2] for r in rois], [r[2:] for r in rois], 5)
    assert g["dst_roi"] == (-4486, 1903, 5336, 1025) and g["padded"] == (5344, 1056)
    assert g["feed_rects"] == [(3520, 0, 1824, 1056), (2400, 0, 1888, 1056), (1120, 0, 1888, 1056), (0, 0, 1792, 1056)]


def test_cubic_table_and_resize_axes():
    assert np.array_equal(capi.host_cubic_table(), orc.cubic_table())
    for ss, ds in [(1920, 960), (1782, 1920), (889, 1080), (640, 1920), (100, 257), (7, 3), (3, 7), (2, 2)]:
        for clamp in (0, 1):
            o, a0, a1 = capi.host_resize_axis(ss, ds, clamp)
            assert (a0.astype(int) + a1.astype(int) == 2048).all()
            if clamp:
                assert o.min() >= 0 and o.max() <= ss - 1


def test_resize_axis_reproduces_oracle_resize():
    """numpy model of the kernel arithmetic driven by the product's axis tables == oracle resize."""
    src = util.synth_frame(54, 96, 3)
    for dw, dh in [(192, 108), (40, 31), (96, 54), (97, 55)]:
        want = orc.resize_bilinear_u8(src, (dw, dh))
        if (dw, dh) == (96, 54):
            assert np.array_equal(want, src)
            continue
        xo, xa0, xa1 = capi.host_resize_axis(96, dw, 1)
        yo, ya0, ya1 = capi.host_resize_axis(54, dh, 0)
        s = src.astype(np.int64)
        x1 = np.minimum(xo + 1, 95)
        y0 = np.clip(yo, 0, 53); y1 = np.clip(yo + 1, 0, 53)
        t = s[:, xo] * xa0[None, :, None] + s[:, x1] * xa1[None, :, None]
        v = (((ya0[:, None, None] * (t[y0] >> 4)) >> 16) + ((ya1[:, None, None] * (t[y1] >> 4)) >> 16) + 2) >> 2
        assert np.array_equal(np.clip(v, 0, 255).astype(np.uint8), want)


def test_undistort_maps_match_oracle():
    cam = calib.CAM_LIJING_390_FOV60_1920
    K = np.array(cam["K"]).reshape(3, 3)
    newK = np.array([[1627.5076, 0, 943.1681], [0, 1622.9720, 571.5369], [0, 0, 1]])
    mx, my = capi.host_undistort_maps(K, cam["distorParams"], newK, 960, 540)
    ox, oy = orc.init_undistort_map(K, cam["distorParams"], newK, 960, 540)
    assert np.array_equal(mx, ox) and np.array_equal(my, oy)


def test_float_pyrdown_and_feather_weight():
    rng = np.random.default_rng(3)
    for w, h in [(64, 32), (33, 17), (5, 3), (1, 1), (128, 7)] + [(w, 9) for w in range(1, 40)]:
        a = rng.random((h, w), np.float32)
        assert np.array_equal(capi.host_pyrdown_f32(a), orc.pyrdown_f32(a))
    cv2 = pytest.importorskip("cv2")
    # ... and against cv2 itself (per-column summation order of pyramids.cpp, geometry.cpp pyrDownColumnRule)
    for w, h in [(w, h) for w in range(1, 30) for h in (1, 2, 5, 8)] + [(1920, 1080), (5400, 968), (517, 301)]:
        a = rng.integers(0, 256, (h, w)).astype(np.float32) * np.float32(1.0 / 255.0)
        assert np.array_equal(capi.host_pyrdown_f32(a), cv2.pyrDown(a)), (w, h)
    m = np.zeros((90, 140), np.uint8)
    m[10:70, 20:120] = 255
    m[30:40, 50:60] = 0
    m[0:5, :] = 255
    for sharp in (0.02, 0.1, 1.0 / 37.3):
        want = cv2.distanceTransform(m, cv2.DIST_L1, 3)
        _, want = cv2.threshold(want * np.float32(sharp), 1.0, 1.0, cv2.THRESH_TRUNC)
        assert np.array_equal(capi.host_feather_weight(m, sharp), want)
    full = np.full((20, 30), 255, np.uint8)
    want = cv2.distanceTransform(full, cv2.DIST_L1, 3)
    _, want = cv2.threshold(want * np.float32(0.05), 1.0, 1.0, cv2.THRESH_TRUNC)
    assert np.array_equal(capi.host_feather_weight(full, 0.05), want)


def test_linear_exact_axis_matches_oracle_restatement():
    """cv::resize(INTER_LINEAR_EXACT) axis tables of pano_set_seam_mask (updateMask tail) vs the oracle's restatement
    (itself pinned live against cv2 in test_oracle_vs_cv2.py)."""
    from oracle import compose
    for ss, ds in ((171, 1690), (103, 1016), (5, 64), (100, 100), (64, 63), (1, 17), (200, 50), (7, 7)):
        o, c = capi.host_linear_exact_axis(ss, ds)
        wo, wc = compose.linear_exact_axis(ss, ds)
        assert np.array_equal(o, wo) and np.array_equal(c, wc), (ss, ds)


@needs_cv2
@pytest.mark.parametrize("warp,size", [("spherical", (1920, 1080)), ("spherical", (960, 540)), ("cylindrical", (1280, 720))])
def test_seam_finder_inputs_equal_cv2(warp, size):
    """pano_host_seam_input == the low-resolution resize + warps of initSeam / updateMask through cv2
    (include/ocvstitcher.hpp:985-1017, 1228-1242): corners, warped images and warped masks bit for bit."""
    import math
    import cv2
    W, H = size
    Ks, Rs, scale = calib.rig("2222", W)
    imgs = util.synth_set(4, H, W, 77)
    swa = min(1.0, math.sqrt(1e5 / (H * W)))
    assert panob200.capi.lib().pano_host_seam_scale(W, H) == swa
    sw = cv2.PyRotationWarper(warp, np.float32(scale * swa))
    kind = panob200.capi.WARP_SPHERICAL if warp == "spherical" else panob200.capi.WARP_CYLINDRICAL
    for i in range(4):
        K = Ks[i].astype(np.float32).copy()
        f = np.float32(swa)
        K[0, 0] *= f; K[0, 2] *= f; K[1, 1] *= f; K[1, 2] *= f
        small = cv2.resize(imgs[i], None, fx=swa, fy=swa, interpolation=cv2.INTER_LINEAR_EXACT)
        corner, want_img = sw.warp(small, K, Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, want_mask = sw.warp(np.full(small.shape[:2], 255, np.uint8), K, Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        roi, got_img, got_mask = panob200.capi.host_seam_input(kind, scale, Ks[i], Rs[i], imgs[i])
        assert tuple(roi[:2]) == tuple(corner) and (roi[3], roi[2]) == want_img.shape[:2]
        assert np.array_equal(got_mask, want_mask)
        assert np.array_equal(got_img, want_img), util.report("seam input %d" % i, got_img, want_img)
