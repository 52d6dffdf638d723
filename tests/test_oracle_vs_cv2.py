"""Pins the oracle (oracle/pano_oracle.c) LIVE against cv2 -- the reference's own OpenCV classes --
on random + adversarial inputs.  Skipped where cv2 is not importable; the committed fixtures in
tests/golden/ (test_oracle_golden.py) carry the same pin without cv2."""
import numpy as np
import pytest

import util
from golden import calib
from oracle import compose, oracle as orc, cv2_reference as ref  # noqa: F401  (cv2_reference imports cv2)

cv2 = pytest.importorskip("cv2")


def eq(a, b, what):
    assert np.array_equal(a, b), util.report(what, a, b)


@pytest.mark.parametrize("name,kind", [("spherical", 0), ("cylindrical", 1)])
def test_rotation_warper_maps(name, kind):
    Ks, Rs, sc = calib.rig("2222", 960)
    for K, R in zip(Ks, Rs):
        w = cv2.PyRotationWarper(name, np.float32(sc))
        assert tuple(w.warpRoi((960, 540), K, R)) == orc.warp_roi(kind, np.float32(sc), K, R, 960, 540)
        _, xm, ym = w.buildMaps((960, 540), K, R)
        _, oxm, oym = orc.build_maps(kind, np.float32(sc), K, R, 960, 540)
        assert np.array_equal(xm, oxm) and np.array_equal(ym, oym)


def test_remap_variants_and_convert_maps():
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (97, 131, 3), np.uint8)
    xm = (rng.random((80, 120), np.float32) * 180 - 25).astype(np.float32)
    ym = (rng.random((80, 120), np.float32) * 150 - 25).astype(np.float32)
    xm[0, :10] = -1; ym[0, :10] = -1
    xm[1, :5] = 1e6; ym[1, :5] = -1e7; xm[2, :5] = 5e9; xm[3, :3] = 40000; ym[3, :3] = -40000
    eq(cv2.remap(src, xm, ym, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT), orc.remap_bilinear_u8(src, xm, ym, "reflect"), "linear reflect")
    eq(cv2.remap(src, xm, ym, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT), orc.remap_bilinear_u8(src, xm, ym, "const"), "linear const")
    m = rng.integers(0, 256, (97, 131), np.uint8)
    eq(cv2.remap(m, xm, ym, cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT), orc.remap_nearest_u8(m, xm, ym), "nearest")
    eq(cv2.remap(src, xm, ym, cv2.INTER_CUBIC), orc.remap_cubic_u8(src, xm, ym), "cubic")
    ixy, fr = cv2.convertMaps(xm, ym, cv2.CV_16SC2)
    oixy, ofr = orc.convert_maps(xm, ym)
    eq(ixy, oixy, "convertMaps ixy"); eq(fr, ofr, "convertMaps frac")


@pytest.mark.parametrize("sw,sh,dw,dh,c", [(131, 97, 64, 40, 3), (1920, 1080, 960, 540, 4), (1782, 889, 1920, 1080, 3),
                                           (100, 100, 257, 33, 3), (640, 360, 1920, 1080, 4), (960, 540, 1920, 1080, 3),
                                           (1920, 1080, 720, 405, 3), (50, 60, 50, 60, 3)])
def test_resize_bilinear(sw, sh, dw, dh, c):
    s = np.random.default_rng(sw + dw).integers(0, 256, (sh, sw, c), np.uint8)
    eq(cv2.resize(s, (dw, dh)), orc.resize_bilinear_u8(s, (dw, dh)), "resize")


@pytest.mark.parametrize("w,h", [(64, 32), (33, 17), (5, 3), (2, 2), (1, 1), (128, 7), (57, 33)])
def test_pyramids_s16(w, h):
    rng = np.random.default_rng(w * 100 + h)
    for lo, hi in ((-600, 600), (-32768, 32767)):
        s = rng.integers(lo, hi, (h, w, 3)).astype(np.int16)
        eq(cv2.pyrDown(s), orc.pyrdown_s16(s), "pyrDown")
        eq(cv2.pyrUp(s), orc.pyrup_s16(s, (2 * w, 2 * h)), "pyrUp even")
        if w > 1 and h > 1:
            eq(cv2.pyrUp(s, dstsize=(2 * w - 1, 2 * h - 1)), orc.pyrup_s16(s, (2 * w - 1, 2 * h - 1)), "pyrUp odd")


def test_float_pyrdown_is_bit_exact():
    """cv::pyrDown on CV_32F (the weight pyramid of MultiBandBlender::feed): OpenCV's vector bodies and scalar border /
    tail loops sum in different orders; with the per-column rule restated the oracle is bit-exact at every size."""
    rng = np.random.default_rng(0)
    sizes = [(h, w) for h in range(1, 41) for w in range(1, 41)] + [(132, 228), (301, 517), (128, 1030), (33, 2701), (545, 675)]
    for h, w in sizes:
        f = rng.random((h, w), np.float32) if (h + w) % 2 else rng.integers(0, 256, (h, w)).astype(np.float32) * np.float32(1.0 / 255.0)
        a, b = cv2.pyrDown(f), orc.pyrdown_f32(f)
        assert a.shape == b.shape and np.array_equal(a, b), (h, w)


def _blend_inputs():
    corners = [(-100, 20), (60, 33), (230, 25)]
    sizes = [(200, 150), (210, 140), (190, 155)]
    imgs = [util.synth_frame(s[1], s[0], 40 + i).astype(np.int16) for i, s in enumerate(sizes)]
    masks = []
    for s in sizes:
        m = np.zeros((s[1], s[0]), np.uint8)
        m[5:-7, 10:-4] = 255
        masks.append(cv2.GaussianBlur(m, (9, 9), 0))
    return corners, sizes, imgs, masks


@pytest.mark.parametrize("nb", [1, 3, 5, 7])
def test_multiband_blender(nb):
    corners, sizes, imgs, masks = _blend_inputs()
    roi = cv2.detail.resultRoi(corners=corners, sizes=sizes)
    bl = cv2.detail_MultiBandBlender(0, nb)
    bl.prepare(roi)
    for i in range(3):
        bl.feed(imgs[i], masks[i], corners[i])
    r, rm = bl.blend(None, None)
    nbe, pwh = orc.mb_prepare(roi, nb)
    ew = []
    for i in range(3):
        _, bd = orc.mb_feed_rect(roi, pwh, nbe, corners[i], sizes[i])
        ws = [cv2.copyMakeBorder(masks[i].astype(np.float32) * np.float32(1 / 255.), bd[0], bd[1], bd[2], bd[3], cv2.BORDER_CONSTANT)]
        for _ in range(nbe):
            ws.append(cv2.pyrDown(ws[-1]))
        ew.append(ws)
    o, om = orc.multiband_blend(imgs, masks, corners, sizes, nb, ew)
    eq(r, o, "multiband"); eq(rm, om, "multiband mask")


def test_feather_no_blend_and_gain():
    corners, sizes, imgs, masks = _blend_inputs()
    roi = cv2.detail.resultRoi(corners=corners, sizes=sizes)
    for sharp in (0.02, 0.1, 1 / 37.3):
        fb = cv2.detail_FeatherBlender(sharp)
        fb.prepare(roi)
        for i in range(3):
            fb.feed(imgs[i], masks[i], corners[i])
        r, rm = fb.blend(None, None)
        o, om = orc.feather_blend(imgs, [ref._feather_weight(m, sharp) for m in masks], corners, sizes)
        eq(r, o, "feather"); eq(rm, om, "feather mask")
    nbl = cv2.detail.Blender_createDefault(cv2.detail.Blender_NO)
    nbl.prepare(roi)
    for i in range(3):
        nbl.feed(imgs[i], masks[i], corners[i])
    r, rm = nbl.blend(None, None)
    o, om = orc.no_blend(imgs, masks, corners, sizes)
    eq(r, o, "no blend"); eq(rm, om, "no blend mask")
    img = util.synth_frame(50, 60, 3)
    gc = cv2.detail_BlocksGainCompensator(32, 32, 1)
    gains = [(np.random.default_rng(2).random((2, 2)).astype(np.float32) + 0.5)]
    gc.setMatGains(gains)
    got = gc.apply(0, (0, 0), img.copy(), np.full((50, 60), 255, np.uint8))
    eq(got, orc.gain_apply_u8(img, cv2.resize(gains[0], (60, 50), interpolation=cv2.INTER_LINEAR)), "blocks gain")
    gs = cv2.detail_GainCompensator(1)
    gs.setMatGains([np.array([[1.2345]], np.float64)])
    eq(gs.apply(0, (0, 0), img.copy(), np.full((50, 60), 255, np.uint8)), orc.gain_apply_u8(img, None, 1.2345), "scalar gain")


def test_undistort_maps_fixed_point_identical():
    cam = calib.CAM_LIJING_390_FOV60_1920
    K = np.array(cam["K"]).reshape(3, 3)
    D = np.array(cam["distorParams"])
    newK, mx, my = ref.undistort_tables(K, D, (1920, 1080))
    ox, oy = orc.init_undistort_map(K, D, newK, 1920, 1080)
    assert (mx != ox).sum() + (my != oy).sum() <= 16          # AVX2/FMA path of cv2: a few 1-ulp floats
    a, af = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    b, bf = orc.convert_maps(ox, oy)
    assert np.array_equal(a, b) and np.array_equal(af, bf)    # what remap consumes is identical


def test_full_call_sequence_matches_cv2_reference():
    """oracle/compose.process == cv2-driven ocvStitcher::process restatement (GraphCut masks)."""
    Ks, Rs, sc = calib.rig("2222", 480)
    imgs = util.synth_set(4, 270, 480, 3)
    t = ref.init_seam(imgs, Ks, Rs, sc)
    from oracle import compose
    ot = compose.build_tables(Ks, Rs, sc, (480, 270))
    assert ot.corners == t.corners and ot.sizes == t.sizes and ot.dst_roi == tuple(t.dst_roi)
    for a, b in zip(ot.warped_masks, t.warped_masks):
        eq(a, b, "warped mask")
    ot.blend_masks = t.blend_masks
    cut = ref.default_cut(t.dst_roi, 200)
    want = ref.process(t, imgs, "multiband", 4, cut=cut)
    got = compose.process(ot, imgs, "multiband", 4, cut=cut)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1, util.report("compose", got, want)     # own float weight pyramid: <= 1 LSB
    eq(ref.process(t, imgs, "no", cut=cut), compose.process(ot, imgs, "no", cut=cut), "no-blend compose")


def test_ring_epilogue_vs_cv2_calls():
    """The two-ring caller step (src/master.cpp:321-326, src/panocamimpl.cpp:354-360) replayed with cv2."""
    from oracle import compose
    rng = np.random.default_rng(9)
    for (uw, uh), (dw, dh) in (((333, 57), (301, 64)), ((200, 40), (200, 40)), ((150, 33), (211, 30))):
        up = rng.integers(0, 256, (uh, uw, 3), np.uint8)
        down = rng.integers(0, 256, (dh, dw, 3), np.uint8)
        r = cv2.resize(up, (dw, dh))
        ret = cv2.vconcat([r, down])
        cv2.rectangle(ret, (0, ret.shape[0] // 2 - 5, ret.shape[1], 10), (0, 0, 0), -1)
        eq(compose.ring_epilogue(up, down, "resize"), ret, "master.cpp epilogue %dx%d" % (uw, uh))
        for fc in (0, 3):
            width, height = min(uw, dw), min(uh, dh) - 2 * fc
            ret = cv2.vconcat([np.ascontiguousarray(up[fc:fc + height, :width]), np.ascontiguousarray(down[fc:fc + height, :width])])
            cv2.rectangle(ret, (0, height - 2, width, 4), (0, 0, 0), -1, 1, 0)
            eq(compose.ring_epilogue(up, down, "crop", finalcut=fc), ret, "panocamimpl epilogue finalcut %d" % fc)


def test_seam_mask_tail_vs_cv2_calls():
    """dilate -> INTER_LINEAR_EXACT -> AND (include/ocvstitcher.hpp:1095-1101, 1251-1257) replayed with cv2."""
    from oracle import compose
    rng = np.random.default_rng(3)
    for (sw, sh), (dw, dh) in (((171, 103), (1690, 1016)), ((40, 30), (97, 131)), ((5, 7), (64, 33)), ((100, 60), (100, 60)),
                               ((64, 64), (63, 65)), ((1, 9), (17, 40)), ((33, 1), (90, 5)), ((200, 100), (50, 30))):
        for binary in (True, False):
            m = (rng.integers(0, 2, (sh, sw)) * 255).astype(np.uint8) if binary else rng.integers(0, 256, (sh, sw), np.uint8)
            full = (rng.integers(0, 8, (dh, dw)) > 0).astype(np.uint8) * 255
            eq(compose.dilate3x3_u8(m), cv2.dilate(m, None), "dilate")
            eq(compose.resize_linear_exact_u8(m, (dw, dh)), cv2.resize(m, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT), "linear exact")
            want = cv2.bitwise_and(cv2.resize(cv2.dilate(m, None), (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT), full)
            eq(compose.seam_mask_tail(m, full), want, "seam tail")


def test_yuyv_ingest_vs_cv2_cvtcolor():
    """cv::cvtColor(COLOR_YUV2BGRA_YUYV) (include/nvcam.hpp:880-886): random frames plus a dense sweep of (Y, U, V)."""
    from oracle import compose
    rng = np.random.default_rng(0)
    f = rng.integers(0, 256, (64, 96, 2), np.uint8)
    eq(compose.yuyv_to_bgra(f), cv2.cvtColor(f, cv2.COLOR_YUV2BGRA_YUYV), "random yuyv")
    yy, uu, vv = np.meshgrid(np.arange(256), np.arange(0, 256, 3), np.arange(0, 256, 5), indexing="ij")
    n = yy.size
    a = np.zeros((1, 2 * n, 2), np.uint8)
    a[0, 0::2, 0] = yy.ravel(); a[0, 1::2, 0] = 255 - yy.ravel(); a[0, 0::2, 1] = uu.ravel(); a[0, 1::2, 1] = vv.ravel()
    eq(compose.yuyv_to_bgra(a), cv2.cvtColor(a, cv2.COLOR_YUV2BGRA_YUYV), "yuv sweep")


@pytest.mark.parametrize("size", [(5336, 1792), (2500, 700), (3840, 1000), (1920, 1080), (1000, 500), (3840, 2160)])
def test_fit2final_oracle_equals_cv2(size):
    """nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189) restated vs the cv2 calls it makes."""
    w, h = size
    f = util.synth_frame(h, w, 5)
    if (w, h) == (1920, 1080):
        want = f.copy()
    else:
        fs = 1920 * 1.0 / w if w > 1920 else 1
        tmp = cv2.resize(f, None, fx=fs, fy=fs)
        want = np.zeros((1080, 1920, 3), np.uint8)
        ox, oy = (1920 - tmp.shape[1]) // 2, (1080 - tmp.shape[0]) // 2
        want[oy:oy + tmp.shape[0], ox:ox + tmp.shape[1]] = tmp
    assert np.array_equal(compose.fit2final(f), want)
