"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the
committed cv2 golden vectors.  Integer/byte work -> the bar is BIT-EXACT (max|d| = 0), library-built weight
pyramids included (the float pyrDown follows OpenCV's per-column summation order)."""
import os

import numpy as np
import pytest

import panob200
import util
from golden import calib
from oracle import compose, oracle as orc

pytestmark = pytest.mark.gpu
SC = panob200.StitcherConfig


def load(name):
    return np.load(os.path.join(util.GOLDEN, name + ".npz"))


def make(Ks, Rs, scale, W, H, warp="spherical", blender="multiband", num_bands=5, cut=None, max_batch=1, sharp=None):
    st = panob200.ocvStitcher(SC(width=W, height=H, num_images=len(Rs), Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                 warp=warp, blender=blender, num_bands=num_bands, cut=cut, max_batch=max_batch,
                                 sharpness=sharp))
    return st


def golden_setup(name, warp, blender="multiband", gains=False):
    g = load(name)
    W, H = int(g["width"]), int(g["height"])
    Ks, Rs, scale = calib.rig("2222", W)
    n = len(Rs)
    nb = int(g["num_bands"])
    st = make(Ks, Rs, scale, W, H, warp, blender, nb, cut=[int(v) for v in g["cut"]],
              sharp=float(g["sharpness"]) if "sharpness" in g.files else None)
    masks = [g["mask%d" % i] for i in range(n)]
    wl = [{l: g["w%d_%d" % (i, l)] for l in range(1, nb + 1)} for i in range(n)] if blender == "multiband" else None
    fw = [g["fw%d" % i] for i in range(n)] if blender == "feather" else None
    assert st.initTables(masks, wl, fw) == 0, st.last_error
    if gains:
        st.set_gain_maps([g["gain%d" % i] for i in range(n)])
    return g, st, util.synth_set(n, H, W, int(g["seed"]))


def assert_equal(name, got, want):
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want), util.report(name, got, want)


# ------------------------------------------------------------------ golden vectors (cv2 4.13)

@pytest.mark.parametrize("name,warp", [("cfg1_small", "spherical"), ("cfg1_cyl_small", "cylindrical")])
def test_multiband_golden_bit_exact(name, warp):
    g, st, imgs = golden_setup(name, warp)
    assert np.array_equal(np.array(st.m_corners), g["corners"]) and np.array_equal(np.array(st.m_sizes), g["sizes"])
    assert st.dst_roi == tuple(g["dst_roi"])
    assert_equal(name, st.process(imgs), g["pano_multiband"])
    assert st.last_launch_count() >= 3


def test_multiband_library_weights_bit_exact():
    """Weight pyramids built by the library on the device (no cv2 in the loop): since the float pyrDown follows
    OpenCV's per-column summation order (pyrDownColumnRule) the panorama equals cv2's byte for byte."""
    g = load("cfg1_small")
    Ks, Rs, scale = calib.rig("2222", 240)
    st = make(Ks, Rs, scale, 240, 135, num_bands=5, cut=[int(v) for v in g["cut"]])
    assert st.initTables([g["mask%d" % i] for i in range(4)]) == 0, st.last_error
    out = st.process(util.synth_set(4, 135, 240, int(g["seed"])))
    assert_equal("library-built weights vs cv2 golden", out, g["pano_multiband"])


def test_feather_no_gain_golden_bit_exact():
    g, st, imgs = golden_setup("cfg1_small", "spherical", "feather")
    assert_equal("feather", st.process(imgs), g["pano_feather"])
    st.set_gain_maps([g["gain%d" % i] for i in range(4)])
    assert_equal("gain+feather", st.process(imgs), g["pano_gain_feather"])
    st.set_gain_maps([None] * 4)
    assert_equal("feather again", st.process(imgs), g["pano_feather"])
    g, st, imgs = golden_setup("cfg1_small", "spherical", "no")
    assert_equal("no", st.process(imgs), g["pano_no"])
    g, st, imgs = golden_setup("cfg1_small", "spherical", "multiband", gains=True)
    assert_equal("gain+multiband", st.process(imgs), g["pano_gain_multiband"])


def test_feather_weights_built_on_device():
    """FeatherBlender weight maps (distanceTransform(mask, DIST_L1, 3) * sharpness, clamped at 1) are built by two scan
    kernels at every mask refresh: bit-identical to the host builder and to cv2's distanceTransform + threshold."""
    Ks, Rs, scale = calib.rig("2222", 320)
    st = make(Ks, Rs, scale, 320, 180, "spherical", "feather", 0, sharp=0.037)
    assert st.initTables() == 0, st.last_error
    rng = np.random.default_rng(5)
    masks = []
    for i, (w, h) in enumerate(st.m_sizes):
        m = st.get_mask(i).copy()
        if i == 1:
            m[:] = 255                                            # no zero pixel anywhere: every weight saturates to 1
        elif i == 2:
            m[rng.integers(0, h, 40), rng.integers(0, w, 40)] = 0   # isolated holes
        elif i == 3:
            m[:, : w // 3] = 0
            m[h // 2, :] = 0
        masks.append(m)
        st.set_mask(i, m)
    for i, m in enumerate(masks):
        got = st.weight_level(i, 0)
        want = panob200.capi.host_feather_weight(m, 0.037)
        assert np.array_equal(got, want), "camera %d: device feather weights differ from the host builder" % i
        try:
            import cv2
        except ImportError:
            continue
        w = cv2.distanceTransform(m, cv2.DIST_L1, 3)
        _, w = cv2.threshold(w * np.float32(0.037), 1.0, 1.0, cv2.THRESH_TRUNC)
        assert np.array_equal(got, w), "camera %d: device feather weights differ from cv2" % i


# ------------------------------------------------------------------ against the oracle

def oracle_case(Ks, Rs, scale, W, H, warp, blender, nb, seed, masks="soft", cut=None, gains=None, sharp=0.05):
    t = compose.build_tables(Ks, Rs, scale, (W, H), warp)
    if masks == "soft":
        t.blend_masks = util.soft_masks(t)
    imgs = util.synth_set(len(Rs), H, W, seed)
    st = make(Ks, Rs, scale, W, H, warp, blender, nb, cut=cut, sharp=sharp)
    fw = None
    if blender == "feather":
        fw = [panob200.capi.host_feather_weight(m, sharp) for m in t.blend_masks]
    assert st.initTables(t.blend_masks if masks == "soft" else None, None, None) == 0, st.last_error
    if gains is not None:
        t.gain_maps = gains
        st.set_gain_maps(gains)
    want = compose.process(t, imgs, blender, nb, feather_weights=fw, cut=cut)
    return t, st, imgs, want


@pytest.mark.parametrize("nb", [0, 1, 2, 4, 7])
def test_band_counts_vs_oracle(nb):
    Ks, Rs, scale = calib.rig("2222", 320)
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 320, 180, "spherical", "multiband", nb, seed=10 + nb)
    assert_equal("nb=%d" % nb, st.process(imgs), want)


def test_default_masks_and_fixed_maps_vs_oracle():
    Ks, Rs, scale = calib.rig("424", 320)
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 320, 180, "spherical", "multiband", 3, seed=3, masks="default")
    assert_equal("default masks", st.process(imgs), want)
    for i in range(len(Rs)):
        ixy, fr = st.fixed_maps(i)
        oixy, ofr = orc.convert_maps(*t.maps[i])
        assert np.array_equal(ixy, oixy) and np.array_equal(fr, ofr)        # remap index maps: bit-exact
        xm, ym = st.warp_maps(i)
        assert np.array_equal(xm, t.maps[i][0]) and np.array_equal(ym, t.maps[i][1])


def test_odd_geometry_three_cameras_cylindrical():
    Ks, Rs, scale = calib.ring(3, 333, 187, 70.0, 38.0)
    roi = compose.build_tables(Ks, Rs, scale, (333, 187), "cylindrical").dst_roi
    cut = [7, 5, roi[2] - 20, roi[3] - 11]
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 333, 187, "cylindrical", "multiband", 4, seed=21, cut=cut)
    assert_equal("odd cyl", st.process(imgs), want)


def test_feather_and_no_blend_vs_oracle_with_scalar_gain():
    Ks, Rs, scale = calib.rig("2222", 320)
    gains = [1.0, 0.83, 1.21, 1.07]
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 320, 180, "spherical", "feather", 0, seed=5, gains=gains, sharp=0.04)
    assert_equal("feather scalar gain", st.process(imgs), want)
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 320, 180, "spherical", "no", 0, seed=6)
    assert_equal("no blend", st.process(imgs), want)


def test_wide_source_uses_64bit_map_entries():
    Ks, Rs, scale = calib.ring(2, 2300, 96, 80.0, 50.0)
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 2300, 96, "cylindrical", "multiband", 3, seed=8)
    assert_equal("map64", st.process(imgs), want)


def test_process_from_pageable_strided_and_pinned_buffers():
    """pano_process takes whatever host memory the caller has: pageable frames with row padding (a cv::Mat ROI) go through
    the library's pinned bounce buffers and worker threads, pinned ones straight to the copy engine; the panorama may land
    in a padded pageable buffer as well.  All variants must give the same bytes as the oracle."""
    import ctypes as C
    import torch
    Ks, Rs, scale = calib.rig("2222", 320)
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 320, 180, "spherical", "multiband", 3, seed=11)
    assert_equal("contiguous pageable", st.process(imgs), want)
    lib, capi = panob200.capi.lib(), panob200.capi
    ow, oh = st.out_size
    # (a) every frame is a window of a wider, taller pageable array; the output is a window too
    big = [np.full((180 + 7, 320 + 13, 3), 201, np.uint8) for _ in imgs]
    views = []
    for b, im in zip(big, imgs):
        v = b[3:183, 5:325]
        v[...] = im
        views.append(v)
    out_big = np.full((oh + 4, ow + 9, 3), 66, np.uint8)
    out_v = out_big[2:2 + oh, 4:4 + ow]
    fp = (C.c_void_p * 4)(*[v.ctypes.data for v in views])
    sp = (C.c_int * 4)(*[v.strides[0] for v in views])
    capi.check(lib.pano_process(st._h, fp, sp, C.c_void_p(out_v.ctypes.data), out_v.strides[0]), st._h)
    assert_equal("strided pageable in / out", np.ascontiguousarray(out_v), want)
    frame = out_big.copy()
    frame[2:2 + oh, 4:4 + ow] = 66
    assert np.all(frame == 66), "bytes outside the output window were touched"
    # (b) pinned frames and a pinned panorama
    pin = [torch.from_numpy(im).pin_memory() for im in imgs]
    pout = torch.empty((oh, ow, 3), dtype=torch.uint8).pin_memory()
    fp = (C.c_void_p * 4)(*[p_.data_ptr() for p_ in pin])
    capi.check(lib.pano_process(st._h, fp, None, C.c_void_p(pout.data_ptr()), ow * 3), st._h)
    assert_equal("pinned in / out", pout.numpy(), want)
    # (c) mixed: pinned frames, pageable panorama; then the reverse
    ret = np.empty((oh, ow, 3), np.uint8)
    capi.check(lib.pano_process(st._h, fp, None, capi.ptr(ret), ow * 3), st._h)
    assert_equal("pinned in, pageable out", ret, want)
    fp2 = (C.c_void_p * 4)(*[im.ctypes.data for im in imgs])
    pout.zero_()
    capi.check(lib.pano_process(st._h, fp2, None, C.c_void_p(pout.data_ptr()), ow * 3), st._h)
    assert_equal("pageable in, pinned out", pout.numpy(), want)
    st.close()


def test_mask_update_at_runtime():
    Ks, Rs, scale = calib.rig("2222", 240)
    t, st, imgs, want = oracle_case(Ks, Rs, scale, 240, 135, "spherical", "multiband", 3, seed=4)
    assert_equal("before", st.process(imgs), want)
    new = [np.minimum(m, np.uint8(200)) for m in t.blend_masks]
    new[1][:, : new[1].shape[1] // 2] = 0
    for i, m in enumerate(new):
        st.set_mask(i, m)                       # the updateMask landing (ocvstitcher.hpp:1257)
    t.blend_masks = new
    want2 = compose.process(t, imgs, "multiband", 3)
    assert_equal("after", st.process(imgs), want2)
    assert not np.array_equal(want, want2)


def test_device_built_weight_pyramid_equals_host_builder():
    """pano_set_mask rebuilds the float weight pyramid and the collapse tile statistics ON THE DEVICE (the static part
    of MultiBandBlender::feed).  Every level must equal the host builder (pano_host_pyrdown_f32 chain on
    mask * (1/255.f)) bit for bit, and stay within float rounding of cv2's own pyramid (golden fixture)."""
    import time
    g = load("cfg1_small")
    Ks, Rs, scale = calib.rig("2222", 240)
    st = make(Ks, Rs, scale, 240, 135, num_bands=5, cut=[int(v) for v in g["cut"]])
    masks = [g["mask%d" % i] for i in range(4)]
    assert st.initTables(masks) == 0, st.last_error
    nb, _, rects = st.blend_geometry()
    for i in range(4):
        l0 = st.weight_level(i, 0)
        x0, y0 = st.m_corners[i][0] - st.dst_roi[0], st.m_corners[i][1] - st.dst_roi[1]
        pad = np.zeros(l0.shape, np.uint8)
        ox, oy = x0 - rects[i][0], y0 - rects[i][1]
        pad[oy:oy + masks[i].shape[0], ox:ox + masks[i].shape[1]] = masks[i]
        cur = pad.astype(np.float32) * np.float32(1.0 / 255.0)
        assert np.array_equal(l0, cur)
        for l in range(1, nb + 1):
            cur = panob200.capi.host_pyrdown_f32(cur)
            got = st.weight_level(i, l)
            assert got.shape == cur.shape and np.array_equal(got.view(np.uint32), cur.view(np.uint32)), "cam %d level %d" % (i, l)
            assert np.array_equal(got, g["w%d_%d" % (i, l)]), "cam %d level %d vs cv2's weight pyramid" % (i, l)
    t0 = time.perf_counter()
    for i in range(4):
        st.set_mask(i, masks[i])
    print("pano_set_mask x4 (240x135 rig): %.2f ms" % (1e3 * (time.perf_counter() - t0)))
    assert_equal("after the device mask refresh", st.process(util.synth_set(4, 135, 240, int(g["seed"]))), g["pano_multiband"])


def test_seam_mask_tail_on_device():
    """pano_set_seam_mask = dilate -> INTER_LINEAR_EXACT -> AND with the warped full mask on the device, then the compose
    with the resulting masks: mask bytes and panorama bit-exact with the oracle's restatement of the OpenCV calls."""
    Ks, Rs, scale = calib.rig("2222", 480)
    t, st, imgs, _ = oracle_case(Ks, Rs, scale, 480, 270, "spherical", "multiband", 4, seed=8)
    rng = np.random.default_rng(12)
    new = []
    for i in range(4):
        w, h = t.sizes[i]
        low = np.zeros((max(1, h // 6), max(1, w // 6)), np.uint8)
        lo, hi = sorted(rng.integers(0, low.shape[1], 2))
        low[:, lo:hi + 1] = 255
        low[rng.integers(0, low.shape[0]), :] = 0
        st.set_seam_mask(i, low)
        want = compose.seam_mask_tail(low, t.warped_masks[i])
        assert_equal("seam mask %d" % i, st.get_mask(i), want)
        new.append(want)
    t.blend_masks = new
    assert_equal("compose with device-built seam masks", st.process(imgs), compose.process(t, imgs, "multiband", 4))


def test_calibration_flow_masks_equal_cv2_init_seam():
    """ocvStitcher.calibration() (initMode 2: fixed parameters -> initSeam: GraphCut on low-res warps on the host, the
    mask tail on the device) yields the masks of the reference call sequence replayed through cv2 (:975-1101)."""
    pytest.importorskip("cv2")
    from oracle import cv2_reference as ref
    Ks, Rs, scale = calib.rig("2222", 480)
    imgs = util.synth_set(4, 270, 480, 21)
    t = ref.init_seam(imgs, Ks, Rs, scale, warp="spherical", seam="gc_color")
    st = panob200.ocvStitcher(SC(width=480, height=270, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                 blender="multiband", num_bands=4, initMode=2))
    assert st.calibration(imgs) == 0, st.last_error
    for i in range(4):
        assert_equal("m_blenderMask[%d]" % i, st.get_mask(i), t.blend_masks[i])
    assert_equal("panorama after calibration()", st.process(imgs), ref.process(t, imgs, "multiband", 4))


def test_batched_device_and_host_apis():
    import torch
    Ks, Rs, scale = calib.rig("2222", 240)
    t = compose.build_tables(Ks, Rs, scale, (240, 135), "spherical")
    t.blend_masks = util.soft_masks(t)
    st = panob200.ocvStitcher(SC(width=240, height=135, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                 blender="multiband", num_bands=4, cut_height=100, max_batch=2))
    assert st.initTables(t.blend_masks) == 0, st.last_error
    B = 5
    sets = [util.synth_set(4, 135, 240, 50 + b) for b in range(B)]
    cut = st.m_cutParams
    want = [compose.process(t, s, "multiband", 4, cut=cut) for s in sets]
    host = torch.from_numpy(np.stack([np.stack(s) for s in sets])).pin_memory()
    ow, oh = st.out_size
    out_dev = torch.empty((B, oh, ow, 3), dtype=torch.uint8, device="cuda")
    st.process_device(host.cuda(), out_dev)
    torch.cuda.synchronize()
    got = out_dev.cpu().numpy()
    for b in range(B):
        assert_equal("device batch %d" % b, got[b], want[b])
    out_host = torch.empty((B, oh, ow, 3), dtype=torch.uint8).pin_memory()
    st.process_batch(host, out_host)
    for b in range(B):
        assert_equal("host batch %d" % b, out_host.numpy()[b], want[b])
    # idempotence: same inputs, same bytes; and strided host frames through pano_process
    st.process_device(host.cuda(), out_dev)
    torch.cuda.synchronize()
    assert np.array_equal(out_dev.cpu().numpy(), got)
    padded = [np.ascontiguousarray(np.pad(f, ((0, 0), (0, 5), (0, 0))))[:, :240] for f in sets[0]]
    assert_equal("strided", st.process(padded), want[0])


def test_gain_map_rounding_edge_cases():
    """BlocksGainCompensator::apply is sat_u8(cvRound(v * g)): gains that produce exact ties (x.5 -> even), overflow,
    zero, negative, infinite and NaN products must round / saturate like the oracle (the warp kernel evaluates it
    with mantissa arithmetic instead of I2F / F2I).  OpenCV quirk included: saturate_cast<uchar>(cvRound(x)) is 0, not
    255, once x leaves the int range (cvtss2si returns INT_MIN) -- verified against cv2.multiply."""
    Ks, Rs, scale = calib.rig("2222", 320)
    special = np.array([0.5, 1.5, 2.5, 0.25, 0.75, 300.0, 0.0, -1.0, np.inf, np.nan, 1.0, 0.9999999, 1.0000001, 127.5 / 255.0,
                        254.5 / 255.0, 3.0e38, 2.2e7, -3.0e38, 8421504.5, 1.0e9 / 255.0], np.float32)
    for blender, nb in (("multiband", 3), ("feather", 0)):
        t = compose.build_tables(Ks, Rs, scale, (320, 180), "spherical")
        t.blend_masks = util.soft_masks(t)
        gains = []
        for i, (w, h) in enumerate(t.sizes):
            yy, xx = np.mgrid[0:h, 0:w]
            gains.append(special[(xx + 3 * yy + i) % len(special)].astype(np.float32))
        imgs = util.synth_set(4, 180, 320, 55)
        imgs[0][:16, :16] = np.arange(256, dtype=np.uint8).reshape(16, 16)[:, :, None]      # every 8-bit value
        st = make(Ks, Rs, scale, 320, 180, "spherical", blender, nb, sharp=0.05)
        assert st.initTables(t.blend_masks) == 0, st.last_error
        st.set_gain_maps(gains)
        t.gain_maps = gains
        fw = [panob200.capi.host_feather_weight(m, 0.05) for m in t.blend_masks] if blender == "feather" else None
        assert_equal("gain edge cases " + blender, st.process(imgs), compose.process(t, imgs, blender, nb, feather_weights=fw))


def test_error_behaviour():
    Ks, Rs, scale = calib.rig("2222", 240)
    st = make(Ks, Rs, scale, 240, 135, num_bands=3, cut=[0, 0, 5000, 50])
    assert st.initTables() == -1 and "cut" in st.last_error          # cv::Mat ROI would assert (:1210)
    st = make(Ks, Rs, scale, 240, 135, num_bands=3)
    assert st.initTables() == 0
    with pytest.raises(panob200.PanoError):
        st.set_mask(0, np.zeros((10, 10), np.uint8))
    with pytest.raises(panob200.PanoError):
        st.process([np.zeros((135, 240, 3), np.uint8)] * 3)
    with pytest.raises(panob200.PanoError):
        st.process([np.zeros((100, 240, 3), np.uint8)] * 4)


def test_error_behaviour_of_the_widened_entry_points():
    """Every new entry point returns PANO_ERR with a message instead of crashing on bad input."""
    import ctypes as C
    capi = panob200.capi
    lib = capi.lib()
    Ks, Rs, scale = calib.rig("2222", 240)
    st = make(Ks, Rs, scale, 240, 135, num_bands=3)
    assert st.initTables() == 0
    with pytest.raises(panob200.PanoError, match="attach a front end first"):
        st.set_frontend_mode(True)                                   # fused mode needs a front end
    st.set_frontend_mode(False)                                      # no-op without one
    with pytest.raises(panob200.PanoError):
        st.set_seam_mask(7, np.zeros((4, 4), np.uint8))              # bad camera index
    assert lib.pano_set_seam_mask(st._h, 0, None, 4, 4, 4) == capi.PANO_ERR
    assert lib.pano_get_weight_level(st._h, 0, 9, capi.ptr(np.zeros(4, np.float32))) == capi.PANO_ERR   # level > num_bands
    assert lib.pano_get_mask(st._h, 0, capi.ptr(np.zeros(4, np.uint8)), 1) == capi.PANO_ERR              # stride < width
    # front end: odd YUYV width, bad format; wrong frame size at attach
    CamConfig = panob200.pkg.nvcam.CamConfig
    with pytest.raises(panob200.PanoError, match="even width"):
        panob200.nvCamFrontEnd(CamConfig(camSrcWidth=241, camSrcHeight=135, undistoredWidth=241, undistoredHeight=135,
                                         outPutWidth=240, outPutHeight=135, undistor=False, srcFormat="yuyv"))
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=320, camSrcHeight=180, undistoredWidth=320, undistoredHeight=180,
                                          outPutWidth=320, outPutHeight=180, undistor=False))
    with pytest.raises(panob200.PanoError, match="delivers 320x180"):
        st.attach_frontend(fe)
    # ring epilogue
    rc = capi.pano_ring_config()
    rc.up_width, rc.up_height, rc.down_width, rc.down_height, rc.mode, rc.finalcut, rc.bar = 100, 20, 90, 24, 1, 10, 4
    h = C.c_void_p()
    assert lib.pano_ring_create(C.byref(rc), C.byref(h)) == capi.PANO_ERR and b"finalcut" in lib.pano_ring_last_error(None)
    rc.mode = 5
    assert lib.pano_ring_create(C.byref(rc), C.byref(h)) == capi.PANO_ERR
    r = panob200.RingComposer((100, 20), (90, 24), "resize")
    with pytest.raises(panob200.PanoError, match="stride"):
        r._check(lib.pano_ring_compose(r._h, capi.ptr(np.zeros((20, 100, 3), np.uint8)), 10, capi.ptr(np.zeros((24, 90, 3), np.uint8)), 270,
                                       capi.ptr(np.zeros((48, 90, 3), np.uint8)), 270))


# ------------------------------------------------------------------ BASELINE full sizes

def test_config1_full_size_vs_oracle_and_properties():
    """4 x 1920x1080 spherical, 5 bands (BASELINE config 1 shape) against the scalar oracle, plus
    size-independent properties: batch slots are independent and deterministic."""
    import torch
    Ks, Rs, scale = calib.rig("2222", 1920)
    t = compose.build_tables(Ks, Rs, scale, (1920, 1080), "spherical")
    t.blend_masks = util.soft_masks(t)
    cut = [0, 64, 5336, 896]
    st = panob200.ocvStitcher(SC(width=1920, height=1080, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                 blender="multiband", num_bands=5, cut=cut, max_batch=2))
    assert st.initTables(t.blend_masks) == 0, st.last_error
    assert st.dst_roi == (-4486, 1903, 5336, 1025)
    imgs = util.synth_set(4, 1080, 1920, 7)
    want = compose.process(t, imgs, "multiband", 5, cut=cut)
    got = st.process(imgs)
    assert_equal("config1 full", got, want)
    other = util.synth_set(4, 1080, 1920, 8)
    batch = torch.from_numpy(np.stack([np.stack(imgs), np.stack(other), np.stack(imgs)])).cuda()
    out = torch.empty((3, 896, 5336, 3), dtype=torch.uint8, device="cuda")
    st.process_device(batch, out)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.array_equal(o[0], got) and np.array_equal(o[2], got) and not np.array_equal(o[1], got)
    # uncovered panorama pixels read as 0 (the reference zero-fills dst every frame)
    full, mask = compose.process(t, imgs, "multiband", 5, return_s16=True)
    hole = mask[64:960] == 0
    assert (got[hole] == 0).all()


def test_config3_shape_gain_feather_and_no_blend():
    """BASELINE config 3: imx424 rig x3 (1920x1080), warp -> block-gain apply on 8U -> 16S ->
    FeatherBlender(sharpness = 1/blend_width); plus Blender::NO (what the yaml's strength 0 selects)."""
    Ks, Rs, scale = calib.rig("424", 1920)
    t = compose.build_tables(Ks, Rs, scale, (1920, 1080), "spherical")
    t.blend_masks = util.soft_masks(t)
    assert t.dst_roi[2:] == (6383, 1131)                       # SURVEY 8d geometry
    bw = compose.blend_width(t.dst_roi, 5.0)
    sharp = float(np.float32(1.0) / np.float32(bw))
    gains = []
    for i, (w, h) in enumerate(t.sizes):                       # smooth per-pixel gain tables (static after init)
        yy, xx = np.mgrid[0:h, 0:w]
        gains.append((0.85 + 0.3 * ((xx * (i + 3) + yy * 7) % 997) / 997.0).astype(np.float32))
    imgs = util.synth_set(4, 1080, 1920, 31)
    fw = [panob200.capi.host_feather_weight(m, sharp) for m in t.blend_masks]
    st = make(Ks, Rs, scale, 1920, 1080, "spherical", "feather", 0, sharp=sharp)
    assert st.initTables(t.blend_masks) == 0, st.last_error
    st.set_gain_maps(gains)
    t.gain_maps = gains
    assert_equal("config3 gain+feather", st.process(imgs), compose.process(t, imgs, "feather", feather_weights=fw))
    st = make(Ks, Rs, scale, 1920, 1080, "spherical", "no", 0)
    assert st.initTables(t.blend_masks) == 0, st.last_error
    t.gain_maps = None
    assert_equal("config3 no-blend", st.process(imgs), compose.process(t, imgs, "no"))


def test_config4_full_size_8cam_4k_cylindrical_7_bands():
    """BASELINE config 4 on ONE GPU: 8 x 3840x2160, cylindrical ring, 7 bands, 64-bit map entries."""
    Ks, Rs, scale = calib.ring(8, 3840, 2160, 65.2, 40.0, focal=3000.0)
    t = compose.build_tables(Ks, Rs, scale, (3840, 2160), "cylindrical")
    t.blend_masks = util.soft_masks(t)
    assert t.dst_roi[2:] == (18076, 2160) and t.corners[0] == (-9038, -1080)
    st = make(Ks, Rs, scale, 3840, 2160, "cylindrical", "multiband", 7)
    assert st.initTables(t.blend_masks) == 0, st.last_error
    nb, padded, rects = st.blend_geometry()
    assert nb == 7 and padded == (18176, 2176)
    imgs = [util.synth_frame(2160, 3840, 400 + i, cell=64) for i in range(8)]
    got = st.process(imgs)
    assert_equal("config4 full", got, compose.process(t, imgs, "multiband", 7))
    assert np.array_equal(st.process(imgs), got)               # idempotent


def test_config2_front_end_chained_into_process():
    """BASELINE config 2 at half scale: 8UC4 camera frames -> cubic undistort -> crop -> resize ->
    spherical warp -> 4-band blend, as ONE call (pano_attach_frontend), vs the oracle's sequential
    undistort -> warp interpolation order."""
    import torch
    cam = calib.CAM_LIJING_390_FOV60_1920
    s = 0.5
    K = np.array(cam["K"], np.float64).reshape(3, 3).copy()
    K[0, 0] *= s; K[0, 2] *= s; K[1, 1] *= s; K[1, 2] *= s
    newK = np.array([[1627.5076 * s, 0, 943.1681 * s], [0, 1622.9720 * s, 571.5369 * s], [0, 0, 1]])
    rect = [34, 52, 891, 444]
    CamConfig = panob200.pkg.nvcam.CamConfig
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=960, camSrcHeight=540, undistoredWidth=960, undistoredHeight=540,
                                          outPutWidth=960, outPutHeight=540, K=K.reshape(-1), distorParams=cam["distorParams"],
                                          rect=rect, newK=newK, max_batch=8))
    mx, my = fe.maps()
    Ks, Rs, scale = calib.rig("2222", 960)
    t = compose.build_tables(Ks, Rs, scale, (960, 540), "spherical")
    t.blend_masks = util.soft_masks(t)
    st = panob200.ocvStitcher(SC(width=960, height=540, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                 blender="multiband", num_bands=4, max_batch=2))
    assert st.initTables(t.blend_masks) == 0, st.last_error
    st.attach_frontend(fe)
    sets = [[util.synth_frame(540, 960, 900 + 10 * b + i, channels=4) for i in range(4)] for b in range(3)]
    want = []
    for frames in sets:
        bgr = [compose.front_end(f, (960, 540), mx, my, rect, (960, 540)) for f in frames]
        want.append(compose.process(t, bgr, "multiband", 4))
    assert_equal("config2 host call", st.process(sets[0]), want[0])
    dev = torch.from_numpy(np.stack([np.stack(f) for f in sets])).cuda()
    ow, oh = st.out_size
    out = torch.empty((3, oh, ow, 3), dtype=torch.uint8, device="cuda")
    st.process_device(dev, out)
    torch.cuda.synchronize()
    for b in range(3):
        assert_equal("config2 device batch %d" % b, out[b].cpu().numpy(), want[b])
    host_in = torch.from_numpy(np.stack([np.stack(f) for f in sets])).pin_memory()
    host_out = torch.empty((3, oh, ow, 3), dtype=torch.uint8).pin_memory()
    st.process_batch(host_in, host_out)
    for b in range(3):
        assert_equal("config2 host batch %d" % b, host_out[b].numpy(), want[b])


def test_config2_fused_single_gather_variant():
    """The single-gather ("fused map") variant of config 2 (pano_set_frontend_mode): the library's composed maps
    match an independent float64 restatement, the compose on those maps is bit-exact with the oracle gathering
    from the raw camera frame, its PSNR against the sequential (parity) result is reported and bounded, and
    switching back restores the bit-exact sequential path."""
    import copy
    cam = calib.CAM_LIJING_390_FOV60_1920
    s = 0.5
    K = np.array(cam["K"], np.float64).reshape(3, 3).copy()
    K[0, 0] *= s; K[0, 2] *= s; K[1, 1] *= s; K[1, 2] *= s
    newK = np.array([[1627.5076 * s, 0, 943.1681 * s], [0, 1622.9720 * s, 571.5369 * s], [0, 0, 1]])
    rect = [34, 52, 891, 444]
    CamConfig = panob200.pkg.nvcam.CamConfig
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=960, camSrcHeight=540, undistoredWidth=960, undistoredHeight=540,
                                          outPutWidth=960, outPutHeight=540, K=K.reshape(-1), distorParams=cam["distorParams"],
                                          rect=rect, newK=newK, max_batch=4))
    mx, my = fe.maps()
    Ks, Rs, scale = calib.rig("2222", 960)
    t = compose.build_tables(Ks, Rs, scale, (960, 540), "spherical")
    t.blend_masks = util.soft_masks(t)
    frames = [util.synth_frame(540, 960, 950 + i, channels=4) for i in range(4)]
    raw_bgr = [np.ascontiguousarray(f[:, :, :3]) for f in frames]
    seq_bgr = [compose.front_end(f, (960, 540), mx, my, rect, (960, 540)) for f in frames]
    for blender, nb in (("multiband", 4), ("feather", 0)):
        st = panob200.ocvStitcher(SC(width=960, height=540, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                     blender=blender, num_bands=nb, sharpness=0.05, max_batch=2))
        assert st.initTables(t.blend_masks) == 0, st.last_error
        fw = [panob200.capi.host_feather_weight(m, 0.05) for m in t.blend_masks] if blender == "feather" else None
        st.attach_frontend(fe)
        want_seq = compose.process(t, seq_bgr, blender, nb, feather_weights=fw)
        assert_equal("sequential " + blender, st.process(frames), want_seq)
        st.set_frontend_mode(True)
        tf = copy.copy(t)
        tf.maps = []
        for i in range(4):
            fx, fy = st.warp_maps(i)
            wx, wy = compose.fused_front_end_maps(t.maps[i][0], t.maps[i][1], (960, 540), (960, 540), mx, my, rect, (960, 540))
            assert np.abs(fx - wx).max() < 1e-3 and np.abs(fy - wy).max() < 1e-3, "composed map of camera %d" % i
            tf.maps.append((fx, fy))
        got = st.process(frames)
        assert_equal("fused " + blender, got, compose.process(tf, raw_bgr, blender, nb, feather_weights=fw))
        p = util.psnr(got, want_seq)
        print("fused single-gather vs sequential (%s): %s" % (blender, util.report("fused", got, want_seq)))
        assert p > 30.0
        st.set_frontend_mode(False)
        assert_equal("sequential again " + blender, st.process(frames), want_seq)
        st.close()


# ------------------------------------------------------------------ column-strip split (config 4)

def _strip_setup(nranks, mode, W=960, H=540, nb=5, ncam=8):
    import torch
    Ks, Rs, scale = calib.ring(ncam, W, H, 65.2, 40.0)
    t = compose.build_tables(Ks, Rs, scale, (W, H), "cylindrical")
    t.blend_masks = util.soft_masks(t)
    imgs = [util.synth_frame(H, W, 700 + i, cell=32) for i in range(ncam)]
    ref_st = make(Ks, Rs, scale, W, H, "cylindrical", "multiband", nb)
    assert ref_st.initTables(t.blend_masks) == 0, ref_st.last_error
    want = ref_st.process(imgs)
    frames = torch.from_numpy(np.stack(imgs)).cuda()
    ranks, panos = [], []
    for r in range(nranks):
        st = make(Ks, Rs, scale, W, H, "cylindrical", "multiband", nb)
        assert st.initTables(t.blend_masks) == 0, st.last_error
        ranks.append(panob200.strips.StripRank(st, r, nranks, mode))
        panos.append(torch.full(want.shape, 77, dtype=torch.uint8, device="cuda"))
    return t, imgs, want, frames, ranks, panos


@pytest.mark.parametrize("nranks,mode", [(2, "exchange"), (4, "exchange"), (8, "exchange"), (4, "redundant")])
def test_strip_split_equals_single_gpu(nranks, mode):
    """Ranks emulated in lockstep on one GPU: the strip-split panorama (halo exchange or redundant
    halo) must equal the undivided one byte for byte -- and that one equals the oracle."""
    import torch
    t, imgs, want, frames, ranks, panos = _strip_setup(nranks, mode)
    assert_equal("undivided vs oracle", want, compose.process(t, imgs, "multiband", 5))
    panob200.strips.compose_local(ranks, frames, panos)
    torch.cuda.synchronize()
    got = panob200.strips.assemble(ranks, panos).cpu().numpy()
    assert_equal("strip split %d/%s" % (nranks, mode), got, want)
    if mode == "exchange":
        assert sum(ranks[0].halo_bytes(p) for p in range(ranks[0].phases)) > 0
        # a rank that skipped the exchange must NOT reproduce the result (the halo matters)
        for p_, r in zip(panos, ranks):
            p_.fill_(0)
        for r, p_ in zip(ranks, panos):
            for ph in range(r.phases):
                r.run_phase(ph, frames, p_, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert not np.array_equal(panob200.strips.assemble(ranks, panos).cpu().numpy(), want)


@pytest.mark.parametrize("nranks,split", [(2, 3), (4, 3), (8, 2), (8, 4), (3, 1), (4, 5)])
def test_strip_split_hybrid_equals_single_gpu(nranks, split):
    """Hybrid decomposition (thin redundant halo below `split`, full width above, one all-gather of g[split]): ranks
    emulated in lockstep on one GPU must reproduce the undivided panorama byte for byte, for split levels from 1 to nb."""
    import torch
    Ks, Rs, scale = calib.ring(8, 960, 540, 65.2, 40.0)
    t = compose.build_tables(Ks, Rs, scale, (960, 540), "cylindrical")
    t.blend_masks = util.soft_masks(t)
    imgs = [util.synth_frame(540, 960, 700 + i, cell=32) for i in range(8)]
    ref_st = make(Ks, Rs, scale, 960, 540, "cylindrical", "multiband", 5)
    assert ref_st.initTables(t.blend_masks) == 0, ref_st.last_error
    want = ref_st.process(imgs)
    frames = torch.from_numpy(np.stack(imgs)).cuda()
    ranks, panos = [], []
    for r in range(nranks):
        st = make(Ks, Rs, scale, 960, 540, "cylindrical", "multiband", 5)
        assert st.initTables(t.blend_masks) == 0, st.last_error
        ranks.append(panob200.strips.StripRank(st, r, nranks, "hybrid", split=split))
        panos.append(torch.full(want.shape, 77, dtype=torch.uint8, device="cuda"))
    for _ in range(2):                                       # twice: stale data of the first frame must not matter
        panob200.strips.compose_hybrid_local(ranks, frames, panos)
    torch.cuda.synchronize()
    got = panob200.strips.assemble(ranks, panos).cpu().numpy()
    assert_equal("hybrid strip split %d ranks, split level %d" % (nranks, split), got, want)


@pytest.mark.parametrize("nranks,mode", [(4, "exchange"), (8, "redundant"), (8, "hybrid")])
def test_strip_rank_reads_only_its_cameras(nranks, mode):
    """pano_strip_cameras: a rank only needs the frames of the cameras whose warped ROI meets its window.  Every rank gets
    its own frame buffer in which all OTHER cameras hold noise; the assembled panorama must still equal the undivided one,
    and with 8 ranks no rank may need all 8 cameras of the ring."""
    import torch
    t, imgs, want, frames, ranks, panos = _strip_setup(nranks, mode)
    g = torch.Generator(device="cuda").manual_seed(5)
    per_rank = []
    for r in ranks:
        need = r.cameras()
        assert need and all(0 <= i < 8 for i in need)
        if nranks == 8:
            assert len(need) < 8, "rank %d claims every camera: %r" % (r.rank, need)
        fr = torch.randint(0, 256, frames.shape, dtype=torch.uint8, device="cuda", generator=g)
        for i in need:
            fr[i] = frames[i]
        per_rank.append(fr)
    if mode == "hybrid":
        panob200.strips.compose_hybrid_local(ranks, per_rank, panos)
    else:
        panob200.strips.compose_local(ranks, per_rank, panos)
    torch.cuda.synchronize()
    got = panob200.strips.assemble(ranks, panos).cpu().numpy()
    assert_equal("strip split with per-rank cameras %d/%s" % (nranks, mode), got, want)
    # and the list is tight: noise in a NEEDED camera changes that rank's columns
    r = ranks[nranks // 2]
    bad = per_rank[r.rank].clone()
    mid = r.cameras()[len(r.cameras()) // 2]
    bad[mid] = 255 - bad[mid]
    if mode == "redundant":
        p2 = torch.zeros_like(panos[0])
        for ph in range(r.phases):
            r.run_phase(ph, bad, p2, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        c0, c1 = r.own_output_columns()
        assert not torch.equal(p2[:, c0:c1], panos[r.rank][:, c0:c1])


@pytest.mark.parametrize("nranks,concurrent", [(2, False), (4, False), (8, False)])
def test_strip_split_peer_memory_exchange(nranks, concurrent):
    """Halo exchange through peer-memory mailboxes (push kernel stores into the neighbour's mailbox and raises a flag,
    wait/unpack kernel consumes it).  Ranks are handles on ONE GPU stepped in lockstep on one stream: every push
    precedes the matching wait, so no kernel ever spins (kernels that wait on one another must not share a GPU --
    B200_PROFILING.md; the truly concurrent exchange is covered on real GPUs by test_strip_split_multi_gpu_torchrun).
    Three consecutive frame-sets (sequence numbers, slot parity) must each equal the undivided panorama byte for byte."""
    import torch
    t, imgs, want, frames, ranks, panos = _strip_setup(nranks, "exchange")
    panob200.strips.p2p_setup_local(ranks)
    ref_st = make(*calib.ring(8, 960, 540, 65.2, 40.0), 960, 540, "cylindrical", "multiband", 5)
    assert ref_st.initTables(t.blend_masks) == 0, ref_st.last_error
    variants = [frames, torch.flip(frames, dims=[2]).contiguous(), frames]
    for k, fr in enumerate(variants):
        for p_ in panos:
            p_.fill_(77)
        panob200.strips.compose_p2p_local(ranks, fr, panos, concurrent=concurrent)
        torch.cuda.synchronize()
        got = panob200.strips.assemble(ranks, panos).cpu().numpy()
        ref = want if k != 1 else ref_st.process([f for f in fr.cpu().numpy()])
        assert_equal("p2p strip split %d ranks, frame %d" % (nranks, k), got, ref)


def test_strip_split_multi_gpu_torchrun():
    """The strip split on REAL GPUs, one process per GPU under torchrun (skipped with fewer than two devices): NCCL
    point-to-point halos, peer-memory mailboxes (kernels of different GPUs really wait on one another here) and the
    redundant halo must all reproduce the undivided panorama on every rank."""
    import json
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(util.ROOT, "tools", "strip_split_nccl.py"), "--small", "--steps", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["n_gpus"] == world
    for mode in ("exchange", "p2p", "redundant", "hybrid"):
        assert rec["modes"][mode]["all_ranks_match_undivided"], (mode, rec)


# ------------------------------------------------------------------ nvCam front end

def test_front_end_golden_bit_exact():
    g = np.load(os.path.join(util.GOLDEN, "frontend_small.npz"))
    CamConfig = panob200.pkg.nvcam.CamConfig
    rect = [int(v) for v in g["rect"]]
    base = dict(K=g["K"].reshape(-1), distorParams=g["D"], rect=rect, newK=g["newK"])
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=480, camSrcHeight=270, undistoredWidth=480, undistoredHeight=270,
                                          outPutWidth=480, outPutHeight=270, undistor=True, **base))
    mx, my = fe.maps()
    ixy, fr = orc.convert_maps(mx, my)
    omx, omy = orc.init_undistort_map(g["K"], g["D"], g["newK"], 480, 270)
    oixy, ofr = orc.convert_maps(omx, omy)
    assert np.array_equal(ixy, oixy) and np.array_equal(fr, ofr)
    assert_equal("same size", fe.getFrame(util.synth_frame(270, 480, 77, channels=4)), g["out_same"])
    big = util.synth_frame(540, 960, 78, channels=4)
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=960, camSrcHeight=540, undistoredWidth=480, undistoredHeight=270,
                                          outPutWidth=360, outPutHeight=203, undistor=True, **base))
    assert_equal("down", fe.getFrame(big), g["out_down"])
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=960, camSrcHeight=540, undistoredWidth=480, undistoredHeight=270,
                                          outPutWidth=360, outPutHeight=203, undistor=False, **base))
    assert_equal("no undistort", fe.getFrame(big), g["out_noud"])


def test_front_end_config2_shape_vs_oracle():
    """1920x1080 8UC4 -> cubic undistort -> crop [69,103,1782,889] -> resize 1920x1080 (config 2)."""
    import torch
    cam = calib.CAM_LIJING_390_FOV60_1920
    newK = np.array([[1627.5076, 0, 943.1681], [0, 1622.9720, 571.5369], [0, 0, 1]])
    CamConfig = panob200.pkg.nvcam.CamConfig
    fe = panob200.nvCamFrontEnd(CamConfig(K=cam["K"], distorParams=cam["distorParams"], rect=cam["rect"], newK=newK, max_batch=2))
    mx, my = fe.maps()
    frames = [util.synth_frame(1080, 1920, 90 + i, channels=4) for i in range(3)]
    want = [compose.front_end(f, (1920, 1080), mx, my, cam["rect"], (1920, 1080)) for f in frames]
    assert_equal("config2 front end", fe.getFrame(frames[0]), want[0])
    dev = torch.from_numpy(np.stack(frames)).cuda()
    out = torch.empty((3, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
    fe.process_device(dev, out)
    torch.cuda.synchronize()
    for i in range(3):
        assert_equal("batch %d" % i, out[i].cpu().numpy(), want[i])


# ------------------------------------------------------------------ two-ring epilogue

@pytest.mark.parametrize("mode,up,down,fc", [("resize", (333, 57), (301, 64), 0), ("resize", (200, 40), (200, 40), 0),
                                              ("resize", (5336, 896), (5000, 880), 0), ("crop", (333, 57), (301, 64), 3),
                                              ("crop", (150, 33), (211, 30), 0)])
def test_ring_epilogue_bit_exact(mode, up, down, fc):
    """resize + vconcat + separator bar of src/master.cpp:321-326 / src/panocamimpl.cpp:354-360 as one kernel."""
    import torch
    rng = np.random.default_rng(31)
    u = rng.integers(0, 256, (2, up[1], up[0], 3), np.uint8)
    d = rng.integers(0, 256, (2, down[1], down[0], 3), np.uint8)
    rc = panob200.RingComposer(up, down, mode, finalcut=fc)
    want = [compose.ring_epilogue(u[b], d[b], mode, finalcut=fc) for b in range(2)]
    assert rc.out_size == (want[0].shape[1], want[0].shape[0])
    assert_equal("ring host", rc.compose(u[0], d[0]), want[0])
    out = torch.empty((2,) + want[0].shape, dtype=torch.uint8, device="cuda")
    rc.compose_device(torch.from_numpy(u).cuda(), torch.from_numpy(d).cuda(), out)
    torch.cuda.synchronize()
    for b in range(2):
        assert_equal("ring device %d" % b, out[b].cpu().numpy(), want[b])
    rc.close()


@pytest.mark.parametrize("size", [(5336, 1792), (2500, 700), (3840, 1000), (1920, 1080), (1000, 500), (3840, 2160), (2007, 333)])
def test_fit2final_bit_exact(size):
    """nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189): scale by fitscale + centred paste on the black canvas as one
    kernel, against the oracle restatement (itself pinned against the cv2 calls in tests/test_oracle_vs_cv2.py)."""
    import torch
    w, h = size
    frames = np.stack([util.synth_frame(h, w, 60 + b) for b in range(2)])
    fc = panob200.FitCanvas((w, h))
    want = [compose.fit2final(frames[b]) for b in range(2)]
    assert_equal("fit2final host", fc.fit2final(frames[0]), want[0])
    out = torch.empty((2, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
    fc.fit2final_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    for b in range(2):
        assert_equal("fit2final device %d" % b, out[b].cpu().numpy(), want[b])
    fc.close()
    with pytest.raises(panob200.PanoError):
        panob200.FitCanvas((1000, 1200))                 # taller than the canvas: the reference's ROI would assert


# ------------------------------------------------------------------ YUYV ingest

def _yuyv_frame(h, w, seed):
    """Valid 8UC2 YUYV bytes from a synthetic 3-channel frame: Y = channel 0, U / V = channels 1 / 2 of the even pixel."""
    f = util.synth_frame(h, w, seed, channels=3)
    y = np.empty((h, w, 2), np.uint8)
    y[:, :, 0] = f[:, :, 0]
    y[:, 0::2, 1] = f[:, 0::2, 1]
    y[:, 1::2, 1] = f[:, 0::2, 2]
    y[0, :6] = [[0, 0], [0, 0], [255, 255], [255, 255], [0, 255], [255, 0]]
    return y


@pytest.mark.parametrize("fused", [False, True])
def test_yuyv_ingest_chained_into_process(fused):
    """8UC2 YUYV camera frames in (half the bytes of 8UC4): cvtColor(COLOR_YUV2BGRA_YUYV) + nvCam front end + compose as
    one call, bit-exact with the oracle's sequential restatement (fused=True: the single-gather variant on the
    converted frames, bit-exact with the oracle gathering through the library's composed maps)."""
    import copy
    import torch
    cam = calib.CAM_LIJING_390_FOV60_1920
    s = 0.5
    K = np.array(cam["K"], np.float64).reshape(3, 3).copy()
    K[0, 0] *= s; K[0, 2] *= s; K[1, 1] *= s; K[1, 2] *= s
    newK = np.array([[1627.5076 * s, 0, 943.1681 * s], [0, 1622.9720 * s, 571.5369 * s], [0, 0, 1]])
    rect = [34, 52, 891, 444]
    CamConfig = panob200.pkg.nvcam.CamConfig
    fe = panob200.nvCamFrontEnd(CamConfig(camSrcWidth=960, camSrcHeight=540, undistoredWidth=960, undistoredHeight=540,
                                          outPutWidth=960, outPutHeight=540, K=K.reshape(-1), distorParams=cam["distorParams"],
                                          rect=rect, newK=newK, max_batch=8, srcFormat="yuyv"))
    mx, my = fe.maps()
    sets = [[_yuyv_frame(540, 960, 1200 + 10 * b + i) for i in range(4)] for b in range(2)]
    bgra = [[compose.yuyv_to_bgra(f) for f in fs] for fs in sets]
    want_fe = compose.front_end(bgra[0][0], (960, 540), mx, my, rect, (960, 540))
    assert_equal("yuyv front end alone", fe.getFrame(sets[0][0]), want_fe)
    Ks, Rs, scale = calib.rig("2222", 960)
    t = compose.build_tables(Ks, Rs, scale, (960, 540), "spherical")
    t.blend_masks = util.soft_masks(t)
    st = panob200.ocvStitcher(SC(width=960, height=540, num_images=4, Ks=Ks, Rs=Rs, warped_image_scale=scale,
                                 blender="multiband", num_bands=4, max_batch=2))
    assert st.initTables(t.blend_masks) == 0, st.last_error
    st.attach_frontend(fe)
    tt = t
    if fused:
        st.set_frontend_mode(True)
        tt = copy.copy(t)
        tt.maps = [st.warp_maps(i) for i in range(4)]
    want = []
    for fs in bgra:
        src = [np.ascontiguousarray(f[:, :, :3]) for f in fs] if fused else [compose.front_end(f, (960, 540), mx, my, rect, (960, 540)) for f in fs]
        want.append(compose.process(tt, src, "multiband", 4))
    assert_equal("yuyv host call", st.process(sets[0]), want[0])
    dev = torch.from_numpy(np.stack([np.stack(f) for f in sets])).cuda()
    ow, oh = st.out_size
    out = torch.empty((2, oh, ow, 3), dtype=torch.uint8, device="cuda")
    st.process_device(dev, out)
    torch.cuda.synchronize()
    host_in = torch.from_numpy(np.stack([np.stack(f) for f in sets])).pin_memory()
    host_out = torch.empty((2, oh, ow, 3), dtype=torch.uint8).pin_memory()
    st.process_batch(host_in, host_out)
    for b in range(2):
        assert_equal("yuyv device batch %d" % b, out[b].cpu().numpy(), want[b])
        assert_equal("yuyv host batch %d" % b, host_out[b].numpy(), want[b])


def test_ring_epilogue_against_cv2_fixture():
    """The two-ring kernel against cv2's own resize + vconcat + rectangle output (tests/golden/widened_small.npz)."""
    g = load("widened_small")
    up, down = g["ring_up"], g["ring_down"]
    rc = panob200.RingComposer((up.shape[1], up.shape[0]), (down.shape[1], down.shape[0]), "resize")
    assert_equal("ring resize vs cv2", rc.compose(up, down), g["ring_resize"])
    rc.close()
    rc = panob200.RingComposer((up.shape[1], up.shape[0]), (down.shape[1], down.shape[0]), "crop", finalcut=3)
    assert_equal("ring crop vs cv2", rc.compose(up, down), g["ring_crop"])
    rc.close()
