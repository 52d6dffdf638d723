"""The oracle (oracle/pano_oracle.c) against the committed golden vectors produced by cv2 4.13
(tests/golden/make_golden.py).  No cv2 and no GPU needed: this is what pins the checker."""
import hashlib
import os

import numpy as np
import pytest

import util
from golden import calib
from oracle import compose, oracle as orc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    return np.load(os.path.join(util.GOLDEN, name + ".npz"))


def golden_tables(g, rig, warp):
    W, H = int(g["width"]), int(g["height"])
    Ks, Rs, scale = calib.rig(rig, W)
    t = compose.build_tables(Ks, Rs, scale, (W, H), warp)
    n = len(Rs)
    t.blend_masks = [g["mask%d" % i] for i in range(n)]
    return t, util.synth_set(n, H, W, int(g["seed"]))


def ext_weights(g, t, nb):
    n = len(t.sizes)
    nbe, pwh = orc.mb_prepare(t.dst_roi, nb)
    out = []
    for i in range(n):
        _, bd = orc.mb_feed_rect(t.dst_roi, pwh, nbe, t.corners[i], t.sizes[i])
        m = t.blend_masks[i]
        w0 = np.zeros((m.shape[0] + bd[0] + bd[1], m.shape[1] + bd[2] + bd[3]), np.float32)
        w0[bd[0]:bd[0] + m.shape[0], bd[2]:bd[2] + m.shape[1]] = m.astype(np.float32) * np.float32(1.0 / 255.0)
        out.append([w0] + [g["w%d_%d" % (i, l)] for l in range(1, nbe + 1)])
    return out


def test_primitive_known_answers():
    want = dict(l.split() for l in open(os.path.join(util.GOLDEN, "primitive_hashes.txt")))
    rng = np.random.default_rng(5)
    src = util.synth_frame(97, 131, 9)
    xm = (rng.integers(-800, 5000, (80, 120)) / 32.0).astype(np.float32)
    ym = (rng.integers(-800, 3800, (80, 120)) / 32.0).astype(np.float32)
    xm[0, :6] = -1; ym[0, :6] = -1
    s16 = rng.integers(-700, 700, (66, 96, 3)).astype(np.int16)
    got = dict(
        remap_linear_reflect=sha(orc.remap_bilinear_u8(src, xm, ym, "reflect")),
        remap_linear_const=sha(orc.remap_bilinear_u8(src, xm, ym, "const")),
        remap_nearest=sha(orc.remap_nearest_u8(src[:, :, 0].copy(), xm, ym)),
        remap_cubic=sha(orc.remap_cubic_u8(src, xm, ym)),
        resize_down=sha(orc.resize_bilinear_u8(src, (64, 40))),
        resize_up=sha(orc.resize_bilinear_u8(src, (300, 211))),
        pyrdown=sha(orc.pyrdown_s16(s16)),
        pyrup=sha(orc.pyrup_s16(s16, (192, 132))),
        pyrdown_odd=sha(orc.pyrdown_s16(np.ascontiguousarray(s16[:65, :95]))),
    )
    assert got == want


@pytest.mark.parametrize("name,warp", [("cfg1_small", "spherical"), ("cfg1_cyl_small", "cylindrical")])
def test_geometry_matches_cv2(name, warp):
    g = load(name)
    t, _ = golden_tables(g, "2222", warp)
    assert np.array_equal(np.array(t.corners), g["corners"])
    assert np.array_equal(np.array(t.sizes), g["sizes"])
    assert tuple(g["dst_roi"]) == t.dst_roi


@pytest.mark.parametrize("name,warp", [("cfg1_small", "spherical"), ("cfg1_cyl_small", "cylindrical")])
def test_multiband_panorama_bit_exact(name, warp):
    g = load(name)
    t, imgs = golden_tables(g, "2222", warp)
    nb = int(g["num_bands"])
    out = compose.process(t, imgs, "multiband", nb, cut=tuple(g["cut"]), ext_weights=ext_weights(g, t, nb))
    assert np.array_equal(out, g["pano_multiband"]), util.report(name, out, g["pano_multiband"])


def test_multiband_own_weights_within_1lsb():
    g = load("cfg1_small")
    t, imgs = golden_tables(g, "2222", "spherical")
    out = compose.process(t, imgs, "multiband", int(g["num_bands"]), cut=tuple(g["cut"]))
    d = np.abs(out.astype(int) - g["pano_multiband"].astype(int))
    assert d.max() <= 1, util.report("own weights", out, g["pano_multiband"])


def test_feather_no_and_gain_bit_exact():
    g = load("cfg1_small")
    t, imgs = golden_tables(g, "2222", "spherical")
    n = len(imgs)
    cut = tuple(g["cut"])
    fw = [g["fw%d" % i] for i in range(n)]
    out = compose.process(t, imgs, "feather", feather_weights=fw, cut=cut)
    assert np.array_equal(out, g["pano_feather"]), util.report("feather", out, g["pano_feather"])
    out = compose.process(t, imgs, "no", cut=cut)
    assert np.array_equal(out, g["pano_no"]), util.report("no", out, g["pano_no"])
    t.gain_maps = [g["gain%d" % i] for i in range(n)]
    out = compose.process(t, imgs, "feather", feather_weights=fw, cut=cut)
    assert np.array_equal(out, g["pano_gain_feather"]), util.report("gain+feather", out, g["pano_gain_feather"])
    nb = int(g["num_bands"])
    out = compose.process(t, imgs, "multiband", nb, cut=cut, ext_weights=ext_weights(g, t, nb))
    assert np.array_equal(out, g["pano_gain_multiband"]), util.report("gain+mb", out, g["pano_gain_multiband"])


def test_front_end_bit_exact():
    g = np.load(os.path.join(util.GOLDEN, "frontend_small.npz"))
    mx, my = orc.init_undistort_map(g["K"], g["D"], g["newK"], 480, 270)
    ixy, frac = orc.convert_maps(mx, my)
    assert sha(ixy) == str(g["map_ixy_sha"]) and sha(frac) == str(g["map_frac_sha"])
    rect = [int(v) for v in g["rect"]]
    a = util.synth_frame(270, 480, 77, channels=4)
    assert np.array_equal(compose.front_end(a, (480, 270), mx, my, rect, (480, 270)), g["out_same"])
    b = util.synth_frame(540, 960, 78, channels=4)
    assert np.array_equal(compose.front_end(b, (480, 270), mx, my, rect, (360, 203)), g["out_down"])
    assert np.array_equal(compose.front_end(b, (480, 270), mx, my, rect, (360, 203), undistort=False), g["out_noud"])


def test_widened_rows_against_cv2_fixture():
    """SURVEY 8(f) restatements (YUYV ingest, updateMask tail, two-ring epilogue) vs cv2 outputs stored in
    tests/golden/widened_small.npz -- the pin that holds without cv2."""
    g = load("widened_small")
    assert np.array_equal(compose.yuyv_to_bgra(g["yuyv"]), g["yuyv_bgra"])
    assert np.array_equal(compose.seam_mask_tail(g["seam_low"], g["seam_full"]), g["seam_mask"])
    assert np.array_equal(compose.ring_epilogue(g["ring_up"], g["ring_down"], "resize"), g["ring_resize"])
    assert np.array_equal(compose.ring_epilogue(g["ring_up"], g["ring_down"], "crop", finalcut=3), g["ring_crop"])
