"""The C++ drop-in wrapper (include/ocvstitcher_b200.hpp) driven from a compiled C++ caller, the way the
reference's executables use the header-only ocvStitcher (src/replay.cpp:206-288).  CPU: the caller compiles
against the headers, links the C ABI library and fails loudly without a device.  GPU: its panoramas are
bit-identical to the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

import util
from golden import calib
from oracle import compose

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "img-stitching_b200", "lib")


def build_demo(tmp_path, name="wrapper_demo"):
    import panob200
    panob200.capi.lib()                       # builds lib/libpanob200.so if stale
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cxx", name + ".cpp"), "-L", LIBDIR, "-lpanob200",
           "-Wl,-rpath," + LIBDIR, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_inputs(path, t, Ks, Rs, scale, W, H, nb, cut, sets):
    with open(path, "wb") as f:
        f.write(struct.pack("<11i", 4, W, H, 0, 2, nb, *cut, len(sets)))
        f.write(struct.pack("<f", float(scale)))
        f.write(np.asarray(Ks, np.float32).tobytes())
        f.write(np.asarray(Rs, np.float32).tobytes())
        for m in t.blend_masks:
            f.write(struct.pack("<2i", m.shape[1], m.shape[0]))
            f.write(np.ascontiguousarray(m, np.uint8).tobytes())
        for s in sets:
            for im in s:
                f.write(np.ascontiguousarray(im, np.uint8).tobytes())


def setup(tmp_path, W=240, H=135, nb=3, nsets=2):
    Ks, Rs, scale = calib.rig("2222", W)
    t = compose.build_tables(Ks, Rs, scale, (W, H), "spherical")
    t.blend_masks = util.soft_masks(t)
    sets = [util.synth_set(4, H, W, 77 + s) for s in range(nsets)]
    cut = (8, 4, t.dst_roi[2] - 16, t.dst_roi[3] - 8)
    inp = str(tmp_path / "in.bin")
    write_inputs(inp, t, Ks, Rs, scale, W, H, nb, cut, sets)
    return t, sets, cut, inp


def test_cxx_caller_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    import torch
    exe = build_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("CUDA device present: covered by the gpu test")
    _, _, _, inp = setup(tmp_path)
    r = subprocess.run([exe, inp, str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CUDA device" in r.stderr and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cxx_caller_panoramas_bit_exact(tmp_path):
    exe = build_demo(tmp_path)
    t, sets, cut, inp = setup(tmp_path)
    outp = str(tmp_path / "out.bin")
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(outp, "rb").read()
    ow, oh, ns = struct.unpack("<3i", raw[:12])
    assert (ow, oh, ns) == (cut[2], cut[3], len(sets))
    got = np.frombuffer(raw, np.uint8, offset=12).reshape(ns, oh, ow, 3)
    for s, frames in enumerate(sets):
        want = compose.process(t, frames, "multiband", 3, cut=cut)
        assert np.array_equal(got[s], want), util.report("c++ caller set %d" % s, got[s], want)


# ------------------------------------------------------------------ whole widened path from C++ (two-ring rig)

def write_params_file(path, Ks, Rs, scale, blocks=2):
    """cameraparaout_<id>.txt as saveCameraParams writes it (include/ocvstitcher.hpp:522-562); the LAST block counts."""
    with open(path, "w") as f:
        for b in range(blocks):
            f.write("2021-11-%02d-10-25-21:\n" % (10 + b))
            for K, R in zip(Ks, Rs):
                k = np.asarray(K, np.float32).reshape(-1) * (1.0 if b == blocks - 1 else 0.5)     # stale earlier block
                f.write(",".join(repr(float(v)) for v in list(k) + list(np.asarray(R, np.float32).reshape(-1))) + ",\n")
            f.write(repr(float(np.float32(scale))) + "\n")


def rig_setup(tmp_path, W=480, H=270, nb=3):
    import panob200
    cam = calib.CAM_LIJING_390_FOV60_1920
    s = W / 1920.0
    K = np.array(cam["K"], np.float64).reshape(3, 3).copy()
    K[0, 0] *= s; K[0, 2] *= s; K[1, 1] *= s; K[1, 2] *= s
    newK = np.array([[1627.5076 * s, 0, 943.1681 * s], [0, 1622.9720 * s, 571.5369 * s], [0, 0, 1]])
    rect = [int(round(69 * s)), int(round(103 * s)), int(1782 * s), int(889 * s)]
    Ks, Rs, scale = calib.rig("2222", W)
    scales = [scale, float(np.float32(scale * 0.9))]
    cfgdir = str(tmp_path) + os.sep
    for r in range(2):
        write_params_file(cfgdir + "cameraparaout_%d.txt" % (1 + r), Ks, Rs, scales[r])
    frames = [[util.synth_frame(H, W, 300 + 10 * r + i, channels=4) for i in range(4)] for r in range(2)]
    inp = str(tmp_path / "rig.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<8i", 4, W, H, nb, *rect))
        f.write(np.concatenate([K.reshape(-1), np.asarray(cam["distorParams"], np.float64)[:4], newK.reshape(-1)]).astype(np.float64).tobytes())
        for r in range(2):
            for im in frames[r]:
                f.write(np.ascontiguousarray(im, np.uint8).tobytes())
    return dict(W=W, H=H, nb=nb, K=K, D=cam["distorParams"], newK=newK, rect=rect, Ks=Ks, Rs=Rs, scales=scales,
                frames=frames, cfgdir=cfgdir, inp=inp)


def test_cxx_pipeline_compiles_and_parses_calibration_files(tmp_path):
    """CPU: the two-ring C++ caller builds; without a device it stops at the first CUDA object with the library's
    message.  (The calibration-file round trip needs the run to get past the front end, so it is checked on the GPU.)"""
    import torch
    exe = build_demo(tmp_path, "pipeline_demo")
    if torch.cuda.is_available():
        pytest.skip("CUDA device present: covered by the gpu test")
    g = rig_setup(tmp_path)
    r = subprocess.run([exe, g["cfgdir"], g["inp"], str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_cxx_two_ring_pipeline_bit_exact(tmp_path):
    """calibration files -> front end chained into two stitchers -> process x2 -> resize + vconcat + bar, all from
    C++ through include/ocvstitcher_b200.hpp, bit-identical to the oracle chain."""
    import panob200
    exe = build_demo(tmp_path, "pipeline_demo")
    g = rig_setup(tmp_path)
    outp = str(tmp_path / "out.bin")
    r = subprocess.run([exe, g["cfgdir"], g["inp"], outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H = g["W"], g["H"]
    from oracle import oracle as orc
    mx, my = orc.init_undistort_map(g["K"], g["D"], g["newK"], W, H)
    panos = []
    for ring in range(2):
        bgr = [compose.front_end(f, (W, H), mx, my, g["rect"], (W, H)) for f in g["frames"][ring]]
        t = compose.build_tables(g["Ks"], g["Rs"], g["scales"][ring], (W, H), "spherical")
        panos.append(compose.process(t, bgr, "multiband", g["nb"]))
    want = compose.ring_epilogue(panos[0], panos[1], "resize")
    raw = open(outp, "rb").read()
    w, h = struct.unpack("<2i", raw[:8])
    assert (w, h) == (want.shape[1], want.shape[0])
    got = np.frombuffer(raw, np.uint8, offset=8).reshape(h, w, 3)
    assert np.array_equal(got, want), util.report("c++ two-ring pipeline", got, want)
    # saveCameraParams round trip: what C++ re-saved parses (Python mirror of initCamParams) to what it loaded
    from img_stitching_b200.stitcher import parse_camera_params_file
    Ks2, Rs2, sc2 = parse_camera_params_file(g["cfgdir"] + "cameraparaout_9.txt", 4)
    assert np.allclose(np.asarray(Ks2), np.asarray(g["Ks"], np.float32), rtol=1e-5)
    assert np.allclose(np.asarray(Rs2), np.asarray(g["Rs"], np.float32), rtol=1e-5, atol=1e-6)
    assert abs(sc2 - g["scales"][0]) < 1e-2


# ------------------------------------------------------------------ the SDK facade from C++ (include/panocam.h:8-28)

def facade_setup(tmp_path, W=480, H=270, nb=3, finalcut=6, canvas=(640, 360)):
    g = rig_setup(tmp_path, W, H, nb)
    inp = str(tmp_path / "facade.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<11i", 4, W, H, nb, *g["rect"], finalcut, *canvas))
        f.write(np.concatenate([g["K"].reshape(-1), np.asarray(g["D"], np.float64)[:4], g["newK"].reshape(-1)]).astype(np.float64).tobytes())
        f.write(np.asarray(g["scales"], np.float32).tobytes())
        f.write(np.asarray(g["Ks"], np.float32).tobytes())
        f.write(np.asarray(g["Rs"], np.float32).tobytes())
        for r in range(2):
            for im in g["frames"][r]:
                f.write(np.ascontiguousarray(im, np.uint8).tobytes())
    g.update(inp=inp, finalcut=finalcut, canvas=canvas)
    return g


def test_cxx_facade_compiles_and_fails_loudly_without_a_device(tmp_path):
    import torch
    exe = build_demo(tmp_path, "panocam_demo")
    if torch.cuda.is_available():
        pytest.skip("CUDA device present: covered by the gpu test")
    g = facade_setup(tmp_path)
    r = subprocess.run([exe, g["inp"], str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_cxx_facade_pano_frame_bit_exact(tmp_path):
    """pano::panocam from C++: init -> setCamFrame x 8 -> calibration(seam finder callback) -> getPanoFrame -> fit2final.
    The seam-finder inputs, the device tail of the masks, both rings' compose, the finalcut crop + bar and the canvas fit
    are restated with the oracle; the stacked frame and the canvas must match byte for byte."""
    import panob200
    from oracle import oracle as orc
    exe = build_demo(tmp_path, "panocam_demo")
    g = facade_setup(tmp_path)
    outp = str(tmp_path / "out.bin")
    r = subprocess.run([exe, g["inp"], outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H = g["W"], g["H"]
    mx, my = orc.init_undistort_map(g["K"], g["D"], g["newK"], W, H)
    panos = []
    for ring in range(2):
        bgr = [compose.front_end(f, (W, H), mx, my, g["rect"], (W, H)) for f in g["frames"][ring]]
        t = compose.build_tables(g["Ks"], g["Rs"], g["scales"][ring], (W, H), "spherical")
        masks = []
        for i in range(4):
            # the application's stand-in seam finder on the library's low-resolution inputs (pinned vs cv2 in test_host_tables)
            _, _, low = panob200.capi.host_seam_input(panob200.capi.WARP_SPHERICAL, g["scales"][ring], g["Ks"][i], g["Rs"][i], bgr[i])
            w = low.shape[1]
            low = low.copy()
            low[:, :w // 5] = 0
            low[:, w - w // 5:] = 0
            masks.append(compose.seam_mask_tail(low, t.warped_masks[i]))
        t.blend_masks = masks
        panos.append(compose.process(t, bgr, "multiband", g["nb"]))
    want = compose.ring_epilogue(panos[0], panos[1], "crop", finalcut=g["finalcut"])
    raw = open(outp, "rb").read()
    w, h = struct.unpack("<2i", raw[:8])
    assert (w, h) == (want.shape[1], want.shape[0])
    got = np.frombuffer(raw, np.uint8, count=w * h * 3, offset=8).reshape(h, w, 3)
    assert np.array_equal(got, want), util.report("c++ facade getPanoFrame", got, want)
    cw, ch = g["canvas"]
    canvas = np.frombuffer(raw, np.uint8, offset=8 + w * h * 3).reshape(ch, cw, 3)
    assert np.array_equal(canvas, compose.fit2final(want, (cw, ch))), "fit2final canvas differs"
