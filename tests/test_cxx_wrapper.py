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


def build_demo(tmp_path):
    import panob200
    panob200.capi.lib()                       # builds lib/libpanob200.so if stale
    exe = str(tmp_path / "wrapper_demo")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cxx", "wrapper_demo.cpp"), "-L", LIBDIR, "-lpanob200",
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
