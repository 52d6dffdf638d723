"""The worker pool that feeds pano_process from pageable host buffers (csrc/host_pool.cpp), checked on the CPU: a small
C++ driver (tests/cxx/host_pool_check.cpp) runs strided 2-D copies of awkward sizes and alignments through 1 / 3 / 8
threads with plain and streaming stores and compares every byte, including the bytes around the destination window."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_pool_copies_every_byte(tmp_path):
    exe = str(tmp_path / "host_pool_check")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread",
           os.path.join(ROOT, "tests", "cxx", "host_pool_check.cpp"),
           os.path.join(ROOT, "img-stitching_b200", "csrc", "host_pool.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout
