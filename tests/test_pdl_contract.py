"""Programmatic dependent launch (csrc/pdl.h): a kernel launched through launch_chain() may become resident while its
predecessor is still running, so its FIRST statement must be pdl_enter() (griddepcontrol.wait) -- a kernel that skipped it
would read data its predecessor has not written yet.  Static check of the sources: every kernel handed to launch_chain()
starts with pdl_enter(), and nothing reaches global memory before it."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "img-stitching_b200", "csrc")


def test_every_chain_kernel_waits_first():
    text = {f: open(os.path.join(CSRC, f)).read() for f in ("kernels.cu", "frontend.cu")}
    launched = set()
    for src in text.values():
        launched |= set(re.findall(r"launch_chain\(\s*([A-Za-z0-9_]+)", src))
    assert len(launched) >= 10, launched
    for name in sorted(launched):
        bodies = []
        for src in text.values():
            for m in re.finditer(r"__global__[^{;]*\b%s\s*\([^{;]*\)\s*\{" % re.escape(name), src):
                bodies.append(src[m.end():m.end() + 400])
        assert bodies, "kernel %s not found" % name
        for b in bodies:
            first = b.strip().split(";")[0].strip()
            assert first == "pdl_enter()", "%s: first statement is %r, not pdl_enter()" % (name, first)
