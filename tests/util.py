"""Shared test helpers: deterministic synthetic frames (integer-only, no cv2), configurations."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def synth_frame(h, w, seed, channels=3, cell=16):
    """Smooth full-range frame + mild noise from pure integer arithmetic (bit-stable everywhere)."""
    rng = np.random.default_rng(seed)
    gh, gw = h // cell + 2, w // cell + 2
    grid = rng.integers(0, 256, (gh, gw, channels), dtype=np.int64)
    ys, xs = np.arange(h), np.arange(w)
    y0, fy, x0, fx = ys // cell, (ys % cell)[:, None, None], xs // cell, (xs % cell)[None, :, None]
    a, b = grid[y0][:, x0], grid[y0][:, x0 + 1]
    c, d = grid[y0 + 1][:, x0], grid[y0 + 1][:, x0 + 1]
    v = (a * (cell - fy) * (cell - fx) + b * (cell - fy) * fx + c * fy * (cell - fx) + d * fy * fx) // (cell * cell)
    v = v + rng.integers(-6, 7, (h, w, channels))
    return np.clip(v, 0, 255).astype(np.uint8)


def synth_set(n, h, w, seed):
    return [synth_frame(h, w, 1000 * seed + i) for i in range(n)]


def soft_masks(tables, seed=0):
    """Deterministic soft seam-like masks without cv2: each camera keeps a central vertical band
    of its warped footprint with a linear 0..255 ramp, ANDed (min) with the warped validity mask."""
    out = []
    for i, (m, (w, h)) in enumerate(zip(tables.warped_masks, tables.sizes)):
        x = np.arange(w)
        lo, hi = int(w * 0.18), int(w * 0.82)
        ramp = np.clip(np.minimum(x - lo, hi - x) * 12 + 128, 0, 255).astype(np.uint8)
        band = np.broadcast_to(ramp[None, :], (h, w))
        out.append(np.minimum(band, m).astype(np.uint8))
    return out


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)


def report(name, got, ref):
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    return "%s: max|d|=%d mismatches=%d/%d psnr=%s" % (name, int(diff.max()), int((diff != 0).sum()), diff.size, psnr(got, ref))
