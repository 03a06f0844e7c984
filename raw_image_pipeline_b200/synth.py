"""Synthetic Bayer frames: counter-based (Philox) and reproducible.

Two distributions:
  * ``"U"``  i.i.d. uniform u8 -- worst case for parity (max gradients, saturation everywhere)
  * ``"N"``  natural-ish: low-frequency RGB field x sensor colour cast + N(0,4) noise,
             mosaiced to the requested CFA (gives the CCC histogram a real peak).
"""
from __future__ import annotations

import numpy as np

# colour at (row%2, col%2) for each encoding: 0=B, 1=G, 2=R  
CFA = {
    "bayer_bggr8": ((0, 1), (1, 2)),
    "bayer_rggb8": ((2, 1), (1, 0)),
    "bayer_gbrg8": ((1, 0), (2, 1)),
    "bayer_grbg8": ((1, 2), (0, 1)),
}


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=int(seed)))


def bayer_frame(rows: int, cols: int, encoding: str = "bayer_rggb8", seed: int = 0,
                dist: str = "N") -> np.ndarray:
    """One ``rows x cols`` u8 Bayer frame."""
    rng = _rng(seed)
    if dist == "U":
        return rng.integers(0, 256, size=(rows, cols), dtype=np.uint8)
    if dist != "N":
        raise ValueError(dist)
    h, w = max(rows // 64, 2), max(cols // 64, 2)
    field = rng.uniform(0.05, 0.95, size=(h, w, 3)).astype(np.float32)
    cast = np.array([0.55, 1.0, 0.7], np.float32)  # (B, G, R)
    pat = CFA[encoding]
    out = np.empty((rows, cols), np.uint8)
    # upsample per CFA phase to avoid materialising a full (rows, cols, 3) float image
    for py in range(2):
        for px in range(2):
            c = pat[py][px]
            plane = _upsample_phase(field[:, :, c], rows, cols, py, px)
            noise = rng.normal(0.0, 4.0, size=plane.shape).astype(np.float32)
            v = plane * (cast[c] * 255.0) + noise
            out[py::2, px::2] = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return out


def _upsample_phase(field2d: np.ndarray, rows: int, cols: int, py: int, px: int) -> np.ndarray:
    """Bicubic upsample evaluated only at the (py, px) CFA phase sites."""
    def axis_weights(n_out, n_in, phase):
        pos = np.arange(phase, n_out, 2)
        x = (pos + 0.5) * (n_in / n_out) - 0.5
        i0 = np.floor(x).astype(np.int64)
        t = (x - i0).astype(np.float32)
        a = -0.5
        w = np.stack([
            ((a * (t + 1) - 5 * a) * (t + 1) + 8 * a) * (t + 1) - 4 * a,
            ((a + 2) * t - (a + 3)) * t * t + 1,
            ((a + 2) * (1 - t) - (a + 3)) * (1 - t) * (1 - t) + 1,
            ((a * (2 - t) - 5 * a) * (2 - t) + 8 * a) * (2 - t) - 4 * a,
        ], axis=1).astype(np.float32)
        idx = np.clip(i0[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)
        return idx, w
    iy, wy = axis_weights(rows, field2d.shape[0], py)
    ix, wx = axis_weights(cols, field2d.shape[1], px)
    tmp = np.einsum("rkx,rk->rx", field2d[iy], wy)       # (rows/2, w)
    out = np.einsum("rxk,xk->rx", tmp[:, ix], wx)        # (rows/2, cols/2)
    return out


def bayer_batch(n: int, rows: int, cols: int, encoding: str = "bayer_rggb8", seed0: int = 0,
                dist: str = "N", out: np.ndarray | None = None) -> np.ndarray:
    """``n`` frames, seeds ``seed0 + i`` (seed = 1000*config + frame index)."""
    if out is None:
        out = np.empty((n, rows, cols), np.uint8)
    for i in range(n):
        out[i] = bayer_frame(rows, cols, encoding, seed0 + i, dist)
    return out
