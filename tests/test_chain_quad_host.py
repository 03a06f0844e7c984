"""CPU-only: the strip kernel's four-pixel form of the chain (chain_quad.cuh, its own table layout) against
chain_pixel() (pixel_math.cuh) -- which tests/test_pixel_math_host.py pins exhaustively to cv2 -- over the whole 2^24
colour cube, for every stage set, with and without a bias, pca-like (G identity) and ccc-like (all three) white-balance
tables, and for the row-tail variant."""
import ctypes

import numpy as np
import pytest

from conftest import CC_EXAMPLE, cube
from oracle import cv2_oracle as O

P = ctypes.c_void_p
CC2 = [1.6016861, -1.0104847, 0.2213647, -0.012442476, 1.0841908, -0.07573348, 0.09123942, -0.8107389, 1.8687756]


def _args(img, mask, cc, bias, enh, wb, gamma):
    cc = np.asarray(cc if cc is not None else np.eye(3).ravel(), np.float64).astype(np.float32)
    bias = np.asarray(bias, np.float64).astype(np.float32)
    enh = np.asarray(enh, np.float64)
    wb = np.ascontiguousarray(wb if wb is not None else np.tile(np.arange(256, dtype=np.uint8), 3))
    gamma = np.ascontiguousarray(gamma if gamma is not None else np.arange(256, dtype=np.uint8))
    m = None if mask is None else np.ascontiguousarray(mask, np.float32)
    return cc, bias, enh, wb, gamma, m


def chain_ref(hostsim, stages, img, width, **kw):
    cc, bias, enh, wb, gamma, m = _args(img, **kw)
    out = np.empty_like(img)
    hostsim.hs_chain(ctypes.c_uint(stages), ctypes.c_long(img.size // 3), ctypes.c_int(width), P(img.ctypes.data),
                     P(m.ctypes.data) if m is not None else None, P(cc.ctypes.data), P(bias.ctypes.data), P(enh.ctypes.data),
                     P(wb.ctypes.data), P(gamma.ctypes.data), P(out.ctypes.data))
    return out


def chain_quad(hostsim, stages, img, tail, **kw):
    cc, bias, enh, wb, gamma, m = _args(img, **kw)
    out = np.empty_like(img)
    hostsim.hs_chain_quad(ctypes.c_uint(stages), ctypes.c_long(img.size // 3), ctypes.c_int(tail), P(img.ctypes.data),
                          P(m.ctypes.data) if m is not None else None, P(cc.ctypes.data), P(bias.ctypes.data), P(enh.ctypes.data),
                          P(wb.ctypes.data), P(gamma.ctypes.data), P(out.ctypes.data))
    return out


def _wb_tables(rng, g_identity):
    b = np.sort(rng.integers(0, 256, 256)).astype(np.uint8)
    r = np.minimum(255, (np.arange(256) * 1.37 + 0.5).astype(np.int64)).astype(np.uint8)
    g = np.arange(256, dtype=np.uint8) if g_identity else np.minimum(255, (np.arange(256) * 1.11).astype(np.int64)).astype(np.uint8)
    return np.concatenate([b, g, r])


@pytest.mark.parametrize("stages", list(range(32)))
def test_chain_quad_equals_chain_pixel_on_the_whole_cube(hostsim, stages):
    img = cube()
    rng = np.random.default_rng(100 + stages)
    kw = dict(mask=rng.uniform(1.0, 2.6, img.shape[:2]).astype(np.float32) if stages & 8 else None,
              cc=CC_EXAMPLE if stages % 4 < 2 else CC2, bias=(0, 0, 0),
              enh=(1.0, 1.2, 1.0) if stages % 3 else (1.1, 0.8, 1.3), wb=_wb_tables(rng, g_identity=(stages % 8) < 4),
              gamma=O.gamma_lut(0.8))
    ref = chain_ref(hostsim, stages, img, width=img.shape[1], **kw)   # width 4096: no row tail
    got = chain_quad(hostsim, stages, img, 0, **kw)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("stages", [16, 24, 31])
def test_chain_quad_row_tail_variant(hostsim, stages):
    img = cube()[:1024]
    rng = np.random.default_rng(7)
    kw = dict(mask=rng.uniform(1.0, 2.6, img.shape[:2]).astype(np.float32) if stages & 8 else None, cc=CC_EXAMPLE, bias=(0, 0, 0),
              enh=(1.0, 1.2, 1.0), wb=_wb_tables(rng, True), gamma=O.gamma_lut(0.8))
    # hs_chain treats columns >= width & ~31 as the tail: width 16 makes every pixel one
    flat = np.ascontiguousarray(img.reshape(-1, 16, 3))
    kwf = dict(kw)
    if kw["mask"] is not None:
        kwf["mask"] = kw["mask"].reshape(-1, 16)
    ref = chain_ref(hostsim, stages, flat, width=16, **kwf)
    got = chain_quad(hostsim, stages, flat, 1, **kwf)
    assert int((got != ref).sum()) == 0


def test_full_chain_quad_against_cv2(hostsim, oracle_built):
    """the strip form against the cv2 stage functions directly (not only against chain_pixel)"""
    import cv2
    rng = np.random.default_rng(11)
    rows, cols = 480, 640
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    mask = O.vignetting_mask(rows, cols, 1.5, 1e-3, 1e-6)
    wb = _wb_tables(rng, True)
    got = chain_quad(hostsim, 31, img, 0, mask=mask, cc=CC_EXAMPLE, bias=(0, 0, 0), enh=(1.0, 1.2, 1.0), wb=wb, gamma=O.gamma_lut(0.8))
    ref = cv2.merge([cv2.LUT(img[..., 0], wb[:256]), cv2.LUT(img[..., 1], wb[256:512]), cv2.LUT(img[..., 2], wb[512:])])
    ref = O.color_calibration(ref, CC_EXAMPLE, (0, 0, 0))
    ref = O.gamma(ref, 0.8)
    ref = O.vignetting(ref, mask)
    ref = O.color_enhancer(ref, 1.0, 1.2, 1.0)
    assert int((got != ref).sum()) == 0
