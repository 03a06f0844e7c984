"""CPU-only: the CCC white-balance arithmetic the kernels use (host build of ccc_math.cuh) against
the cv2 replay of convolutional_color_constancy.cpp (oracle.cv2_oracle.CCC)."""
import ctypes
import os

import cv2
import numpy as np
import pytest

from conftest import ROOT
from oracle import cv2_oracle as O

P = ctypes.c_void_p
MODEL = os.path.join(ROOT, "raw_image_pipeline_b200", "config", "ccc_model.bin")
DATA = os.path.join(ROOT, "tests", "golden")


def _small(hostsim, img):
    out = np.empty((270, 360, 3), np.uint8)
    hostsim.hs_ccc_small(P(img.ctypes.data), ctypes.c_int(img.shape[0]), ctypes.c_int(img.shape[1]), P(out.ctypes.data))
    return out


def _images():
    rng = np.random.default_rng(3)
    yield "noise 540x720", rng.integers(0, 256, (540, 720, 3), dtype=np.uint8)
    yield "noise 1080x1920", rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    yield "noise 271x361", rng.integers(0, 256, (271, 361, 3), dtype=np.uint8)
    yield "noise 135x180 (upscale)", rng.integers(0, 256, (135, 180, 3), dtype=np.uint8)
    yield "noise 333x1001", rng.integers(0, 256, (333, 1001, 3), dtype=np.uint8)
    base = cv2.resize(rng.integers(0, 256, (12, 16, 3), dtype=np.uint8), (1440, 1080), interpolation=cv2.INTER_CUBIC)
    yield "smooth 1080x1440", base


@pytest.mark.parametrize("name,img", list(_images()), ids=lambda v: v if isinstance(v, str) else "")
def test_resize_to_small_image_is_bit_exact(hostsim, name, img):
    ref = cv2.resize(img, (360, 270))
    got = _small(hostsim, np.ascontiguousarray(img))
    assert int((got != ref).sum()) == 0, name


def weight_table(n):
    """hist(u,v) after k sequential `+= 1/97200` fp32 additions (ccc.cpp:237-263)."""
    w = np.float32(1.0) / np.float32(97200)
    t = np.zeros(n + 1, np.float32)
    for k in range(1, n + 1):
        t[k] = np.float32(t[k - 1] + w)
    return t


@pytest.mark.parametrize("thr", [(0.9, 0.1), (0.8, 0.2)])
def test_histogram_counts_reproduce_the_oracle_histogram(hostsim, thr):
    rng = np.random.default_rng(11)
    ccc = O.CCC(MODEL)
    ccc.bright_thr, ccc.dark_thr = np.float32(thr[0]), np.float32(thr[1])
    wt = weight_table(97200)
    for small in (rng.integers(0, 256, (270, 360, 3), dtype=np.uint8),
                  np.clip(rng.normal(120, 30, (270, 360, 3)), 0, 255).astype(np.uint8),
                  np.full((270, 360, 3), 204, np.uint8), np.full((270, 360, 3), 51, np.uint8)):
        ref = ccc.histogram(small.astype(np.float32))
        counts = np.zeros(65536, np.uint32)
        hostsim.hs_ccc_counts.restype = ctypes.c_long
        used = hostsim.hs_ccc_counts(P(small.ctypes.data), ctypes.c_long(270 * 360), ctypes.c_float(float(np.float32(255) * ccc.bright_thr)),
                                     ctypes.c_float(float(np.float32(255) * ccc.dark_thr)), ctypes.c_float(-1.421875),
                                     ctypes.c_float(1.0 / 64.0), P(counts.ctypes.data))
        assert used == ccc.n_samples
        got = wt[counts].reshape(256, 256)
        assert np.array_equal(got, ref)


def test_gray_formula_exhaustive(hostsim):
    """The mask depends on gray <= / > thresholds: the gray value itself must match cv2 bit for bit."""
    from conftest import cube
    img = cube().astype(np.float32)
    ref = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    b, g, r = img[..., 0].astype(np.float64), img[..., 1].astype(np.float64), img[..., 2].astype(np.float64)
    f = np.float32
    t = (g.astype(np.float32) * f(0.587)).astype(np.float64)
    t = (b * np.float64(f(0.114)) + t).astype(np.float32).astype(np.float64)   # exact product + one rounding == fmaf
    got = (r * np.float64(f(0.299)) + t).astype(np.float32)
    # (double evaluation of a*b+c rounds once to double first; a mismatch with fmaf needs a 2^-29 coincidence)
    assert int((got != ref).sum()) <= 2


def test_gains_table_matches_oracle(hostsim):
    ccc = O.CCC(MODEL)
    libm = ctypes.CDLL("libm.so.6"); libm.expf.restype = ctypes.c_float; libm.expf.argtypes = [ctypes.c_float]
    f = np.float32
    tab = np.array([f(1.0) / f(libm.expf(float(-(f(f(k) * f(1 / 64)) + f(-1.421875))))) for k in range(256)], np.float32)
    for uv in ((111, 139), (92, 92), (0, 255), (255, 0), (128, 128), (17, 200)):
        ccc.uv_pos = uv
        ref = ccc.gains()
        got = np.zeros(3, np.float32)
        hostsim.hs_ccc_gains(ctypes.c_int(uv[0]), ctypes.c_int(uv[1]), P(tab.ctypes.data), P(got.ctypes.data))
        assert [f(x) for x in ref] == list(got)
