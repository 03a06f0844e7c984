"""CPU oracle for the RAW hot path -- TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  ``raw_image_pipeline_b200`` must never import anything from
``oracle/``.

What it is
----------
A call-for-call replay of the reference's **CPU** path
(``raw_image_pipeline/include/raw_image_pipeline/raw_image_pipeline.hpp:143-172``)
through the same OpenCV functions the reference calls, using the Python ``cv2``
wheel (4.13.0 in this image).  The reference C++ cannot be compiled here (no OpenCV
C++ headers, Eigen, yaml-cpp, boost, catkin), and the path's arithmetic lives in
OpenCV (un-vendored, ``find_package(OpenCV REQUIRED)``
``raw_image_pipeline/CMakeLists.txt:29``; README pins 4.2) and Eigen3 (2x2
``inverse()``, ``white_balance.cpp:114-115``).  So:

* every OpenCV call is made for real through cv2 (same function, same flags);
* Eigen's fixed-size 2x2 inverse is restated in fp32 (``_eigen_inverse2f``);
* the vignetting-mask double loop (``vignetting_correction.cpp:32-63``) is restated in
  C (``oracle/vignetting_mask.c``) and evaluated by this box's libm;
* ``std::exp(float)`` in ``computeGains`` is evaluated by glibc ``expf`` via ctypes.

Pinning status: the reference ships no golden vectors / KATs for this path
(SURVEY.md section 4), so parity is pinned by OpenCV's actual behaviour (cv2 4.13.0)
plus the restated glue above; the Eigen 2x2 inverse is the one piece with no runnable
reference behind it ("parity unpinned" for that 8-flop solve; sensitivity measured as
nil, see tests/test_oracle.py::test_pca_lut_insensitive_to_ulp).
"""
from __future__ import annotations

import ctypes
import math
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import cv2
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBM = ctypes.CDLL("libm.so.6")
_LIBM.expf.restype = ctypes.c_float
_LIBM.expf.argtypes = [ctypes.c_float]

BAYER_CODES = {
    # debayer.cpp:48-70 : encoding -> cv::demosaicing code (then COLOR_RGB2BGR swap)
    "bayer_bggr8": cv2.COLOR_BayerBG2BGR,
    "bayer_gbrg8": cv2.COLOR_BayerGB2BGR,
    "bayer_grbg8": cv2.COLOR_BayerGR2BGR,
    "bayer_rggb8": cv2.COLOR_BayerRG2BGR,
}
# debayer.hpp:74-81 (note the missing comma in the reference list)
BAYER_TYPES = [
    "bayer_bggr8", "bayer_gbrg8", "bayer_grbg8", "bayer_rggb8" "bayer_bggr16",
    "bayer_gbrg16", "bayer_grbg16", "bayer_rggb16",
]


def to_u8(x: np.ndarray) -> np.ndarray:
    """cv::Mat::convertTo(CV_8U) for a float Mat: saturate_cast<uchar>(cvRound(v))."""
    return cv2.add(x, 0.0, dtype=cv2.CV_8U)


# --------------------------------------------------------------------------------------
# vignetting mask (C restatement of vignetting_correction.cpp:32-63)
# --------------------------------------------------------------------------------------
_mask_lib = None


def _load_mask_lib():
    global _mask_lib
    if _mask_lib is None:
        path = os.path.join(_HERE, "_build", "libvignetting_mask.so")
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} missing: run `make -C oracle` (or __graft_entry__.build()) first")
        lib = ctypes.CDLL(path)
        lib.oracle_vignetting_mask.restype = ctypes.c_int
        lib.oracle_vignetting_mask.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_void_p]
        _mask_lib = lib
    return _mask_lib


def vignetting_mask(rows: int, cols: int, scale: float, a2: float, a4: float) -> np.ndarray:
    """Mask the reference builds for a ``rows x cols`` image (vignetting_correction.cpp:69
    calls precomputeVignettingMask(image.cols, image.rows))."""
    lib = _load_mask_lib()
    out = np.empty((rows, cols), np.float32)
    rc = lib.oracle_vignetting_mask(rows, cols, scale, a2, a4, out.ctypes.data)
    assert rc == 0
    return out


# --------------------------------------------------------------------------------------
# parameters (defaults = raw_image_pipeline.cpp:58-153 YAML defaults)
# --------------------------------------------------------------------------------------
@dataclass
class OracleParams:
    debayer_enabled: bool = True
    debayer_encoding: str = "auto"
    # EXTENSION beyond the reference (which throws for 16-bit Bayer, debayer.cpp:76-78): accept bayer_*16, see debayer16()
    debayer_allow_16bit: bool = False
    flip_enabled: bool = False
    flip_angle: int = 0
    wb_enabled: bool = False
    wb_method: str = "ccc"
    wb_clipping_percentile: float = 20.0
    wb_bright_thr: float = 0.8
    wb_dark_thr: float = 0.1
    wb_temporal_consistency: bool = True
    cc_enabled: bool = False
    cc_available: bool = True
    cc_matrix: List[float] = field(default_factory=lambda: [1, 0, 0, 0, 1, 0, 0, 0, 1])
    cc_bias: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    gamma_enabled: bool = False
    gamma_method: str = "custom"
    gamma_k: float = 0.8
    vig_enabled: bool = False
    vig_scale: float = 1.5
    vig_a2: float = 1e-3
    vig_a4: float = 1e-6
    enh_enabled: bool = False
    # *member* values after the cross-wired setters (color_enhancer.cpp:23-33)
    enh_hue_gain: float = 1.0
    enh_saturation_gain: float = 1.0
    enh_value_gain: float = 1.0
    und_enabled: bool = False
    und_available: bool = True
    und_model: str = "equidistant"
    und_balance: float = 0.0
    und_fov_scale: float = 1.0
    und_width: int = 720
    und_height: int = 540
    und_new_width: Optional[int] = None
    und_new_height: Optional[int] = None
    und_K: List[float] = field(default_factory=lambda: [
        347.548139773951, 0.0, 342.454373227748, 0.0, 347.434712422309, 271.368057185649,
        0.0, 0.0, 1.0])
    und_D: List[float] = field(default_factory=lambda: [
        -0.0396482888762527, -0.00367688950406141, 0.00391742438164282, -0.00178738156007817])
    und_R: List[float] = field(default_factory=lambda: [1, 0, 0, 0, 1, 0, 0, 0, 1])
    # reference bug B-7: the mask is regenerated on every frame for non-square images
    regen_mask_every_frame: bool = False


# --------------------------------------------------------------------------------------
# stages
# --------------------------------------------------------------------------------------
def debayer(image: np.ndarray, encoding: str) -> Tuple[np.ndarray, str]:
    """debayer.cpp:45-79 (CPU overload)."""
    if encoding in BAYER_CODES:
        out = cv2.demosaicing(image, BAYER_CODES[encoding])
        out = cv2.cvtColor(out, cv2.COLOR_RGB2BGR)
        return out, "bgr8"
    if encoding == "rgb8":
        return cv2.cvtColor(image, cv2.COLOR_RGB2BGR), encoding  # CPU branch keeps encoding (B-2/8b)
    if encoding in BAYER_TYPES:
        raise ValueError("Encoding [" + encoding + "] is a valid pattern but is not supported!")
    return image, encoding


def debayer16(image: np.ndarray, encoding: str) -> Tuple[np.ndarray, str]:
    """EXTENSION (the reference lists the bayer_*16 names but throws for them).  Defined as what the
    reference's own 8-bit code does, one depth up, followed by the 16 -> 8 bit reduction ROS' cv_bridge applies
    (convertTo(CV_8U, 255 / 65535)): cv::demosaicing on the CV_16UC1 frame with the 8-bit path's code and R/B swap, then
    saturate_cast<uchar>(v * (1 / 257.f)); the rest of the chain (cv::LUT, cvtColor ... all 8-bit only) runs on that."""
    code = BAYER_CODES[encoding[:-2] + "8"]
    out = cv2.demosaicing(image, code)
    out = cv2.cvtColor(out, cv2.COLOR_RGB2BGR)
    return cv2.convertScaleAbs(out, alpha=1.0 / 257.0), "bgr8"


def flip(image: np.ndarray, angle: int) -> np.ndarray:
    """flip.cpp:37-58."""
    if angle == 90:
        return cv2.flip(cv2.transpose(image), 1)
    if angle == 180:
        return cv2.flip(image, -1)
    if angle == 270:
        return cv2.flip(cv2.transpose(image), 0)
    return image


def _eigen_inverse2f(a, b, c, d):
    """Eigen fixed-size 2x2 inverse in fp32 (compute_inverse<...,2>): invdet = 1/det,
    det = a*d - c*b (plain -O3 x86-64: no FMA), result = [d,-b;-c,a]*invdet."""
    f = np.float32
    a, b, c, d = f(a), f(b), f(c), f(d)
    det = f(f(a * d) - f(c * b))
    invdet = f(f(1.0) / det)
    return f(d * invdet), f(f(-b) * invdet), f(f(-c) * invdet), f(a * invdet)


def pca_coefficients(image: np.ndarray):
    """white_balance.cpp:76-115 -> ((alpha_b, beta_b), (alpha_r, beta_r)) as fp32."""
    ch = cv2.split(image)
    bf = ch[0].astype(np.float32)
    rf = ch[2].astype(np.float32)
    b2 = cv2.multiply(bf, bf)
    r2 = cv2.multiply(rf, rf)
    sum_r2 = cv2.sumElems(r2)[0]
    sum_b2 = cv2.sumElems(b2)[0]
    sum_g = cv2.sumElems(ch[1])[0]
    sum_r = cv2.sumElems(rf)[0]
    sum_b = cv2.sumElems(bf)[0]
    _, max_r, _, _ = cv2.minMaxLoc(rf)
    _, max_g, _, _ = cv2.minMaxLoc(ch[1])
    _, max_b, _, _ = cv2.minMaxLoc(bf)
    _, max_r2, _, _ = cv2.minMaxLoc(r2)
    _, max_b2, _, _ = cv2.minMaxLoc(b2)
    f = np.float32
    out = []
    with np.errstate(all="ignore"):
        for (s2, s1, m2, m1) in ((sum_b2, sum_b, max_b2, max_b), (sum_r2, sum_r, max_r2, max_r)):
            i00, i01, i10, i11 = _eigen_inverse2f(s2, s1, m2, m1)
            v0, v1 = f(sum_g), f(max_g)
            alpha = f(f(i00 * v0) + f(i01 * v1))
            beta = f(f(i10 * v0) + f(i11 * v1))
            out.append((alpha, beta))
    return out, (bf, rf, b2, r2, ch[1])


def white_balance_pca(image: np.ndarray) -> np.ndarray:
    """white_balance.cpp:73-136."""
    (cb, cr), (bf, rf, b2, r2, g) = pca_coefficients(image)
    b_point = cv2.addWeighted(b2, float(cb[0]), bf, float(cb[1]), 0.0)
    r_point = cv2.addWeighted(r2, float(cr[0]), rf, float(cr[1]), 0.0)
    _, b_point = cv2.threshold(b_point, 255, 255, cv2.THRESH_TRUNC)
    _, r_point = cv2.threshold(r_point, 255, 255, cv2.THRESH_TRUNC)
    return cv2.merge([to_u8(b_point), g, to_u8(r_point)])


class CCC:
    """convolutional_color_constancy.cpp (CPU overloads)."""

    def __init__(self, model_path: str):
        self.small_size = (360, 270)
        self.bin_size = np.float32(1.0 / 64.0)
        self.uv0 = np.float32(-1.421875)
        self.bright_thr = np.float32(0.9)
        self.dark_thr = np.float32(0.1)
        self.temporal = False
        self.first_frame = True
        self.load_model(model_path)

    def load_model(self, path):  # :116-207
        d = open(path, "rb").read()
        if d[:8] == b"RIPCCC1\0":
            # this repo's layout (tools/convert_ccc_model.py): same numbers, already transposed
            w, h = struct.unpack("ii", d[8:16])
            a = np.frombuffer(d[16:16 + 8 * w * h], dtype=np.float32)
            self.filter = np.ascontiguousarray(a[:w * h].reshape(h, w))
            self.bias = np.ascontiguousarray(a[w * h:].reshape(h, w))
        else:
            # the reference's model/default.bin: transposed right after reading (:131-132)
            w, h = struct.unpack("ii", d[:8])
            a = np.frombuffer(d[8:8 + 8 * w * h], dtype=np.float32)
            self.filter = np.ascontiguousarray(a[:w * h].reshape(h, w).T)
            self.bias = np.ascontiguousarray(a[w * h:].reshape(h, w).T)
        self.w, self.h = w, h
        self.filter_fft = cv2.dft(self.filter, flags=0, nonzeroRows=h)
        self.bias_fft = cv2.dft(self.bias, flags=0, nonzeroRows=h)
        self.uv_pos = (h // 2, w // 2)  # cv::Point(x, y)
        self.kf = cv2.KalmanFilter(2, 2, 0, cv2.CV_32F)
        self.kf.statePre = np.array([[self.uv_pos[0]], [self.uv_pos[1]]], np.float32)
        self.kf.statePost = np.array([[self.uv_pos[0]], [self.uv_pos[1]]], np.float32)
        self.kf.transitionMatrix = np.eye(2, dtype=np.float32)
        self.kf.processNoiseCov = np.eye(2, dtype=np.float32)
        self.kf.measurementMatrix = np.eye(2, dtype=np.float32)
        self.kf.measurementNoiseCov = 10 * np.eye(2, dtype=np.float32)

    def histogram(self, small_f: np.ndarray) -> np.ndarray:  # :210-271
        gray = cv2.cvtColor(small_f, cv2.COLOR_BGR2GRAY)
        # `255 * bright_thr_` is int*float -> float in C++, then widened to the double argument
        _, upper = cv2.threshold(gray, float(np.float32(255) * self.bright_thr), 255, cv2.THRESH_BINARY_INV)
        _, lower = cv2.threshold(gray, float(np.float32(255) * self.dark_thr), 255, cv2.THRESH_BINARY)
        mask = cv2.bitwise_and(upper, lower)
        with np.errstate(all="ignore"):
            lg = cv2.log(small_f)
        lb, lgn, lr = lg[..., 0], lg[..., 1], lg[..., 2]
        ok = np.isfinite(lb) & np.isfinite(lgn) & np.isfinite(lr) & ~(mask < 1.0)
        with np.errstate(all="ignore"):
            uf = ((lgn - lr) - self.uv0) / self.bin_size  # fp32, left to right
            vf = ((lgn - lb) - self.uv0) / self.bin_size
        uf = uf[ok].astype(np.float64)
        vf = vf[ok].astype(np.float64)
        rnd = lambda x: np.trunc(x + np.copysign(0.5, x)).astype(np.int64)  # std::round
        u = np.clip(rnd(uf), 0, 255)
        v = np.clip(rnd(vf), 0, 255)
        hist = np.zeros((self.h, self.w), np.float32)
        n = small_f.shape[0] * small_f.shape[1]
        wgt = np.float32(1.0) / np.float32(n)
        np.add.at(hist, (u, v), wgt)  # sequential fp32 accumulation, raster order
        self.n_samples = int(ok.sum())
        return hist

    def response(self, hist):  # :273-298
        hist_fft = cv2.dft(hist, flags=0, nonzeroRows=self.h)
        resp_fft = cv2.mulSpectrums(self.filter_fft, hist_fft, 0)
        resp_fft = cv2.add(resp_fft, self.bias_fft)
        resp = cv2.dft(resp_fft, flags=cv2.DFT_INVERSE | cv2.DFT_REAL_OUTPUT, nonzeroRows=self.h)
        _, _, _, max_loc = cv2.minMaxLoc(resp)
        self.response_map = resp
        return max_loc  # (x, y)

    def kalman(self):  # :300-340
        if self.first_frame:
            self.first_frame = False
            self.kf.statePost = np.array([[self.uv_pos[0]], [self.uv_pos[1]]], np.float32)
        else:
            self.kf.predict()
            meas = np.array([[self.uv_pos[0]], [self.uv_pos[1]]], np.float32)
            est = self.kf.correct(meas)
            self.uv_pos = (int(est[0, 0]), int(est[1, 0]))  # float -> int truncation

    def gains(self):  # :342-381
        f = np.float32
        Lu = f(f(f(self.uv_pos[0]) * self.bin_size) + self.uv0)
        Lv = f(f(f(self.uv_pos[1]) * self.bin_size) + self.uv0)
        z = f(1.0)
        gr = f(z / f(_LIBM.expf(float(-Lu))))
        gg = z
        gb = f(z / f(_LIBM.expf(float(-Lv))))
        fac = min(min(gr, gg), gb)
        return f(gb / fac), f(gg / fac), f(gr / fac)

    def balance_white(self, src: np.ndarray) -> np.ndarray:  # :91-113
        small = cv2.resize(src, self.small_size)
        small_f = small.astype(np.float32)
        hist = self.histogram(small_f)
        self.hist = hist
        self.uv_pos = self.response(hist)
        if self.temporal:
            self.kalman()
        gb, gg, gr = self.gains()
        self.last_gains = (gb, gg, gr)
        return cv2.multiply(src, (float(gb), float(gg), float(gr), 0.0))


def color_calibration(image: np.ndarray, matrix, bias) -> np.ndarray:
    """color_calibration.cpp:91-104. ``matrix`` is row-major 3x3 (YAML order), stored as Matx33f."""
    M = np.asarray(matrix, np.float64).reshape(3, 3).astype(np.float32)
    rows, cols = image.shape[:2]
    flat_f = image.reshape(rows * cols, 3).astype(np.float32)
    mixed = cv2.gemm(flat_f, np.ascontiguousarray(M.T), 1.0, None, 0.0)
    image_f = mixed.reshape(rows, cols, 3)
    image_f = cv2.add(image_f, (float(bias[0]), float(bias[1]), float(bias[2]), 0.0))
    return to_u8(image_f)


def gamma_lut(k: float) -> np.ndarray:
    """gamma_correction.cpp:35-42."""
    lut = np.zeros(256, np.uint8)
    for i in range(256):
        f = np.float32(i / 255.0)
        f = np.float32(math.pow(float(f), k))
        v = float(f) * 255.0
        lut[i] = int(min(max(np.rint(v), 0), 255))
    return lut


def gamma(image: np.ndarray, k: float) -> np.ndarray:
    return cv2.LUT(image, gamma_lut(k))


def vignetting(image: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """vignetting_correction.cpp:68-93."""
    lab = cv2.cvtColor(image, cv2.COLOR_BGR2Lab)
    ch = list(cv2.split(lab))
    lf = ch[0].astype(np.float32)
    lf = cv2.multiply(lf, mask, scale=1.0, dtype=cv2.CV_32F)
    ch[0] = to_u8(lf)
    lab = cv2.merge(ch)
    return cv2.cvtColor(lab, cv2.COLOR_Lab2BGR)


def color_enhancer(image: np.ndarray, hue_gain, saturation_gain, value_gain) -> np.ndarray:
    """color_enhancer.cpp:38-47 (gains are the *member* values)."""
    hsv = cv2.cvtColor(image, cv2.COLOR_BGR2HSV)
    hsv = cv2.multiply(hsv, (float(hue_gain), float(saturation_gain), float(value_gain), 0.0))
    return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)


def undistortion_maps(K, D, R, size, new_size, balance, fov_scale):
    """undistortion.cpp:197-238 -> (new_K 3x3 float64, map_x, map_y fp32 at ``size``)."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    D = np.asarray(D, np.float64).reshape(4, 1)
    R = np.asarray(R, np.float64).reshape(3, 3)
    newK = cv2.fisheye.estimateNewCameraMatrixForUndistortRectify(
        K, D, tuple(size), R, balance=balance, new_size=tuple(new_size), fov_scale=fov_scale)
    mx, my = cv2.fisheye.initUndistortRectifyMap(K, D, R, newK, tuple(size), cv2.CV_32F)
    return newK, mx, my


def remap(image, mx, my):
    """undistortion.cpp:240-245."""
    return cv2.remap(image, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)


# --------------------------------------------------------------------------------------
# the pipeline (raw_image_pipeline.hpp:143-172)
# --------------------------------------------------------------------------------------
class OraclePipeline:
    def __init__(self, params: OracleParams, ccc_model_path: Optional[str] = None):
        self.p = params
        self._mask = None
        self._mask_key = None
        self._maps = None
        self._maps_key = None
        self.ccc = CCC(ccc_model_path) if ccc_model_path else None
        self.stages: Dict[str, np.ndarray] = {}

    def _get_mask(self, rows, cols):
        p = self.p
        key = (rows, cols, p.vig_scale, p.vig_a2, p.vig_a4)
        if self._mask_key != key or (p.regen_mask_every_frame and rows != cols):
            self._mask = vignetting_mask(rows, cols, p.vig_scale, p.vig_a2, p.vig_a4)
            self._mask_key = key
        return self._mask

    def maps(self):
        p = self.p
        nw = p.und_new_width if p.und_new_width is not None else p.und_width
        nh = p.und_new_height if p.und_new_height is not None else p.und_height
        key = (tuple(p.und_K), tuple(p.und_D), tuple(p.und_R), p.und_width, p.und_height, nw, nh,
               p.und_balance, p.und_fov_scale)
        if self._maps_key != key:
            self._maps = undistortion_maps(p.und_K, p.und_D, p.und_R, (p.und_width, p.und_height),
                                           (nw, nh), p.und_balance, p.und_fov_scale)
            self._maps_key = key
        return self._maps

    def apply(self, image: np.ndarray, encoding: str, keep_stages: bool = False):
        p = self.p
        st = self.stages = {}
        if p.debayer_allow_16bit and encoding in ("bayer_rggb16", "bayer_bggr16", "bayer_gbrg16", "bayer_grbg16"):
            image, encoding = debayer16(image, encoding)  # extension, see debayer16()
        else:
            image, encoding = debayer(image, encoding)  # always runs (B-1)
        if keep_stages: st["debayer"] = image
        if p.flip_enabled:
            image = flip(image, p.flip_angle)
        self.dist_debayered = image  # FlipModule snapshot (flip.hpp:36-45)
        if keep_stages: st["flip"] = image
        ch3 = image.ndim == 3 and image.shape[2] == 3
        if p.wb_enabled and ch3:
            if p.wb_method == "pca":
                image = white_balance_pca(image)
            elif p.wb_method == "ccc":
                self.ccc.bright_thr = np.float32(p.wb_bright_thr)
                self.ccc.dark_thr = np.float32(p.wb_dark_thr)
                self.ccc.temporal = p.wb_temporal_consistency
                image = self.ccc.balance_white(image)
            elif p.wb_method in ("simple", "gray_world", "grey_world", "learned"):
                raise NotImplementedError("xphoto white balance is out of scope (no cv2.xphoto)")
            else:
                raise ValueError("White Balance method [" + p.wb_method + "] not supported. "
                                 "Supported algorithms: 'simple', 'gray_world', 'learned', 'ccc', 'pca'")
        if keep_stages: st["white_balance"] = image
        if p.cc_enabled and ch3 and p.cc_available:
            image = color_calibration(image, p.cc_matrix, p.cc_bias)
        if keep_stages: st["color_calibration"] = image
        if p.gamma_enabled:
            image = gamma(image, p.gamma_k)  # default == custom on CPU (gamma_correction.cpp:58-60)
        if keep_stages: st["gamma"] = image
        if p.vig_enabled:
            image = vignetting(image, self._get_mask(image.shape[0], image.shape[1]))
        if keep_stages: st["vignetting"] = image
        if p.enh_enabled and ch3:
            image = color_enhancer(image, p.enh_hue_gain, p.enh_saturation_gain, p.enh_value_gain)
        if keep_stages: st["color_enhancer"] = image
        self.dist_color = image  # UndistortionModule snapshot (undistortion.hpp:68)
        if p.und_enabled and p.und_available and p.und_model != "none":
            _, mx, my = self.maps()
            image = remap(image, mx, my)
        if keep_stages: st["undistortion"] = image
        self.processed = image
        return image, encoding
