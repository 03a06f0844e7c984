#!/usr/bin/env python3
"""Generates raw_image_pipeline_b200/csrc/cv_tables.inc: the integer lookup tables behind
OpenCV's 8-bit BGR<->Lab and BGR->HSV conversions (the arithmetic the reference reaches
through cv::cvtColor in vignetting_correction.cpp:73,92 and color_enhancer.cpp:40,46).

The tables are constants of OpenCV's algorithm (imgproc/src/color_lab.cpp initLabTabs,
color_hsv.cpp), not of the reference.  They are rebuilt here from their published
formulas and then *validated exhaustively* (all 2^24 input triples, both directions)
against the cv2 installed in this image before the file is written; the two entries where
the closed form differs from OpenCV's softfloat evaluation are listed in EXCEPTIONS.

Run:  python tools/gen_cv_tables.py          (needs cv2; takes ~1 min)
"""
import os
import sys

import cv2
import numpy as np

f32 = np.float32
BASE = 16384
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "raw_image_pipeline_b200",
                   "csrc", "cv_tables.inc")

# ---- sRGBGammaTab_b: G[i] = rint(2040 * gamma(i/255f)) ---------------------------------
def _gamma_fwd(x):
    x = x.astype(np.float32)
    lo = x * (f32(1) / f32(12.92))
    hi = np.power((x.astype(np.float64) + 0.055) / 1.055, 2.4).astype(np.float32)
    return np.where(x <= f32(0.04045), lo, hi).astype(np.float32)

G = np.rint(f32(2040) * _gamma_fwd((np.arange(256) / f32(255)).astype(np.float32))).astype(np.int64)

# ---- LabCbrtTab_b: C[t] = rint(32768 * f(t/2040)), t = 0..2040 --------------------------
EXCEPTIONS = {49: 9454}   # closed form gives 9455 (value within 2e-4 of a .5 tie)
_t = (np.arange(2041).astype(np.float32) / f32(2040))
C = np.rint(32768 * np.where(_t < f32(0.008856), _t.astype(np.float64) * 7.787 + 16 / 116,
                             np.cbrt(_t.astype(np.float64)))).astype(np.int64)
for k, v in EXCEPTIONS.items():
    C[k] = v
assert C[324] == 17745

# ---- LabToYF_b --------------------------------------------------------------------------
Y = np.zeros(256, np.int64)
IFY = np.zeros(256, np.int64)
for i in range(256):
    if i <= 20:
        Y[i] = int(np.rint(f32(i * BASE * 20 * 9) / f32(17 * 29 * 29 * 29)))
        IFY[i] = int(np.rint(f32(BASE) * (f32(16) / f32(116) + f32(i * 5) / f32(3 * 17 * 29))))
    else:
        fy = f32(f32(i * 100 * BASE) / f32(255 * 116) + f32(16 * BASE) / f32(116))
        IFY[i] = int(np.rint(fy))
        Y[i] = int(np.rint(f32(f32(fy * fy) * fy) / f32(BASE * BASE)))

# ---- sRGBInvGammaTab_b: IG[i] = rint(255 * invgamma(i/4096)) ----------------------------
def _inv_gamma(x):
    x = x.astype(np.float64)
    return np.where(x <= 0.0031308, x * 12.92, 1.055 * np.power(x, 1 / 2.4) - 0.055)

IG = np.clip(np.rint(255 * _inv_gamma(np.arange(4096) / 4096.0)), 0, 255).astype(np.int64)

# ---- HSV division tables ----------------------------------------------------------------
SDIV = np.zeros(256, np.int64)
HDIV = np.zeros(256, np.int64)
for i in range(1, 256):
    SDIV[i] = int(np.rint((255 << 12) / (1.0 * i)))
    HDIV[i] = int(np.rint((180 << 12) / (6.0 * i)))


# ---- numpy models used for the exhaustive validation -----------------------------------
def D(x, n):
    return (x + (1 << (n - 1))) >> n


def cdiv(a, b):
    return np.where(a >= 0, a // b, -((-a) // b))


def bgr2lab(img):
    B = G[img[..., 0]]; Gg = G[img[..., 1]]; R = G[img[..., 2]]
    fX = C[D(R * 1777 + Gg * 1541 + B * 778, 12)]
    fY = C[D(R * 871 + Gg * 2929 + B * 296, 12)]
    fZ = C[D(R * 73 + Gg * 448 + B * 3575, 12)]
    L = D(296 * fY - 1336934, 15)
    A = D(500 * (fX - fY) + 128 * 32768, 15)
    Bb = D(200 * (fY - fZ) + 128 * 32768, 15)
    return np.stack([np.clip(L, 0, 255), np.clip(A, 0, 255), np.clip(Bb, 0, 255)], -1).astype(np.uint8)


def lab2bgr(lab):
    L = lab[..., 0].astype(np.int64); a = lab[..., 1].astype(np.int64); b = lab[..., 2].astype(np.int64)
    yy = Y[L]; fy = IFY[L]
    adiv = ((5 * a * 53687 + 128) >> 13) - 4194
    bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1

    def T(v):
        return np.where(v <= 3390, cdiv(v * 108, 841) - 290, cdiv(cdiv(v * v, BASE) * v, BASE))
    x = T(fy + adiv); z = T(fy - bdiv)
    outs = []
    for c0, c1, c2 in ((12615, -6296, -2223), (-3773, 7684, 185), (217, -836, 4715)):
        outs.append(IG[np.clip(D(c0 * x + c1 * yy + c2 * z, 14), 0, 4095)])
    R, Gg, B = outs
    return np.stack([B, Gg, R], -1).astype(np.uint8)


def bgr2hsv(img):
    b = img[..., 0].astype(np.int64); g = img[..., 1].astype(np.int64); r = img[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r); d = v - np.minimum(np.minimum(b, g), r)
    s = (d * SDIV[v] + 2048) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (h * HDIV[d] + 2048) >> 12
    h = np.where(h < 0, h + 180, h)
    return np.stack([np.clip(h, 0, 255), s, v], -1).astype(np.uint8)


def main():
    v = np.arange(256, dtype=np.uint8)
    a, b, c = np.meshgrid(v, v, v, indexing="ij")
    cube = np.stack([a, b, c], -1).reshape(4096, 4096, 3)
    for name, fn, code in (("BGR2Lab", bgr2lab, cv2.COLOR_BGR2Lab), ("Lab2BGR", lab2bgr, cv2.COLOR_Lab2BGR),
                           ("BGR2HSV", bgr2hsv, cv2.COLOR_BGR2HSV)):
        bad = int((fn(cube) != cv2.cvtColor(cube, code)).sum())
        print(f"{name}: {bad} mismatching values over 2^24 triples (cv2 {cv2.__version__})")
        if bad:
            sys.exit("table validation failed")

    def arr(ctype, name, values, per_line=16):
        vals = [str(int(x)) for x in values]
        lines = [", ".join(vals[i:i + per_line]) for i in range(0, len(vals), per_line)]
        return f"static const {ctype} {name}[{len(vals)}] = {{\n  " + ",\n  ".join(lines) + "\n};\n"

    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_cv_tables.py -- do not edit.\n#pragma once\n"
                f"// Validated exhaustively (2^24 triples, BGR2Lab / Lab2BGR / BGR2HSV) against cv2 {cv2.__version__}.\n"
                "// OpenCV 8-bit colour-conversion constants (imgproc color_lab.cpp / color_hsv.cpp).\n")
        f.write(arr("unsigned short", "kSrgbGammaTab", G))
        f.write(arr("unsigned short", "kLabCbrtTab", C))
        f.write(arr("unsigned int", "kLabToYF", [(int(IFY[i]) << 16) | int(Y[i]) for i in range(256)], 8))
        f.write(arr("unsigned char", "kSrgbInvGammaTab", IG, 32))
        f.write(arr("int", "kHsvSdiv", SDIV, 8))
        f.write(arr("int", "kHsvHdiv", HDIV, 8))
        # cv::log on CV_32F (core mathfuncs, table + polynomial, not libm) for the 256 values an 8-bit
        # channel can take, as IEEE-754 bit patterns; entry 0 is -inf.  Position-independent [checked below].
        x = np.arange(256, dtype=np.float32)
        with np.errstate(all="ignore"):
            lg = cv2.log(x.reshape(1, -1)).ravel()
            for n, off in ((1, 5), (7, 100), (33, 200), (256, 0)):
                assert np.array_equal(cv2.log(np.ascontiguousarray(x[off:off + n]).reshape(1, -1)).ravel(), lg[off:off + n])
        f.write("// cv::log(float(i)), i = 0..255, as raw IEEE-754 bits (reinterpret as float)\n")
        f.write(arr("unsigned int", "kCvLogTabBits", lg.view(np.uint32), 8))
        # cv::KalmanFilter(2, 2, 0, CV_32F) with A = H = Q = I, R = 10 I, P0 = 0 (ccc.cpp:176-203): the gain
        # is data-independent, stays a multiple of I and reaches its fp32 fixed point after < 40 steps.
        # Tabulated from cv2 because correct() solves through an fp32 SVD whose rounding has no closed form.
        kf = cv2.KalmanFilter(2, 2, 0, cv2.CV_32F)
        kf.transitionMatrix = np.eye(2, dtype=np.float32); kf.processNoiseCov = np.eye(2, dtype=np.float32)
        kf.measurementMatrix = np.eye(2, dtype=np.float32); kf.measurementNoiseCov = 10 * np.eye(2, dtype=np.float32)
        gains = []
        for k in range(64):
            kf.predict(); kf.correct(np.array([[k % 7], [3 * k % 11]], np.float32))
            g = kf.gain
            assert g[0, 1] == 0 and g[1, 0] == 0 and g[0, 0] == g[1, 1]
            gains.append(g[0, 0])
        gains = np.array(gains, np.float32)
        assert np.all(gains[40:] == gains[40])
        f.write("// gain of the k-th cv::KalmanFilter::correct() call of the CCC tracker (fixed point from entry 39 on)\n")
        f.write(arr("unsigned int", "kCccKalmanGainBits", gains[:40].view(np.uint32), 8))
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
