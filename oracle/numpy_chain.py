"""Second oracle -- TEST INFRASTRUCTURE ONLY: a from-scratch numpy restatement of the arithmetic the
reference's CPU path reaches through OpenCV, with NO cv2 call in the pixel path.

``oracle/cv2_oracle.py`` replays the reference call for call through the cv2 wheel; this module restates what
those calls compute, so that the two oracles pin each other (tests/test_oracle.py compares them on seeded
frames, every stage) and so that parity does not rest on a library being bit-stable across builds.
Only ``tests/`` may import it.

Each function cites the reference call site it restates and the appendix that pins the arithmetic.
The 8-bit colour-conversion tables are rebuilt from their published formulas (OpenCV imgproc
color_lab.cpp ``initLabTabs``, color_hsv.cpp) exactly as ``tools/gen_cv_tables.py`` does.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
BASE = 16384


# ---- OpenCV 8-bit colour tables (A.7-A.9) ---------------------------------------------------------------
def _gamma_fwd(x):
    x = x.astype(np.float32)
    lo = x * (f32(1) / f32(12.92))
    hi = np.power((x.astype(np.float64) + 0.055) / 1.055, 2.4).astype(np.float32)
    return np.where(x <= f32(0.04045), lo, hi).astype(np.float32)


G = np.rint(f32(2040) * _gamma_fwd((np.arange(256) / f32(255)).astype(np.float32))).astype(np.int64)
_t = (np.arange(2041).astype(np.float32) / f32(2040))
C = np.rint(32768 * np.where(_t < f32(0.008856), _t.astype(np.float64) * 7.787 + 16 / 116,
                             np.cbrt(_t.astype(np.float64)))).astype(np.int64)
C[49] = 9454  # OpenCV's softfloat evaluation lands on the other side of a .5 tie (tools/gen_cv_tables.py EXCEPTIONS)
Y = np.zeros(256, np.int64)
IFY = np.zeros(256, np.int64)
for _i in range(256):
    if _i <= 20:
        Y[_i] = int(np.rint(f32(_i * BASE * 20 * 9) / f32(17 * 29 * 29 * 29)))
        IFY[_i] = int(np.rint(f32(BASE) * (f32(16) / f32(116) + f32(_i * 5) / f32(3 * 17 * 29))))
    else:
        _fy = f32(f32(_i * 100 * BASE) / f32(255 * 116) + f32(16 * BASE) / f32(116))
        IFY[_i] = int(np.rint(_fy))
        Y[_i] = int(np.rint(f32(f32(_fy * _fy) * _fy) / f32(BASE * BASE)))


def _inv_gamma(x):
    x = x.astype(np.float64)
    return np.where(x <= 0.0031308, x * 12.92, 1.055 * np.power(x, 1 / 2.4) - 0.055)


IG = np.clip(np.rint(255 * _inv_gamma(np.arange(4096) / 4096.0)), 0, 255).astype(np.int64)
SDIV = np.zeros(256, np.int64)
HDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    SDIV[_i] = int(np.rint((255 << 12) / (1.0 * _i)))
    HDIV[_i] = int(np.rint((180 << 12) / (6.0 * _i)))


def _D(x, n):
    return (x + (1 << (n - 1))) >> n


def _cdiv(a, b):  # C integer division (truncation toward zero)
    return np.where(a >= 0, a // b, -((-a) // b))


def sat_u8(x):
    """saturate_cast<uchar>: round half to even, clamp (NaN -> 0)."""
    x = np.nan_to_num(np.asarray(x, np.float64), nan=0.0, posinf=255.0, neginf=0.0)
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


# ---- debayer.cpp:45-79 == cv::demosaicing(bilinear) + R/B swap (A.1) -------------------------------------
# colour at (row % 2, col % 2): 0 = B, 1 = G, 2 = R
CFA = {"bayer_bggr8": ((0, 1), (1, 2)), "bayer_rggb8": ((2, 1), (1, 0)),
       "bayer_gbrg8": ((1, 0), (2, 1)), "bayer_grbg8": ((1, 2), (0, 1))}


def debayer(raw: np.ndarray, encoding: str) -> np.ndarray:
    pat = CFA[encoding]
    H, W = raw.shape
    p = raw.astype(np.int64)
    out = np.zeros((H, W, 3), np.int64)
    c = p[1:-1, 1:-1]
    n, s, w, e = p[:-2, 1:-1], p[2:, 1:-1], p[1:-1, :-2], p[1:-1, 2:]
    nw, ne, sw, se = p[:-2, :-2], p[:-2, 2:], p[2:, :-2], p[2:, 2:]
    cross = (n + s + w + e + 2) >> 2
    diag = (nw + ne + sw + se + 2) >> 2
    horiz = (w + e + 1) >> 1
    vert = (n + s + 1) >> 1
    yy, xx = np.mgrid[1:H - 1, 1:W - 1]
    inner = np.zeros((H - 2, W - 2, 3), np.int64)
    for py in range(2):
        for px in range(2):
            m = ((yy & 1) == py) & ((xx & 1) == px)
            col = pat[py][px]
            if col == 1:  # green site: the row's other colour comes from the horizontal neighbours
                hcol = pat[py][px ^ 1]
                inner[..., 1][m] = c[m]
                inner[..., hcol][m] = horiz[m]
                inner[..., 2 - hcol][m] = vert[m]
            else:
                inner[..., col][m] = c[m]
                inner[..., 1][m] = cross[m]
                inner[..., 2 - col][m] = diag[m]
    out[1:-1, 1:-1] = inner
    out[1:-1, 0] = out[1:-1, 1]; out[1:-1, -1] = out[1:-1, -2]   # border: copy the neighbouring interior pixel
    out[0] = out[1]; out[-1] = out[-2]
    return out.astype(np.uint8)


# ---- flip.cpp:37-58 (A.1b) --------------------------------------------------------------------------------
def flip(img: np.ndarray, angle: int) -> np.ndarray:
    if angle == 90:
        return np.ascontiguousarray(np.rot90(img, -1))
    if angle == 180:
        return np.ascontiguousarray(img[::-1, ::-1])
    if angle == 270:
        return np.ascontiguousarray(np.rot90(img, 1))
    return img


# ---- white_balance.cpp:73-136 (A.2) -------------------------------------------------------------------------
def white_balance_pca(img: np.ndarray) -> np.ndarray:
    b = img[..., 0].astype(np.int64); g = img[..., 1].astype(np.int64); r = img[..., 2].astype(np.int64)
    luts = []
    with np.errstate(all="ignore"):
        for x in (b, r):
            s1, s2, m1 = f32(float(x.sum())), f32(float((x * x).sum())), f32(float(x.max()))
            m2 = f32(float(int(x.max()) ** 2))
            sg, mg = f32(float(g.sum())), f32(float(g.max()))
            det = f32(f32(s2 * m1) - f32(m2 * s1))
            inv = f32(f32(1) / det)
            i00, i01, i10, i11 = f32(m1 * inv), f32(f32(-s1) * inv), f32(f32(-m2) * inv), f32(s2 * inv)
            alpha = f32(f32(i00 * sg) + f32(i01 * mg))
            beta = f32(f32(i10 * sg) + f32(i11 * mg))
            v = np.arange(256, dtype=np.float64)
            y = (v * v * np.float64(alpha) + v * np.float64(beta)).astype(np.float32)  # addWeighted: double MAC, one rounding
            luts.append(sat_u8(np.minimum(y, f32(255))))                               # THRESH_TRUNC then convertTo(CV_8U)
    out = img.copy()
    out[..., 0] = luts[0][img[..., 0]]
    out[..., 2] = luts[1][img[..., 2]]
    return out


# ---- color_calibration.cpp:91-104 (A.4) ----------------------------------------------------------------------
def color_calibration(img: np.ndarray, matrix, bias) -> np.ndarray:
    M = np.asarray(matrix, np.float64).reshape(3, 3).astype(np.float32)
    p = img.astype(np.float32)
    out = np.empty(img.shape, np.float32)
    for j in range(3):
        t0 = p[..., 0] * M[j, 0]; t1 = p[..., 1] * M[j, 1]; t2 = p[..., 2] * M[j, 2]   # separately rounded products
        out[..., j] = ((t0 + t1) + t2) + f32(bias[j])
    return sat_u8(out)


# ---- gamma_correction.cpp:35-56 (A.5) -----------------------------------------------------------------------
def gamma(img: np.ndarray, k: float) -> np.ndarray:
    i = np.arange(256)
    f = (i / 255.0).astype(np.float32)
    f = np.power(f.astype(np.float64), k).astype(np.float32)
    return sat_u8(f.astype(np.float64) * 255.0)[img]


# ---- vignetting_correction.cpp:32-93 (A.6-A.8) ----------------------------------------------------------------
def vignetting_mask(rows: int, cols: int, scale: float, a2: float, a4: float) -> np.ndarray:
    j = np.arange(cols, dtype=np.float64) - cols / 2.0
    i = np.arange(rows, dtype=np.float64) - rows / 2.0
    r = np.sqrt(np.power(j, 2)[None, :] + np.power(i, 2)[:, None])
    k = (np.power(r, 2) * a2 + np.power(r, 4) * a4).astype(np.float32)
    kmax = k.max()
    if float(kmax) > 0:
        k = k * f32(1.0 / float(kmax))
    k = k * f32(scale)
    return k + f32(1.0)


def bgr2lab(img):
    B = G[img[..., 0]]; Gg = G[img[..., 1]]; R = G[img[..., 2]]
    fX = C[_D(R * 1777 + Gg * 1541 + B * 778, 12)]
    fY = C[_D(R * 871 + Gg * 2929 + B * 296, 12)]
    fZ = C[_D(R * 73 + Gg * 448 + B * 3575, 12)]
    L = _D(296 * fY - 1336934, 15)
    A = _D(500 * (fX - fY) + 128 * 32768, 15)
    Bb = _D(200 * (fY - fZ) + 128 * 32768, 15)
    return np.stack([np.clip(L, 0, 255), np.clip(A, 0, 255), np.clip(Bb, 0, 255)], -1).astype(np.uint8)


def lab2bgr(lab):
    L = lab[..., 0].astype(np.int64); a = lab[..., 1].astype(np.int64); b = lab[..., 2].astype(np.int64)
    yy = Y[L]; fy = IFY[L]
    adiv = ((5 * a * 53687 + 128) >> 13) - 4194
    bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1

    def T(v):
        return np.where(v <= 3390, _cdiv(v * 108, 841) - 290, _cdiv(_cdiv(v * v, BASE) * v, BASE))
    x = T(fy + adiv); z = T(fy - bdiv)
    outs = []
    for c0, c1, c2 in ((12615, -6296, -2223), (-3773, 7684, 185), (217, -836, 4715)):
        outs.append(IG[np.clip(_D(c0 * x + c1 * yy + c2 * z, 14), 0, 4095)])
    R, Gg, B = outs
    return np.stack([B, Gg, R], -1).astype(np.uint8)


def vignetting(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    lab = bgr2lab(img)
    lab[..., 0] = sat_u8(lab[..., 0].astype(np.float32) * mask)  # cv::multiply fp32, convertTo(CV_8U)
    return lab2bgr(lab)


# ---- color_enhancer.cpp:38-47 (A.9) ---------------------------------------------------------------------------
def bgr2hsv(img):
    b = img[..., 0].astype(np.int64); g = img[..., 1].astype(np.int64); r = img[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r); d = v - np.minimum(np.minimum(b, g), r)
    s = (d * SDIV[v] + 2048) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (h * HDIV[d] + 2048) >> 12
    h = np.where(h < 0, h + 180, h)
    return np.stack([np.clip(h, 0, 255), s, v], -1).astype(np.uint8)


def _fma32(a, b, c):
    """fmaf on fp32 arrays: exact product and sum in float64 (48 + 24 significant bits fit), one rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def hsv2bgr(hsv):
    """HSV2RGB_b as the optimised OpenCV build runs it: fused 1 - s*f, results truncated in the 32-pixel
    vector chunks of each row and rounded in the scalar tail (width % 32 pixels)."""
    h = hsv[..., 0].astype(np.float32) * f32(6.0 / 180.0)
    h = np.where(h >= f32(6), h - f32(6), h).astype(np.float32)
    sec = np.floor(h)
    f = (h - sec).astype(np.float32)
    sec = np.clip(sec.astype(np.int64), 0, 5)
    s = hsv[..., 1].astype(np.float32) * f32(1.0 / 255.0)
    v = hsv[..., 2].astype(np.float32) * f32(1.0 / 255.0)
    one = np.ones_like(s)
    t = [v, v * (one - s), v * _fma32(-s, f, one), v * _fma32(-s, (one - f).astype(np.float32), one)]
    tab = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    stack = np.stack(t, -1)
    out = np.empty(hsv.shape, np.float32)
    for c in range(3):
        out[..., c] = np.take_along_axis(stack, tab[sec, c][..., None], -1)[..., 0] * f32(255)
    W = hsv.shape[1]
    res = np.trunc(out).astype(np.int64) & 255
    tail = (np.arange(W) >= (W & ~31))
    res[:, tail] = sat_u8(out[:, tail])
    return res.astype(np.uint8)


def color_enhancer(img, hue_gain, saturation_gain, value_gain):
    hsv = bgr2hsv(img).astype(np.float64)
    for c, gain in enumerate((hue_gain, saturation_gain, value_gain)):
        hsv[..., c] = hsv[..., c] * float(gain)   # cv::multiply(u8, Scalar): double product
    return hsv2bgr(sat_u8(hsv))


# ---- undistortion.cpp:240-245 == cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) (A.10) -------------------------------
def remap(img: np.ndarray, mx: np.ndarray, my: np.ndarray) -> np.ndarray:
    def fix(m):
        v = m.astype(np.float32) * f32(32)
        bad = ~np.isfinite(v) | (np.abs(v) >= 2147483648.0)
        return np.where(bad, -2 ** 31, np.rint(np.where(bad, 0, v))).astype(np.int64)
    sx, sy = fix(mx), fix(my)
    ix, iy, ax, ay = sx >> 5, sy >> 5, sx & 31, sy & 31
    H, W = img.shape[:2]
    src = img.reshape(H, W, -1).astype(np.int64)
    acc = np.full(mx.shape + (src.shape[2],), 16384, np.int64)
    for dy, dx, w in ((0, 0, (32 - ay) * (32 - ax) * 32), (0, 1, (32 - ay) * ax * 32), (1, 0, ay * (32 - ax) * 32), (1, 1, ay * ax * 32)):
        yy, xx = iy + dy, ix + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        tap = np.where(ok[..., None], src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0)
        acc += w[..., None] * tap
    out = np.clip(acc >> 15, 0, 255).astype(np.uint8)
    return out.reshape(mx.shape + img.shape[2:])
