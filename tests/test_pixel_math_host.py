"""CPU-only: the kernels' per-pixel arithmetic (host build of pixel_math.cuh/frame_math.cuh)
against cv2, exhaustively where the domain is finite (SURVEY 8c 'self-made known-answer tests')."""
import ctypes

import cv2
import numpy as np
import pytest

from conftest import CC_EXAMPLE, cube
from oracle import cv2_oracle as O

P = ctypes.c_void_p


def _run3(fn, img, with_width=False):
    out = np.empty_like(img)
    if with_width:
        fn(ctypes.c_long(img.shape[0] * img.shape[1]), ctypes.c_int(img.shape[1]), P(img.ctypes.data), P(out.ctypes.data))
    else:
        fn(ctypes.c_long(img.shape[0] * img.shape[1]), P(img.ctypes.data), P(out.ctypes.data))
    return out


@pytest.mark.parametrize("name,code", [("hs_bgr2lab", cv2.COLOR_BGR2Lab), ("hs_lab2bgr", cv2.COLOR_Lab2BGR),
                                       ("hs_bgr2hsv", cv2.COLOR_BGR2HSV), ("hs_hsv2bgr", cv2.COLOR_HSV2BGR)])
def test_colour_conversions_exhaustive(hostsim, name, code):
    img = cube()
    got = _run3(getattr(hostsim, name), img, with_width=(name == "hs_hsv2bgr"))
    ref = cv2.cvtColor(img, code)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("width", [31, 45, 100])
@pytest.mark.parametrize("name,code", [("hs_bgr2lab", cv2.COLOR_BGR2Lab), ("hs_lab2bgr", cv2.COLOR_Lab2BGR),
                                       ("hs_bgr2hsv", cv2.COLOR_BGR2HSV), ("hs_hsv2bgr", cv2.COLOR_HSV2BGR)])
def test_colour_conversions_exhaustive_in_row_tails(hostsim, name, code, width):
    """cv2 finishes each row's last (width % 32) pixels with scalar code; HSV2BGR rounds there
    instead of truncating.  All 2^24 triples again, laid out so that they land in row tails."""
    flat = cube().reshape(-1, 3)
    pad = (-len(flat)) % width
    img = np.ascontiguousarray(np.concatenate([flat, flat[:pad]]).reshape(-1, width, 3))
    got = _run3(getattr(hostsim, name), img, with_width=(name == "hs_hsv2bgr"))
    ref = cv2.cvtColor(img, code)
    assert int((got != ref).sum()) == 0


def _chain(hostsim, stages, img, mask=None, cc=None, bias=(0, 0, 0), enh=(1, 1, 1), wb=None, gamma=None):
    n = img.shape[0] * img.shape[1]
    cc = np.asarray(cc if cc is not None else np.eye(3).ravel(), np.float64).astype(np.float32)
    bias = np.asarray(bias, np.float64).astype(np.float32)
    enh = np.asarray(enh, np.float64)
    wb = np.ascontiguousarray(wb if wb is not None else np.tile(np.arange(256, dtype=np.uint8), 3))
    gamma = np.ascontiguousarray(gamma if gamma is not None else np.arange(256, dtype=np.uint8))
    out = np.empty_like(img)
    m = None if mask is None else np.ascontiguousarray(mask, np.float32)
    hostsim.hs_chain(ctypes.c_uint(stages), ctypes.c_long(n), ctypes.c_int(img.shape[1]), P(img.ctypes.data),
                     P(m.ctypes.data) if m is not None else None, P(cc.ctypes.data), P(bias.ctypes.data),
                     P(enh.ctypes.data), P(wb.ctypes.data), P(gamma.ctypes.data), P(out.ctypes.data))
    return out


@pytest.mark.parametrize("matrix,bias", [(CC_EXAMPLE, (0, 0, 0)),
                                         ([1.6016861, -1.0104847, 0.2213647, -0.012442476, 1.0841908, -0.07573348,
                                           0.09123942, -0.8107389, 1.8687756], (3.25, -7.5, 0.49))])
def test_colour_calibration_exhaustive(hostsim, matrix, bias):
    img = cube()
    got = _chain(hostsim, 2, img, cc=matrix, bias=bias)
    ref = O.color_calibration(img, matrix, bias)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("gains", [(1.0, 1.2, 1.0), (1.0, 1.5, 1.0), (1.1, 0.8, 1.3), (1.5, 2.0, 0.7)])
def test_enhancer_exhaustive(hostsim, gains):
    img = cube()
    got = _chain(hostsim, 16, img, enh=gains)
    ref = O.color_enhancer(img, *gains)
    assert int((got != ref).sum()) == 0


def test_vignetting_stage(hostsim, oracle_built):
    rng = np.random.default_rng(5)
    rows, cols = 540, 720
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    for (s, a2, a4) in [(1.5, 1e-3, 1e-6), (0.7, 2e-3, 0.0), (3.0, 1e-4, 1e-7)]:
        mask = O.vignetting_mask(rows, cols, s, a2, a4)
        got = _chain(hostsim, 8, img, mask=mask)
        ref = O.vignetting(img, mask)
        assert int((got != ref).sum()) == 0


def test_vignetting_L_times_mask_dense(hostsim):
    """2M colour triples x random mask values in the range the mask can take."""
    img = cube()[:512]
    rng = np.random.default_rng(7)
    mask = rng.uniform(1.0, 2.6, img.shape[:2]).astype(np.float32)
    got = _chain(hostsim, 8, img, mask=mask)
    ref = O.vignetting(img, mask)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("rows,cols", [(480, 640), (270, 362), (33, 17)])
def test_gamma_and_full_chain_random(hostsim, oracle_built, rows, cols):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    mask = O.vignetting_mask(rows, cols, 1.5, 1e-3, 1e-6)
    glut = O.gamma_lut(0.8)
    wb = np.concatenate([rng.permutation(256).astype(np.uint8), np.arange(256, dtype=np.uint8),
                         np.sort(rng.integers(0, 256, 256)).astype(np.uint8)])
    got = _chain(hostsim, 31, img, mask=mask, cc=CC_EXAMPLE, enh=(1.0, 1.2, 1.0), wb=wb, gamma=glut)
    ref = cv2.merge([cv2.LUT(img[..., 0], wb[:256]), cv2.LUT(img[..., 1], wb[256:512]), cv2.LUT(img[..., 2], wb[512:])])
    ref = O.color_calibration(ref, CC_EXAMPLE, (0, 0, 0))
    ref = O.gamma(ref, 0.8)
    ref = O.vignetting(ref, mask)
    ref = O.color_enhancer(ref, 1.0, 1.2, 1.0)
    assert int((got != ref).sum()) == 0


CFA_ID = {"bayer_rggb8": 0, "bayer_grbg8": 1, "bayer_gbrg8": 2, "bayer_bggr8": 3}


@pytest.mark.parametrize("enc", list(CFA_ID))
@pytest.mark.parametrize("shape", [(480, 640), (11, 13), (10, 12), (3, 3), (4, 9), (33, 130)])
@pytest.mark.parametrize("angle", [0, 90, 180, 270])
def test_demosaic_and_flip(hostsim, enc, shape, angle):
    rng = np.random.default_rng(abs(hash((enc, shape, angle))) % (1 << 32))
    raw = rng.integers(0, 256, shape, dtype=np.uint8)
    ref, _ = O.debayer(raw, enc)
    ref = O.flip(ref, angle)
    for mode in (0, 1, 2):
        out = np.empty(ref.shape, np.uint8)
        hostsim.hs_demosaic(P(raw.ctypes.data), shape[0], shape[1], CFA_ID[enc], angle, mode, P(out.ctypes.data))
        assert out.shape == ref.shape
        assert int((out != ref).sum()) == 0, (enc, shape, angle, mode)


@pytest.mark.parametrize("ch", [1, 3])
def test_remap_random_maps(hostsim, ch):
    rng = np.random.default_rng(3)
    rows, cols = 97, 131
    src = rng.integers(0, 256, (rows, cols, ch) if ch == 3 else (rows, cols), dtype=np.uint8)
    orows, ocols = 120, 150
    mx = rng.uniform(-3, cols + 2, (orows, ocols)).astype(np.float32)
    my = rng.uniform(-3, rows + 2, (orows, ocols)).astype(np.float32)
    # exact grid points, half-way points, far out-of-range and infinities
    mx[0, :10] = np.arange(10); my[0, :10] = 5.0
    mx[1, :10] = np.arange(10) + 0.5; my[1, :10] = 4.5
    mx[2, :4] = [1e9, -1e9, np.inf, -np.inf]; my[2, :4] = [3, 3, 3, 3]
    mx[3, :4] = [cols - 1, cols - 1.0 + 1 / 64, cols - 0.5, -0.984375]; my[3, :4] = [rows - 1, rows - 0.5, 0, -0.5]
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out = np.empty_like(ref)
    hostsim.hs_remap(P(src.ctypes.data), rows, cols, ch, P(mx.ctypes.data), P(my.ctypes.data), orows, ocols,
                     P(out.ctypes.data))
    assert int((out != ref).sum()) == 0


def test_remap_from_bgrx_intermediate_matches_cv2(hostsim):
    """The 4-byte-per-pixel gather (one load per tap, dp4a horizontal step) is the same function."""
    rng = np.random.default_rng(77)
    rows, cols = 97, 131
    src = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    bgrx = np.zeros((rows, cols, 4), np.uint8); bgrx[..., :3] = src
    orows, ocols = 120, 152
    mx = rng.uniform(-3, cols + 2, (orows, ocols)).astype(np.float32)
    my = rng.uniform(-3, rows + 2, (orows, ocols)).astype(np.float32)
    mx[0, :10] = np.arange(10); my[0, :10] = 5.0
    mx[1, :10] = np.arange(10) + 0.5; my[1, :10] = 4.5
    mx[2, :4] = [1e9, -1e9, np.inf, -np.inf]; my[2, :4] = [3, 3, 3, 3]
    mx[3, :4] = [cols - 1, cols - 1.0 + 1 / 64, cols - 0.5, -0.984375]; my[3, :4] = [rows - 1, rows - 0.5, 0, -0.5]
    mx[4, :6] = [cols - 2, cols - 1.5, -1, -0.5, 0, cols - 1]; my[4, :6] = [rows - 2, rows - 1.5, 0, -1, -0.5, rows - 1]
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out = np.empty_like(ref)
    hostsim.hs_remap_bgrx(P(bgrx.ctypes.data), rows, cols, P(mx.ctypes.data), P(my.ctypes.data), orows, ocols, P(out.ctypes.data))
    assert int((out != ref).sum()) == 0
    # all 256 x 256 weight combinations on extreme pixel values
    src2 = rng.choice(np.array([0, 1, 127, 128, 254, 255], np.uint8), (8, 8, 3))
    b2 = np.zeros((8, 8, 4), np.uint8); b2[..., :3] = src2
    fx, fy = np.meshgrid(np.arange(32, dtype=np.float32) / 32, np.arange(32, dtype=np.float32) / 32)
    mx2 = np.tile(fx, (6, 6)) + np.repeat(np.repeat(np.arange(6, dtype=np.float32)[None, :], 6, 0), 32, 0).repeat(32, 1)
    my2 = np.tile(fy, (6, 6)) + np.repeat(np.repeat(np.arange(6, dtype=np.float32)[:, None], 6, 1), 32, 0).repeat(32, 1)
    ref2 = cv2.remap(src2, mx2, my2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out2 = np.empty_like(ref2)
    hostsim.hs_remap_bgrx(P(b2.ctypes.data), 8, 8, P(np.ascontiguousarray(mx2).ctypes.data), P(np.ascontiguousarray(my2).ctypes.data),
                          192, 192, P(out2.ctypes.data))
    assert int((out2 != ref2).sum()) == 0


def test_remap_packed_fixed_point_map_matches_cv2(hostsim):
    """the 4-byte packed map (int16 displacements in 1/32 px) gives cv::remap's result exactly;
    maps whose displacement does not fit are refused (the library then keeps the float map)."""
    rng = np.random.default_rng(78)
    rows, cols = 300, 420
    src = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    bgrx = np.zeros((rows, cols, 4), np.uint8); bgrx[..., :3] = src
    xs, ys = np.meshgrid(np.arange(cols, dtype=np.float32), np.arange(rows, dtype=np.float32))
    mx = (xs + rng.uniform(-200, 200, xs.shape)).astype(np.float32)
    my = (ys + rng.uniform(-150, 150, ys.shape)).astype(np.float32)
    mx[0, :6] = [np.nan, np.inf, -np.inf, 1e9, -1e9, -1.0]; my[1, :4] = [np.nan, -1.0, rows - 0.5, rows]
    mx[2, :4] = [-1.03125, -0.96875, cols - 1, cols - 0.03125]
    mxc, myc = mx.copy(), my.copy()
    mxc[np.isnan(mxc)] = -1e9; myc[np.isnan(myc)] = -1e9   # what the library uploads (rip_api.cu build_host_map)
    ref = cv2.remap(src, mxc, myc, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out = np.empty_like(ref)
    ok = hostsim.hs_remap_bgrx_packed(P(bgrx.ctypes.data), rows, cols, P(mxc.ctypes.data), P(myc.ctypes.data), rows, cols, P(out.ctypes.data))
    assert ok == 1 and int((out != ref).sum()) == 0
    far = mxc.copy(); far[10, 10] = 5.0 + 1100.0   # inside a (hypothetical) wide image but > 1023 px away: must be refused
    wide = np.zeros((rows, 2000, 4), np.uint8)
    ok = hostsim.hs_remap_bgrx_packed(P(wide.ctypes.data), rows, 2000, P(far.ctypes.data), P(myc.ctypes.data), rows, cols, P(out.ctypes.data))
    assert ok == 0


def test_remap_tile_kernel_arithmetic_equals_gather_arithmetic(hostsim):
    """k_remap_tile's test-free path (address = row base + dyi * box_w + dxi from the packed entry, taps from a zero-padded
    box, vertical weights x 64, two byte permutes) gives the value of the gather path for every fraction pair (ax, ay),
    for negative displacements, and at the image border (zeros from the box where the gather drops a tap)."""
    import ctypes
    rng = np.random.default_rng(79)
    rows, cols, BW, BH = 40, 56, 176, 48
    src = np.zeros((rows, cols, 4), np.uint8); src[..., :3] = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    src32 = src.view(np.uint32).reshape(rows, cols)
    bx0, by0 = -8, -3                      # a box that sticks out of the image on every side
    box = np.zeros((BH, BW), np.uint32)
    box[-by0:-by0 + rows, -bx0:-bx0 + cols] = src32
    hostsim.hs_remap_tile_pixel.restype = ctypes.c_uint32
    hostsim.hs_remap_pack.restype = ctypes.c_uint32
    hostsim.hs_remap_pixel_packed.restype = ctypes.c_uint32
    ok = ctypes.c_int()
    n = 0
    for y in (0, 5, 17):
        for x in (0, 3, 30):
            for fy in range(32):
                for fx in range(32):
                    # source positions from just outside the upper-left corner to just outside the lower-right one
                    for base_x, base_y in ((-1.0, -1.0), (4.0, 2.0), (cols - 2.0, rows - 2.0), (cols - 1.0, rows - 1.0), (20.0, 0.0)):
                        mx, my = base_x + fx / 32.0, base_y + fy / 32.0
                        e = hostsim.hs_remap_pack(ctypes.c_float(mx), ctypes.c_float(my), x, y, rows, cols, ctypes.byref(ok))
                        assert ok.value == 1
                        if (e & 0xffff) == 0x8000:
                            continue       # "far": the tile is not flagged, the kernel's tested path handles it
                        a = hostsim.hs_remap_tile_pixel(P(box.ctypes.data), BW, bx0, by0, ctypes.c_uint32(e), x, y)
                        b = hostsim.hs_remap_pixel_packed(P(src32.ctypes.data), rows, cols, ctypes.c_uint32(e), x, y)
                        assert a == b, (x, y, mx, my, hex(a), hex(b))
                        n += 1
    assert n > 40000


def test_remap_identity_is_exact():
    """cv::remap with an identity map returns the image (guards the 32768-weight corner)."""
    rng = np.random.default_rng(9)
    src = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    xs, ys = np.meshgrid(np.arange(50, dtype=np.float32), np.arange(40, dtype=np.float32))
    assert np.array_equal(cv2.remap(src, xs, ys, cv2.INTER_LINEAR), src)


def _stats(img):
    b = img[..., 0].astype(np.uint64); g = img[..., 1].astype(np.uint64); r = img[..., 2].astype(np.uint64)
    return np.array([b.sum(), (b * b).sum(), r.sum(), (r * r).sum(), g.sum(), b.max(), g.max(), r.max()], np.uint64)


@pytest.mark.parametrize("seed", range(6))
def test_pca_lut_matches_oracle(hostsim, seed):
    rng = np.random.default_rng(seed)
    rows, cols = 300, 400
    base = rng.integers(0, 256, (rows, cols, 3)).astype(np.float32)
    base *= np.array([rng.uniform(0.3, 1.0), rng.uniform(0.5, 1.0), rng.uniform(0.3, 1.0)], np.float32)
    img = base.astype(np.uint8)
    st = _stats(img)
    lut_b = np.empty(256, np.uint8); lut_r = np.empty(256, np.uint8); coeff = np.empty(4, np.float32)
    hostsim.hs_pca_lut(P(st.ctypes.data), P(lut_b.ctypes.data), P(lut_r.ctypes.data), P(coeff.ctypes.data))
    (cb, cr), _ = O.pca_coefficients(img)
    assert coeff[0] == cb[0] and coeff[1] == cb[1] and coeff[2] == cr[0] and coeff[3] == cr[1]
    ref = O.white_balance_pca(img)
    got = img.copy()
    got[..., 0] = lut_b[img[..., 0]]
    got[..., 2] = lut_r[img[..., 2]]
    assert np.array_equal(got, ref)
    # the LUT itself over all 256 inputs (not only the values present in the frame)
    x = np.arange(256, dtype=np.float32).reshape(1, 256)
    for lut, (al, be) in ((lut_b, cb), (lut_r, cr)):
        y = cv2.addWeighted(x * x, float(al), x, float(be), 0.0)
        _, y = cv2.threshold(y, 255, 255, cv2.THRESH_TRUNC)
        assert np.array_equal(O.to_u8(y).ravel(), lut)


@pytest.mark.parametrize("gain", [1.0, 1.0164, 1.366838, 2.117, 0.9, 3.7, 1.1, 1.3, 0.7])
def test_gain_lut(hostsim, gain):
    g = float(np.float32(gain))
    lut = np.empty(256, np.uint8)
    hostsim.hs_gain_lut(ctypes.c_float(g), P(lut.ctypes.data))
    x = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    ref = cv2.multiply(x, (g, g, g, 0.0))
    assert np.array_equal(ref[0, :, 0], lut)


# ---- 16-bit Bayer extension --------------------------------------------------------------------------
def test_reduce16to8_equals_cv2_convert_for_every_value(hostsim):
    v = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    ref = cv2.convertScaleAbs(v, alpha=1.0 / 257.0)   # == Mat::convertTo(CV_8U, 1 / 257.f)
    hostsim.hs_reduce16to8.restype = ctypes.c_int
    got = np.array([hostsim.hs_reduce16to8(int(x)) for x in range(0, 65536, 1)], np.uint8).reshape(256, 256)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("enc", ["bayer_rggb16", "bayer_grbg16", "bayer_gbrg16", "bayer_bggr16"])
@pytest.mark.parametrize("shape", [(480, 640), (11, 13), (10, 12), (3, 3), (4, 9)])
def test_demosaic16_against_cv2(hostsim, enc, shape):
    rng = np.random.default_rng(abs(hash((enc, shape))) % (1 << 32))
    raw = rng.integers(0, 65536, shape, dtype=np.uint16)
    if shape == (480, 640):
        raw[:100] = rng.integers(0, 1024, (100, 640), dtype=np.uint16) << 6   # 10-bit sensor data, MSB-aligned
    out = np.empty(shape + (3,), np.uint8)
    hostsim.hs_demosaic16(P(raw.ctypes.data), ctypes.c_int(shape[0]), ctypes.c_int(shape[1]), ctypes.c_int(CFA_ID[enc[:-2] + "8"]),
                          P(out.ctypes.data))
    ref, renc = O.debayer16(raw, enc)
    assert renc == "bgr8"
    assert int((out != ref).sum()) == 0
