"""CPU-only: the C-ABI library loads, exports every declared symbol, and its host logic
(YAML loading, setter semantics, host-computed tables) matches the reference / the oracle.
No pixel work happens here (that needs the GPU and is covered by test_gpu_parity.py)."""
import os
import re

import cv2
import numpy as np
import pytest

from conftest import CALIB_720, ROOT, scaled_calib
from oracle import cv2_oracle as O
from raw_image_pipeline_b200 import RawImagePipeline, RawImagePipelineError
from raw_image_pipeline_b200 import _lib as L

CONFIG = os.path.join(ROOT, "raw_image_pipeline_b200", "config")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rip_b200.h")).read()
    declared = set(re.findall(r"RIP_API\s+[\w\s\*]+?\b(rip_\w+)\s*\(", header))
    assert len(declared) >= 27
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = RawImagePipeline(False, "", "", "")
    with pytest.raises(RawImagePipelineError) as e:
        p.process(np.zeros((16, 16), np.uint8), "bayer_rggb8")
    assert e.value.code == L.RIP_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_default_constructor_matches_reference_defaults():
    p = RawImagePipeline(False)  # raw_image_pipeline.cpp:16-21 + config/pipeline_params_example.yaml
    assert p.is_debayer_enabled() and not p.is_flip_enabled() and p.is_white_balance_enabled()
    assert not p.is_color_calibration_enabled() and not p.is_gamma_correction_enabled()
    assert not p.is_vignetting_correction_enabled() and p.is_undistortion_enabled()
    assert not p.is_color_enhancer_enabled()  # flag is read from `run_color_enhancer` (App. B-4)
    assert p._get_string("white_balance/method") == "ccc"
    assert p._get_double("undistortion/fov_scale") == 0.8
    # loadParams calls setHueGain three times (-> value_gain_ member = YAML value_gain)
    assert p._get_double("color_enhancer/value_gain_member") == 1.0
    assert p._get_double("color_enhancer/saturation_gain_member") == 1.0  # 1.5 in the YAML never lands
    assert p.get_dist_image_width() == 720 and p.get_dist_image_height() == 540
    assert p.get_dist_distortion_model() == "equidistant"
    assert p.get_rect_distortion_model() == "none"  # undistortion enabled (undistortion.cpp:94-104)
    np.testing.assert_allclose(p.get_color_calibration_matrix().ravel()[:3], [2.4276948, 0.21479778, -0.30818], rtol=1e-7)
    assert p.get_color_calibration_bias().shape == (4, 1)
    assert "Loading raw_image_pipeline params from file" in p.log()


def test_four_argument_constructor_and_missing_files():
    p = RawImagePipeline(False, "", "", "")  # empty calibration path: no camera calibration loaded
    assert p.get_dist_distortion_model() == "none" and p.get_dist_image_width() == 0
    q = RawImagePipeline(False, "/nonexistent/params.yaml", "/nonexistent/calib.yaml", "/nonexistent/color.yaml")
    log = q.log()
    assert "Warning: parameters file doesn't exist" in log
    assert "Warning: Calibration file doesn't exist" in log
    assert "Warning: Color calibration file doesn't exist" in log
    assert q.get_dist_image_width() == 320 and q.get_dist_image_height() == 240  # undistortion.cpp:181
    assert q.get_dist_distortion_model() == "none"


def test_enhancer_setters_are_cross_wired_like_the_reference():
    p = RawImagePipeline(False, "", "", "")
    p.set_color_enhancer_hue_gain(1.1)         # -> value_gain_   (color_enhancer.cpp:23-25)
    p.set_color_enhancer_saturation_gain(1.2)  # -> saturation_gain_
    p.set_color_enhancer_value_gain(1.3)       # -> hue_gain_     (color_enhancer.cpp:31-33)
    assert p._get_double("color_enhancer/value_gain_member") == 1.1
    assert p._get_double("color_enhancer/saturation_gain_member") == 1.2
    assert p._get_double("color_enhancer/hue_gain_member") == 1.3


def test_setters_getters_roundtrip_and_errors():
    p = RawImagePipeline(False, "", "", "")
    for name in ("debayer", "flip", "white_balance", "color_calibration", "gamma_correction", "vignetting_correction",
                 "color_enhancer", "undistortion"):
        getattr(p, "set_" + name)(True)
        assert getattr(p, "is_" + name + "_enabled")()
        getattr(p, "set_" + name)(False)
        assert not getattr(p, "is_" + name + "_enabled")()
    p.set_flip_angle(270)
    assert p._get_int("flip/angle") == 270
    p.set_color_calibration_matrix([1, 2, 3, 4, 5, 6, 7, 8, 9.5])
    assert p.get_color_calibration_matrix()[2, 2] == np.float32(9.5)
    with pytest.raises(ValueError):
        p.set_color_calibration_matrix([1, 2, 3])
    with pytest.raises(RawImagePipelineError) as e:
        p._set_bool("no/such/key", True)
    assert e.value.code == L.RIP_ERR_UNKNOWN_KEY
    # unknown white-balance method: std::invalid_argument with the reference's text (white_balance.hpp:81-85)
    p.set_white_balance(True)
    p.set_white_balance_method("magic")
    with pytest.raises(ValueError, match=r"White Balance method \[magic\] not supported"):
        p.process(np.zeros((8, 8), np.uint8), "bayer_rggb8")
    p.set_white_balance_method("simple")
    with pytest.raises(RawImagePipelineError) as e:
        p.process(np.zeros((8, 8), np.uint8), "bayer_rggb8")
    assert e.value.code == L.RIP_ERR_UNSUPPORTED
    p.set_white_balance(False)
    # 16-bit Bayer names listed by the reference throw (debayer.cpp:76-78, App. B-3)
    with pytest.raises(ValueError, match="is a valid pattern but is not supported"):
        p.process(np.zeros((8, 8), np.uint8), "bayer_rggb16")


def test_output_shape_rules():
    p = RawImagePipeline(False, "", "", "")
    assert p.output_shape((480, 640), "bayer_rggb8") == (480, 640, 3)
    p.set_flip(True); p.set_flip_angle(90)
    assert p.output_shape((480, 640), "bayer_rggb8") == (640, 480, 3)
    p.set_flip_angle(45)  # any other angle is a no-op even when enabled (flip.cpp:54-57)
    assert p.output_shape((480, 640), "bayer_rggb8") == (480, 640, 3)
    assert p.output_shape((480, 640, 3), "bgr8") == (480, 640, 3)


@pytest.mark.parametrize("k", [0.8, 1.0, 0.45, 2.2, 1.3])
def test_gamma_lut_matches_oracle(k):
    p = RawImagePipeline(False, "", "", "")
    p.set_gamma_correction_k(k)
    lut = np.frombuffer(p.debug_table("gamma_lut"), np.uint8)
    assert np.array_equal(lut, O.gamma_lut(k))


@pytest.mark.parametrize("shape", [(480, 640), (540, 720), (1080, 1920), (11, 13), (12, 10), (101, 64), (64, 101), (512, 512)])
@pytest.mark.parametrize("par", [(1.5, 1e-3, 1e-6), (0.7, 2e-3, 0.0), (2.0, 0.0, 1e-7)])
def test_vignetting_mask_matches_oracle(oracle_built, shape, par):
    p = RawImagePipeline(False, "", "", "")
    p.set_vignetting_correction_parameters(*par)
    rows, cols = shape
    mask = np.frombuffer(p.debug_table("vignetting_mask", rows, cols), np.float32).reshape(rows, cols)
    ref = O.vignetting_mask(rows, cols, *par)
    assert np.array_equal(mask.view(np.uint32), ref.view(np.uint32))


def test_vignetting_mask_12mp_bit_exact(oracle_built):
    p = RawImagePipeline(False, "", "", "")
    mask = np.frombuffer(p.debug_table("vignetting_mask", 3040, 4032), np.float32).reshape(3040, 4032)
    ref = O.vignetting_mask(3040, 4032, 1.5, 1e-3, 1e-6)
    assert np.array_equal(mask.view(np.uint32), ref.view(np.uint32))


def _setup_undistortion(p, calib, balance, fov, new_size=None):
    p.set_undistortion_image_size(calib["width"], calib["height"])
    if new_size:
        p.set_undistortion_new_image_size(*new_size)
    p.set_undistortion_camera_matrix(calib["K"])
    p.set_undistortion_distortion_coeffs(calib["D"])
    p.set_undistortion_distortion_model("equidistant")
    p.set_undistortion_rectification_matrix([1, 0, 0, 0, 1, 0, 0, 0, 1])
    p.set_undistortion_projection_matrix([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0])
    p.set_undistortion_balance(balance)
    p.set_undistortion_fov_scale(fov)


@pytest.mark.parametrize("size", [(720, 540), (1920, 1080), (640, 480)])
@pytest.mark.parametrize("balance,fov", [(0.0, 0.8), (0.5, 1.2), (1.0, 1.0)])
def test_fisheye_new_camera_matrix_and_maps_match_cv2(size, balance, fov):
    calib = scaled_calib(*size)
    p = RawImagePipeline(False, "", "", "")
    _setup_undistortion(p, calib, balance, fov)
    newK, mx, my = O.undistortion_maps(calib["K"], calib["D"], np.eye(3).ravel(), size, size, balance, fov)
    np.testing.assert_allclose(p.get_rect_camera_matrix(), newK, rtol=0, atol=1e-9)
    np.testing.assert_allclose(p.get_rect_projection_matrix()[:, :3], newK, rtol=0, atol=1e-9)
    assert np.array_equal(p.get_rect_distortion_coefficients(), np.zeros((1, 4)))
    m = np.frombuffer(p.debug_table("undistortion_map"), np.float32).reshape(size[1], size[0], 2)
    # what matters to cv::remap is the 1/32-pixel quantised coordinate
    assert int((np.rint(m[..., 0] * 32) != np.rint(mx * 32)).sum()) == 0
    assert int((np.rint(m[..., 1] * 32) != np.rint(my * 32)).sum()) == 0
    assert float(np.abs(m[..., 0] - mx).max()) <= 1e-3 and float(np.abs(m[..., 1] - my).max()) <= 1e-3


@pytest.mark.parametrize("size,balance,fov", [((720, 540), 0.0, 0.8), ((768, 556), 0.0, 0.45), ((768, 556), 0.5, 2.5),
                                              ((1920, 1080), 1.0, 1.2)])
def test_undistortion_tile_table_and_padded_map(size, balance, fov):
    """Host products of the tile undistortion kernel (include/rip_b200.h: debug tables): the packed entries are the
    integers cv::remap derives; the tile table's origin is the footprint's upper-left corner (x floored to 4 pixels); a
    tile carries the FAST flag exactly when every tap of every pixel lies inside the 176x48 box at that origin and no
    entry is "far"; padding entries of the tile-padded map point at the edge pixel's source position."""
    TW, TH, BW, BH = 128, 24, 176, 48
    w, h = size
    calib = scaled_calib(w, h)
    p = RawImagePipeline(False, "", "", "")
    _setup_undistortion(p, calib, balance, fov)
    m = np.frombuffer(p.debug_table("undistortion_map"), np.float32).reshape(h, w, 2)
    packed = np.frombuffer(p.debug_table("undistortion_packed_map", h, w), np.uint32).reshape(h, w)
    dx = (packed & 0xffff).astype(np.uint16).view(np.int16).astype(np.int64)
    dy = (packed >> 16).astype(np.uint16).view(np.int16).astype(np.int64)
    far = dx == -32768
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    sx = np.rint(m[..., 0].astype(np.float64) * 32).astype(np.int64); sy = np.rint(m[..., 1].astype(np.float64) * 32).astype(np.int64)
    any_tap = ((sx >> 5) >= -1) & ((sx >> 5) < w) & ((sy >> 5) >= -1) & ((sy >> 5) < h)
    assert np.array_equal(far, ~any_tap)
    assert np.array_equal((32 * xs + dx)[~far], sx[~far]) and np.array_equal((32 * ys + dy)[~far], sy[~far])

    tiles_x, tiles_y = -(-w // TW), -(-h // TH)
    table = np.frombuffer(p.debug_table("undistortion_tile_table", h, w), np.int32).reshape(tiles_y, tiles_x, 4)
    padded = np.frombuffer(p.debug_table("undistortion_tile_map", h, w), np.uint32).reshape(tiles_y * TH, tiles_x * TW)
    assert np.array_equal(padded[:h, :w], packed)
    pdx = (padded & 0xffff).astype(np.uint16).view(np.int16).astype(np.int64)
    pdy = (padded >> 16).astype(np.uint16).view(np.int16).astype(np.int64)
    pfar = pdx == -32768
    pxs, pys = np.meshgrid(np.arange(tiles_x * TW), np.arange(tiles_y * TH))
    psx, psy = 32 * pxs + pdx, 32 * pys + pdy
    # padding: same source position as the clamped (edge) pixel, or far where the edge pixel is
    cx, cy = np.minimum(pxs, w - 1), np.minimum(pys, h - 1)
    assert np.array_equal(pfar, far[cy, cx])
    assert np.array_equal(psx[~pfar], (32 * xs + dx)[cy, cx][~pfar]) and np.array_equal(psy[~pfar], (32 * ys + dy)[cy, cx][~pfar])
    n_fast = 0
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            sl = (slice(ty * TH, ty * TH + TH), slice(tx * TW, tx * TW + TW))
            ok = ~pfar[sl]
            bx0, by0, flags, _ = table[ty, tx]
            if not ok.any():
                assert (bx0, by0, flags) == (0, 0, 0)
                continue
            ix, iy = (psx[sl] >> 5)[ok], (psy[sl] >> 5)[ok]
            assert bx0 == (ix.min() & ~3) and by0 == iy.min()
            fits = ix.max() + 1 - bx0 <= BW - 1 and iy.max() + 1 - by0 <= BH - 1
            assert flags == (1 if fits and ok.all() else 0), (tx, ty)
            n_fast += int(flags)
    if fov in (0.8, 0.45):
        assert n_fast == tiles_x * tiles_y  # the example map, also zoomed in 2.2x: every tile takes the test-free path
    if fov == 2.5:
        assert n_fast == 0                  # zoomed out 2.5x: footprints overflow the box, corners map outside the source


def test_fisheye_new_size_and_1p6mp_calibration_file():
    p = RawImagePipeline(False, "", os.path.join(CONFIG, "alphasense_calib_1.6mp_example.yaml"), "")
    assert (p.get_dist_image_width(), p.get_dist_image_height()) == (1440, 1080)
    p.set_undistortion_new_image_size(720, 540)
    K = p.get_dist_camera_matrix(); D = p.get_dist_distortion_coefficients().reshape(4, 1)
    ref = cv2.fisheye.estimateNewCameraMatrixForUndistortRectify(K, D, (1440, 1080), np.eye(3), balance=0.0,
                                                                 new_size=(720, 540), fov_scale=0.8)  # default YAML
    np.testing.assert_allclose(p.get_rect_camera_matrix(), ref, rtol=0, atol=1e-9)
    assert (p.get_rect_image_width(), p.get_rect_image_height()) == (720, 540)


def test_yaml_subset_parser_handles_reference_files(tmp_path):
    y = tmp_path / "params.yaml"
    y.write_text("# comment\nflip:\n  enabled: yes   # trailing\n  angle: 180\n"
                 "gamma_correction: \n  enabled: True\n  method: 'custom'\n  k: 1.25\n"
                 "color_enhancer:\n  run_color_enhancer: true\n  value_gain: 1.4\n"
                 "undistortion:\n  enabled: false\n")
    p = RawImagePipeline(False, str(y), "", "")
    assert p.is_flip_enabled() and p._get_int("flip/angle") == 180
    assert p.is_gamma_correction_enabled() and p._get_double("gamma_correction/k") == 1.25
    assert p.is_color_enhancer_enabled() and p._get_double("color_enhancer/value_gain_member") == 1.4
    assert not p.is_undistortion_enabled() and not p.is_white_balance_enabled()  # defaults of cpp:77,148
    c = tmp_path / "color.yaml"
    c.write_text("matrix:\n  rows: 3\n  cols: 3\n  data: [1.5, 0, 0,\n         0, 1.25, 0,\n         0, 0, 2]\nbias:\n  data: [1, 2, 3]\n")
    p.load_color_calibration(str(c))
    assert np.array_equal(p.get_color_calibration_matrix(), np.diag([1.5, 1.25, 2]).astype(np.float32))
    assert np.array_equal(p.get_color_calibration_bias().ravel(), [1, 2, 3, 0])


# the methods the reference's pybind module defines (raw_image_pipeline_python/src/raw_image_pipeline_python.cpp:16-73)
PYBIND_METHODS = """
apply get_dist_camera_matrix get_dist_distortion_coefficients get_dist_distortion_model
get_dist_image_height get_dist_image_width get_dist_projection_matrix get_dist_rectification_matrix
get_rect_camera_matrix get_rect_distortion_coefficients get_rect_distortion_model get_rect_image_height
get_rect_image_width get_rect_projection_matrix get_rect_rectification_matrix load_params
process reset_white_balance_temporal_consistency set_color_calibration set_color_calibration_bias
set_color_calibration_matrix set_color_enhancer set_color_enhancer_hue_gain set_color_enhancer_saturation_gain
set_color_enhancer_value_gain set_debayer set_debayer_encoding set_debug
set_flip set_flip_angle set_gamma_correction set_gamma_correction_k
set_gamma_correction_method set_gpu set_undistortion set_undistortion_balance
set_undistortion_camera_matrix set_undistortion_distortion_coeffs set_undistortion_distortion_model set_undistortion_fov_scale
set_undistortion_image_size set_undistortion_new_image_size set_undistortion_projection_matrix set_undistortion_rectification_matrix
set_vignetting_correction set_vignetting_correction_parameters set_white_balance set_white_balance_method
set_white_balance_percentile set_white_balance_saturation_threshold set_white_balance_temporal_consistency
""".split()


def test_python_mirror_has_every_method_of_the_reference_pybind_module():
    """Row b: apply_pipeline.py-style scripts switch over by changing one import."""
    missing = [m for m in PYBIND_METHODS if not callable(getattr(RawImagePipeline, m, None))]
    assert not missing, missing
    ref = "/root/reference/raw_image_pipeline_python/src/raw_image_pipeline_python.cpp"
    if os.path.exists(ref):  # this container only: the list above is the file's
        import re
        defs = set(re.findall(r'\.def\("([a-z_0-9]+)"', open(ref).read()))
        assert defs == set(PYBIND_METHODS), defs ^ set(PYBIND_METHODS)


def test_pinned_result_pool_bookkeeping():
    """_PinnedPool (pipeline.py) with a stand-in allocator: buffers are reused, return when the last view dies, the pool
    is bounded, and a closed pool frees what comes back (no GPU involved: the allocator is faked with malloc)."""
    import ctypes, gc
    from raw_image_pipeline_b200.pipeline import _PinnedPool
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p; libc.malloc.argtypes = [ctypes.c_size_t]; libc.free.argtypes = [ctypes.c_void_p]

    class FakeLib:
        live = set()
        def rip_pinned_alloc(self, n, pp):
            p = libc.malloc(n); pp._obj.value = p; self.live.add(p); return 0
        def rip_pinned_free(self, p):
            p = p if isinstance(p, int) else p.value
            self.live.discard(p); libc.free(p); return 0

    lib = FakeLib()
    pool = _PinnedPool(lib)
    a = pool.take(1000); a[:] = 7
    view = a[10:20]
    addr = a.ctypes.data
    del a; gc.collect()
    assert len(pool._free) == 0, "a live view keeps the buffer out of the pool"
    assert int(view.sum()) == 70
    del view; gc.collect()
    assert len(pool._free) == 1
    b = pool.take(500)                       # a smaller request reuses the buffer
    assert b.ctypes.data == addr and b.size == 500
    held = [pool.take(64) for _ in range(_PinnedPool.MAX_SLOTS + 3)]
    assert sum(h is None for h in held) >= 3, "the pool is bounded; callers fall back to pageable arrays"
    n_live = len(lib.live)
    pool.close()
    del b, held; gc.collect()
    assert len(lib.live) == 0 and n_live > 0, "buffers coming back to a closed pool are freed"
