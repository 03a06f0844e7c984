"""GPU parity tests proper: the CUDA path (through the C ABI, via the Python mirror of the
reference's pybind class) against the cv2 oracle on the same seeded inputs.  Bit-exact is the
bar (+-1 LSB end to end needs every stage exact)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import CC_EXAMPLE, ROOT, scaled_calib
from oracle import cv2_oracle as O
from raw_image_pipeline_b200 import RawImagePipeline, synth

pytestmark = pytest.mark.gpu

ENCODINGS = ["bayer_rggb8", "bayer_grbg8", "bayer_gbrg8", "bayer_bggr8"]


def make_pair(rows, cols, *, flip=0, wb=None, cc=False, gamma=None, vig=None, enh=None, undistort=None,
              cc_matrix=CC_EXAMPLE, cc_bias=(0.0, 0.0, 0.0)):
    """Build a (RawImagePipeline, OraclePipeline) pair with identical settings."""
    p = RawImagePipeline(False, "", "", "")
    op = O.OracleParams()
    for name in ("white_balance", "color_calibration", "gamma_correction", "vignetting_correction", "color_enhancer",
                 "undistortion", "flip"):
        getattr(p, "set_" + name)(False)
    if flip:
        p.set_flip(True); p.set_flip_angle(flip)
        op.flip_enabled = True; op.flip_angle = flip
    if wb:
        p.set_white_balance(True); p.set_white_balance_method(wb)
        p.set_white_balance_saturation_threshold(0.8, 0.2); p.set_white_balance_temporal_consistency(False)
        op.wb_enabled = True; op.wb_method = wb; op.wb_bright_thr = 0.8; op.wb_dark_thr = 0.2
        op.wb_temporal_consistency = False
    if cc:
        # calibration_available_ was set by the constructor's default colour-calibration load
        p.set_color_calibration(True); p.set_color_calibration_matrix(cc_matrix); p.set_color_calibration_bias(cc_bias)
        op.cc_enabled = True; op.cc_matrix = list(cc_matrix); op.cc_bias = list(cc_bias)
    if gamma is not None:
        p.set_gamma_correction(True); p.set_gamma_correction_method("custom"); p.set_gamma_correction_k(gamma)
        op.gamma_enabled = True; op.gamma_k = gamma
    if vig is not None:
        p.set_vignetting_correction(True); p.set_vignetting_correction_parameters(*vig)
        op.vig_enabled = True; op.vig_scale, op.vig_a2, op.vig_a4 = vig
    if enh is not None:  # (hue_gain_, saturation_gain_, value_gain_) *member* values
        p.set_color_enhancer(True)
        p.set_color_enhancer_value_gain(enh[0])       # cross-wired: -> hue_gain_
        p.set_color_enhancer_saturation_gain(enh[1])
        p.set_color_enhancer_hue_gain(enh[2])         # cross-wired: -> value_gain_
        op.enh_enabled = True; op.enh_hue_gain, op.enh_saturation_gain, op.enh_value_gain = enh
    if undistort is not None:
        balance, fov = undistort
        frows, fcols = (cols, rows) if flip in (90, 270) else (rows, cols)
        calib = scaled_calib(fcols, frows)
        # calibration_available_ is only set by loadCalibration in the reference: load, then override
        p.load_camera_calibration(os.path.join(ROOT, "raw_image_pipeline_b200", "config", "alphasense_calib_example.yaml"))
        p.set_undistortion_image_size(fcols, frows)
        p.set_undistortion_camera_matrix(calib["K"]); p.set_undistortion_distortion_coeffs(calib["D"])
        p.set_undistortion_balance(balance); p.set_undistortion_fov_scale(fov)
        p.set_undistortion(True)
        op.und_enabled = True; op.und_K = calib["K"]; op.und_D = calib["D"]; op.und_width = fcols; op.und_height = frows
        op.und_balance = balance; op.und_fov_scale = fov
    return p, O.OraclePipeline(op, os.path.join(ROOT, "raw_image_pipeline_b200", "config", "ccc_model.bin"))


def assert_same(got, ref, what=""):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    diff = got.astype(np.int16) - ref.astype(np.int16)
    nbad = int(np.count_nonzero(diff))
    assert nbad == 0, f"{what}: {nbad} differing values, max |diff| = {int(np.abs(diff).max())}"


FULL = dict(flip=180, wb="pca", cc=True, gamma=0.8, vig=(1.5, 1e-3, 1e-6), enh=(1.0, 1.2, 1.0), undistort=(0.0, 0.8))


# ---- config 1: 640x480 rggb8, debayer + gamma --------------------------------------------------
def test_config1_debayer_gamma(oracle_built):
    raw = synth.bayer_frame(480, 640, "bayer_rggb8", 1000, "U")
    p, o = make_pair(480, 640, gamma=0.8)
    ref, enc = o.apply(raw, "bayer_rggb8")
    assert_same(p.process(raw, "bayer_rggb8"), ref, "config1")
    assert enc == "bgr8"
    assert p.kernel_launches() >= 1


@pytest.mark.parametrize("enc", ENCODINGS)
@pytest.mark.parametrize("shape", [(480, 640), (11, 13), (10, 12), (3, 3), (4, 9), (33, 130), (67, 259), (540, 722)])
def test_debayer_all_patterns_and_ragged_sizes(enc, shape):
    raw = synth.bayer_frame(shape[0], shape[1], enc, 7, "U")
    p, o = make_pair(*shape)
    ref, _ = o.apply(raw, enc)
    assert_same(p.process(raw, enc), ref, f"debayer {enc} {shape}")


@pytest.mark.parametrize("angle", [90, 180, 270])
@pytest.mark.parametrize("shape", [(480, 640), (33, 130), (67, 259)])
def test_flip(angle, shape):
    raw = synth.bayer_frame(shape[0], shape[1], "bayer_grbg8", 11, "U")
    p, o = make_pair(*shape, flip=angle)
    ref, _ = o.apply(raw, "bayer_grbg8")
    assert_same(p.process(raw, "bayer_grbg8"), ref, f"flip {angle} {shape}")


# ---- single stages on top of debayer -----------------------------------------------------------
@pytest.mark.parametrize("dist", ["U", "N"])
@pytest.mark.parametrize("kw", [
    dict(wb="pca"), dict(cc=True), dict(cc=True, cc_bias=(3.25, -7.5, 0.49)), dict(gamma=0.45), dict(gamma=2.2),
    dict(vig=(1.5, 1e-3, 1e-6)), dict(vig=(0.7, 2e-3, 0.0)), dict(enh=(1.0, 1.2, 1.0)), dict(enh=(1.1, 0.8, 1.3)),
    dict(undistort=(0.0, 0.8)), dict(undistort=(1.0, 1.2)), dict(flip=180, vig=(1.5, 1e-3, 1e-6)),
    dict(flip=90, vig=(1.5, 1e-3, 1e-6), undistort=(0.0, 0.8)),
], ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items()))
def test_single_stage(oracle_built, kw, dist):
    rows, cols = 540, 720
    raw = synth.bayer_frame(rows, cols, "bayer_bggr8", 21, dist)
    p, o = make_pair(rows, cols, **kw)
    ref, _ = o.apply(raw, "bayer_bggr8")
    assert_same(p.process(raw, "bayer_bggr8"), ref, str(kw))


# ---- config 2: 1920x1080 bggr8 full chain ---------------------------------------------------------
@pytest.mark.parametrize("dist", ["U", "N"])
def test_config2_full_chain_1080p(oracle_built, dist):
    rows, cols = 1080, 1920
    raw = synth.bayer_frame(rows, cols, "bayer_bggr8", 2000, dist)
    p, o = make_pair(rows, cols, **FULL)
    ref, _ = o.apply(raw, "bayer_bggr8", keep_stages=True)
    got = p.process(raw, "bayer_bggr8")
    assert_same(p.get_dist_debayered_image(), o.stages["flip"], "debayered+flipped image")
    assert_same(p.get_dist_color_image(), o.stages["color_enhancer"], "pre-undistortion colour image")
    assert_same(got, ref, "config2 rect image")
    assert_same(p.get_processed_image(), ref, "processed image")
    assert p.get_rect_mask().size == 0
    # PCA coefficients agree with the oracle's fp32 solve bit for bit
    (cb, cr), _ = O.pca_coefficients(o.stages["flip"])
    coeff = p._get_doubles("stats/pca_coefficients")
    assert [np.float32(c) for c in coeff] == [cb[0], cb[1], cr[0], cr[1]]


@pytest.mark.parametrize("enc", ENCODINGS)
@pytest.mark.parametrize("flip", [0, 180])
def test_fast_and_generic_kernels_agree(oracle_built, enc, flip):
    """Shapes that qualify for the TMA fast path (width % 16 == 0) also run through the generic kernels
    (debug/force_generic_kernels): both must equal the oracle, incl. frame borders and partial edge tiles."""
    rows, cols = 70, 208   # partial tiles in both directions, width % 32 != 0 (cv2 row tail inside the fast path)
    raw = synth.bayer_frame(rows, cols, enc, 99, "U")
    kw = dict(FULL); kw["flip"] = flip; kw.pop("undistort")
    p, o = make_pair(rows, cols, **kw)
    ref, _ = o.apply(raw, enc)
    assert_same(p.process(raw, enc), ref, f"fast path {enc} flip {flip}")
    p._set_bool("debug/force_generic_kernels", True)
    assert_same(p.process(raw, enc), ref, f"generic path {enc} flip {flip}")


def test_generic_kernels_full_chain_1080p(oracle_built):
    rows, cols = 1080, 1920
    raw = synth.bayer_frame(rows, cols, "bayer_bggr8", 2001, "N")
    p, o = make_pair(rows, cols, **FULL)
    p._set_bool("debug/force_generic_kernels", True)
    ref, _ = o.apply(raw, "bayer_bggr8")
    assert_same(p.process(raw, "bayer_bggr8"), ref, "generic kernels, config 2")


# ---- config 3: 4032x3040 full chain (one frame against the oracle, batch by property) ------------
def test_config3_full_chain_12mp_one_frame(oracle_built):
    rows, cols = 3040, 4032
    raw = synth.bayer_frame(rows, cols, "bayer_rggb8", 3000, "N")
    p, o = make_pair(rows, cols, **FULL)
    ref, _ = o.apply(raw, "bayer_rggb8")
    assert_same(p.process(raw, "bayer_rggb8"), ref, "config3 12MP")


def test_batch_device_equals_per_frame_calls(oracle_built):
    """Size-independent property: the batched device entry point gives, frame by frame, exactly
    what single-frame apply() gives (per-frame WB statistics stay per frame)."""
    import torch
    rows, cols, n = 1080, 1920, 6
    frames = synth.bayer_batch(n, rows, cols, "bayer_bggr8", 4000, "N")
    frames[1] = synth.bayer_frame(rows, cols, "bayer_bggr8", 4001, "U")
    p, o = make_pair(rows, cols, **FULL)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, "bayer_bggr8", d_out.data_ptr(), host=False,
                        stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    host = p.process_batch(frames, "bayer_bggr8")
    for i in range(n):
        single = p.process(frames[i], "bayer_bggr8")
        assert_same(got[i], single, f"device batch frame {i}")
        assert_same(host[i], single, f"host batch frame {i}")
    ref, _ = o.apply(frames[1], "bayer_bggr8")
    assert_same(got[1], ref, "batch frame 1 vs oracle")


def test_undistortion_batch_paths_packed_and_float_map(oracle_built):
    """Batch entry points keep the pre-undistortion image in the 4-byte intermediate and read the packed
    fixed-point map; with debug/force_float_map the same gather reads the fp32 map. Both equal the oracle."""
    rows, cols, n = 540, 720, 3
    frames = synth.bayer_batch(n, rows, cols, "bayer_rggb8", 4100, "U")
    for balance, fov in ((0.0, 0.8), (1.0, 1.2)):
        kw = dict(FULL); kw["undistort"] = (balance, fov)
        p, o = make_pair(rows, cols, **kw)
        packed = p.process_batch(frames, "bayer_rggb8")
        p._set_bool("debug/force_float_map", True)
        floatmap = p.process_batch(frames, "bayer_rggb8")
        for i in range(n):
            ref, _ = o.apply(frames[i], "bayer_rggb8")
            assert_same(packed[i], ref, f"packed map frame {i} {balance} {fov}")
            assert_same(floatmap[i], ref, f"float map frame {i} {balance} {fov}")


@pytest.mark.parametrize("balance,fov", [(0.0, 0.8), (1.0, 1.2), (0.0, 0.45), (0.5, 2.5)])
def test_undistortion_tile_kernel_equals_gather_kernel(oracle_built, balance, fov):
    """The TMA-staged tile kernel (source boxes in shared memory) and the global-memory gather give the oracle's bytes:
    mild and zoomed-in maps (every tile takes the test-free path), zoomed-out maps whose footprints overflow the box
    (per-pixel fallback) with pixels mapping outside the source ("far" entries, zero border), and a width/height that
    leave partial tiles."""
    rows, cols, n = 540 + 16, 720 + 48, 2
    frames = synth.bayer_batch(n, rows, cols, "bayer_gbrg8", 4200, "N")
    kw = dict(FULL); kw["undistort"] = (balance, fov)
    p, o = make_pair(rows, cols, **kw)
    tile = p.process_batch(frames, "bayer_gbrg8")
    p._set_bool("debug/force_gather_remap", True)
    gather = p.process_batch(frames, "bayer_gbrg8")
    for i in range(n):
        ref, _ = o.apply(frames[i], "bayer_gbrg8")
        assert_same(tile[i], ref, f"tile kernel frame {i} {balance} {fov}")
        assert_same(gather[i], ref, f"gather kernel frame {i} {balance} {fov}")


# ---- 3-channel inputs (apply_pipeline.py usage: a bgr8 PNG) ---------------------------------------
@pytest.mark.parametrize("enc", ["bgr8", "rgb8"])
def test_colour_input(oracle_built, enc):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (270, 362, 3), dtype=np.uint8)
    kw = dict(FULL); kw.pop("undistort")
    p, o = make_pair(270, 362, **kw)
    ref, renc = o.apply(img, enc)
    got = p.process(img, enc)
    assert_same(got, ref, enc)
    work = img.copy()
    assert p.apply(work, enc) is True  # same shape -> in place, like the reference binding
    assert_same(work, ref, enc + " in place")


def test_exhaustive_colour_cube_through_the_kernel(oracle_built):
    """All 2^24 BGR triples through CC -> gamma -> vignetting -> enhancer on the GPU."""
    v = np.arange(256, dtype=np.uint8)
    a, b, c = np.meshgrid(v, v, v, indexing="ij")
    cube = np.ascontiguousarray(np.stack([a, b, c], -1).reshape(4096, 4096, 3))
    p, o = make_pair(4096, 4096, cc=True, gamma=0.8, vig=(1.5, 1e-3, 1e-6), enh=(1.0, 1.2, 1.0))
    ref, _ = o.apply(cube, "bgr8")
    assert_same(p.process(cube, "bgr8"), ref, "2^24 cube")


def test_batch_device_with_caller_dist_color_buffer(oracle_built):
    """rip_apply_batch_device's optional `d_dist_color` receives the pre-undistortion BGR8 frames
    (getDistColorImage of every frame of the batch)."""
    import torch
    rows, cols, n = 270, 368, 3
    frames = synth.bayer_batch(n, rows, cols, "bayer_grbg8", 4200, "N")
    p, o = make_pair(rows, cols, **FULL)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device="cuda")
    d_col = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, "bayer_grbg8", d_out.data_ptr(), host=False,
                        dist_color_ptr=d_col.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for i in range(n):
        ref, _ = o.apply(frames[i], "bayer_grbg8", keep_stages=True)
        assert_same(d_col[i].cpu().numpy(), o.stages["color_enhancer"], f"dist colour frame {i}")
        assert_same(d_out[i].cpu().numpy(), ref, f"rect frame {i}")


@pytest.mark.parametrize("flip", [0, 90, 180])
def test_mono_input_passes_through_the_colour_modules(oracle_built, flip):
    """mono8 (1-channel, not Bayer): debayer leaves it alone, white balance / colour calibration / enhancer skip images
    without 3 channels, flip + gamma LUT + undistortion still apply (SURVEY 8b 'Input kinds')."""
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (270, 362), dtype=np.uint8)
    p, o = make_pair(270, 362, flip=flip, wb="pca", cc=True, gamma=0.8, enh=(1.0, 1.2, 1.0), undistort=(0.0, 0.8))
    ref, enc = o.apply(img, "mono8", keep_stages=True)
    got = p.process(img, "mono8")
    assert enc == "mono8" and got.ndim == 2
    assert_same(got, ref, f"mono8 flip {flip}")
    assert_same(p.get_dist_debayered_image(), o.stages["flip"], "mono debayered+flipped")
    assert_same(p.get_dist_color_image(), o.stages["color_enhancer"], "mono pre-undistortion")
    p.set_vignetting_correction(True)
    with pytest.raises(ValueError):
        p.process(img, "mono8")


def test_two_gpus_from_one_process(oracle_built):
    """One pipeline instance per GPU in the same process (rip_set_device): both produce the oracle's bytes,
    concurrently from two host threads (one host worker per GPU, no collective)."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rows, cols = 540, 720
    frames = [synth.bayer_frame(rows, cols, "bayer_rggb8", 4300 + i, "N") for i in range(2)]
    _, o = make_pair(rows, cols, **FULL)
    refs = [o.apply(f, "bayer_rggb8")[0] for f in frames]
    got = [None, None]
    errors = []

    def worker(dev):
        try:
            p, _ = make_pair(rows, cols, **FULL)
            p._check(p._lib.rip_set_device(p._h, dev))
            for _ in range(3):
                got[dev] = p.process(frames[dev], "bayer_rggb8")
            batch = p.process_batch(np.stack([frames[dev]] * 3), "bayer_rggb8")   # 4-byte intermediate + tile undistortion
            assert_same(batch[2], refs[dev], f"batch on device {dev}")
        except BaseException as e:  # an exception in a thread would otherwise be lost
            errors.append((dev, e))

    threads = [threading.Thread(target=worker, args=(d,)) for d in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for d in range(2):
        assert got[d] is not None
        assert_same(got[d], refs[d], f"device {d}")


@pytest.mark.parametrize("flip", [0, 180])
@pytest.mark.parametrize("rows", [33, 65, 97])
def test_fast_path_when_the_last_frame_row_opens_a_tile(oracle_built, rows, flip):
    """rows % 32 == 1: the border rule for the frame's last (first, when rotated) row reaches beyond the tile halo."""
    cols = 48
    raw = synth.bayer_frame(rows, cols, "bayer_gbrg8", 58, "N")
    kw = dict(FULL); kw["flip"] = flip; kw.pop("undistort")
    p, o = make_pair(rows, cols, **kw)
    ref, _ = o.apply(raw, "bayer_gbrg8")
    assert_same(p.process(raw, "bayer_gbrg8"), ref, f"rows {rows} flip {flip}")


@pytest.mark.parametrize("pad_in,pad_out", [(48, 32), (7, 5)])
def test_batch_device_with_padded_frame_strides(oracle_built, pad_in, pad_out):
    """Frames `in_frame_stride` / `out_frame_stride` bytes apart with gaps: 16-byte multiples stay on the TMA fast path,
    odd gaps fall back to the generic kernels; the gaps are never written."""
    import torch
    rows, cols, n = 64, 96, 3
    frames = synth.bayer_batch(n, rows, cols, "bayer_rggb8", 4400, "U")
    kw = dict(FULL); kw.pop("undistort")
    p, o = make_pair(rows, cols, **kw)
    in_stride, out_stride = rows * cols + pad_in, rows * cols * 3 + pad_out
    h_in = np.full(n * in_stride, 0xAB, np.uint8)
    for i in range(n):
        h_in[i * in_stride:i * in_stride + rows * cols] = frames[i].ravel()
    d_in = torch.from_numpy(h_in).cuda()
    d_out = torch.full((n * out_stride,), 0xCD, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, "bayer_rggb8", d_out.data_ptr(), host=False,
                        stream=torch.cuda.current_stream().cuda_stream, in_frame_stride=in_stride, out_frame_stride=out_stride)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for i in range(n):
        ref, _ = o.apply(frames[i], "bayer_rggb8")
        assert_same(got[i * out_stride:i * out_stride + rows * cols * 3].reshape(rows, cols, 3), ref, f"frame {i}")
        assert (got[i * out_stride + rows * cols * 3:(i + 1) * out_stride] == 0xCD).all()


def test_strided_input_rows_and_noop_flip_angles(oracle_built):
    """`step` larger than a row (a cv::Mat ROI / numpy view of a wider buffer), and flip angles the reference ignores
    (flip.cpp:37-58: anything but 90 / 180 / 270 leaves the image alone even when the module is enabled)."""
    rows, cols = 70, 112
    wide = synth.bayer_frame(rows, cols + 40, "bayer_grbg8", 4500, "U")
    view = wide[:, 8:8 + cols]
    assert not view.flags.c_contiguous
    p, o = make_pair(rows, cols, gamma=0.8, enh=(1.0, 1.2, 1.0))
    ref, _ = o.apply(np.ascontiguousarray(view), "bayer_grbg8")
    assert_same(p.process(view, "bayer_grbg8"), ref, "strided rows")
    for angle in (0, 45, 360, -90):
        p.set_flip(True); p.set_flip_angle(angle)
        assert_same(p.process(np.ascontiguousarray(view), "bayer_grbg8"), ref, f"flip angle {angle} is a no-op")
