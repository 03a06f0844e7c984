"""GPU tests of the single-frame entry point the reference's callers use (rip_apply == RawImagePipeline::apply): pinned
staging + CUDA-graph replay, the lazily recomputed image getters, the rect-mask extension, the multi-GPU front end."""
import numpy as np
import pytest

from raw_image_pipeline_b200 import MultiGpuPipeline, synth
from test_gpu_parity import FULL, assert_same, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wb", ["pca", "ccc"])
def test_apply_replays_a_cuda_graph_and_stays_exact(oracle_built, wb):
    """Five different frames through one pipeline: from the third on the chain runs as a graph replay (pca; ccc keeps
    host-side state per frame and stays un-captured); every result, and the getters, equal the oracle."""
    rows, cols, enc = 1080, 1920, "bayer_bggr8"
    kw = dict(FULL); kw["wb"] = wb
    p, o = make_pair(rows, cols, **kw)
    for i in range(5):
        raw = synth.bayer_frame(rows, cols, enc, 2100 + i, "N" if i % 2 else "U")
        ref, _ = o.apply(raw, enc, keep_stages=True)
        assert_same(p.process(raw, enc), ref, f"frame {i}")
        if i in (0, 3):
            assert_same(p.get_dist_color_image(), o.stages["color_enhancer"], f"frame {i}: pre-undistortion colour image")
            assert_same(p.get_dist_debayered_image(), o.stages["flip"], f"frame {i}: debayered image")
            assert_same(p.get_processed_image(), ref, f"frame {i}: processed image")
    replays = p._get_int("stats/graph_replays")
    assert replays >= 3 if wb == "pca" else replays == 0
    # a setter invalidates the captured graph: the next frames are right under the new configuration
    p.set_gamma_correction_k(0.6)
    o.p.gamma_k = 0.6
    for i in range(3):
        raw = synth.bayer_frame(rows, cols, enc, 2200 + i, "N")
        ref, _ = o.apply(raw, enc)
        assert_same(p.process(raw, enc), ref, f"after set_gamma_correction_k, frame {i}")
    # switching the graph off gives the same bytes
    p._set_bool("apply/cuda_graph", False)
    assert_same(p.process(raw, enc), ref, "graph disabled")


def test_apply_with_changing_shapes(oracle_built):
    """Alternating frame sizes (and a strided view) through one pipeline: buffers are re-sized, the graph re-captured."""
    kw = dict(FULL); kw.pop("undistort")
    p, o = make_pair(64, 96, **kw)
    for i, (rows, cols) in enumerate([(64, 96), (64, 96), (64, 96), (128, 160), (128, 160), (128, 160), (64, 96), (70, 112)]):
        raw = synth.bayer_frame(rows, cols, "bayer_rggb8", 2300 + i, "U")
        ref, _ = o.apply(raw, "bayer_rggb8")
        assert_same(p.process(raw, "bayer_rggb8"), ref, f"{rows}x{cols} call {i}")


def test_rect_mask_extension(oracle_built):
    """getRectMask(): empty like the reference's (undistortion.hpp:136, never written) unless "undistortion/rect_mask" is
    set; then u8, 255 exactly where all four taps of cv::remap's bilinear interpolation lie inside the source image."""
    rows, cols, enc = 540, 720, "bayer_rggb8"
    raw = synth.bayer_frame(rows, cols, enc, 2400, "N")
    for balance, fov in ((0.0, 0.8), (1.0, 1.2)):
        kw = dict(FULL); kw["undistort"] = (balance, fov)
        p, o = make_pair(rows, cols, **kw)
        p.process(raw, enc)
        assert p.get_rect_mask().size == 0
        p._set_bool("undistortion/rect_mask", True)
        p.process(raw, enc)
        mask = p.get_rect_mask()
        _, mx, my = o.maps()
        sx, sy = np.rint(mx.astype(np.float32) * np.float32(32)).astype(np.int64), np.rint(my.astype(np.float32) * np.float32(32)).astype(np.int64)
        ix, iy = sx >> 5, sy >> 5
        want = np.where((ix >= 0) & (ix + 1 < cols) & (iy >= 0) & (iy + 1 < rows), 255, 0).astype(np.uint8)
        assert mask.shape == want.shape and mask.dtype == np.uint8
        assert int(np.count_nonzero(mask != want)) == 0
        if (balance, fov) == (1.0, 1.2):
            assert 0 < int((want == 0).sum()) < want.size  # this map does leave the source image
        # where the mask is set, a constant-colour source gives exactly that colour (no border contribution)
    no_undistort = dict(FULL); no_undistort.pop("undistort")
    p, _ = make_pair(rows, cols, **no_undistort)
    p._set_bool("undistortion/rect_mask", True)
    p.process(raw, enc)
    assert p.get_rect_mask().size == 0  # no undistortion, no mask


def test_multi_gpu_pipeline_shards_a_batch(oracle_built):
    """MultiGpuPipeline over every GPU of the box (1 on the driver's test box, 2+ under gpurun --gpus N): one call,
    contiguous chunks per GPU, every frame equal to the oracle; camera streams stay on their GPU."""
    import torch
    n_gpus = torch.cuda.device_count()
    rows, cols, enc, n = 270, 368, "bayer_grbg8", 7
    frames = synth.bayer_batch(n, rows, cols, enc, 2500, "N")
    mp = MultiGpuPipeline(n_gpus=n_gpus, use_gpu=False, params_path="")
    p1, o = make_pair(rows, cols, **FULL)   # configure the replicas like the single pipeline of make_pair
    for name in ("white_balance", "color_calibration", "gamma_correction", "vignetting_correction", "color_enhancer", "undistortion", "flip"):
        getattr(mp, "set_" + name)(False)
    mp.set_flip(True); mp.set_flip_angle(180)
    mp.set_white_balance(True); mp.set_white_balance_method("pca")
    from conftest import CC_EXAMPLE
    mp.set_color_calibration(True); mp.set_color_calibration_matrix(CC_EXAMPLE); mp.set_color_calibration_bias((0.0, 0.0, 0.0))
    mp.set_gamma_correction(True); mp.set_gamma_correction_method("custom"); mp.set_gamma_correction_k(0.8)
    mp.set_vignetting_correction(True); mp.set_vignetting_correction_parameters(1.5, 1e-3, 1e-6)
    mp.set_color_enhancer(True); mp.set_color_enhancer_value_gain(1.0); mp.set_color_enhancer_saturation_gain(1.2); mp.set_color_enhancer_hue_gain(1.0)
    assert mp.n_gpus == n_gpus and sum(e - b for b, e in mp.shard(n)) == n
    no_und = dict(FULL); no_und.pop("undistort")
    _, o2 = make_pair(rows, cols, **no_und)
    out = mp.process_batch(frames, enc)
    for i in range(n):
        ref, _ = o2.apply(frames[i], enc)
        assert_same(out[i], ref, f"multi-GPU batch frame {i}")
    streams = [frames[:3], frames[3:5], frames[5:]]
    outs = mp.process_streams(streams, enc)
    k = 0
    for s_i, s_out in enumerate(outs):
        for j in range(len(s_out)):
            ref, _ = o2.apply(frames[k], enc)
            assert_same(s_out[j], ref, f"stream {s_i} frame {j}")
            k += 1
    assert mp.kernel_launches() > 0


@pytest.mark.parametrize("enc", ["bayer_rggb16", "bayer_bggr16"])
@pytest.mark.parametrize("flip", [0, 90, 180])
def test_16bit_bayer_extension(oracle_built, enc, flip):
    """bayer_*16 (uint16 frames): rejected like the reference does unless set_debayer_allow_16bit(True); then demosaiced at
    16 bits, reduced to 8 (oracle/cv2_oracle.py debayer16) and run through the full chain -- single frame and batch."""
    rows, cols = 270, 368
    rng = np.random.default_rng(77)
    frames = (rng.integers(0, 4096, (3, rows, cols), dtype=np.uint16) << 4)   # 12-bit sensor data, MSB-aligned
    frames[1] = rng.integers(0, 65536, (rows, cols), dtype=np.uint16)
    kw = dict(FULL); kw["flip"] = flip
    p, o = make_pair(rows, cols, **kw)
    if enc != "bayer_bggr16":   # (the reference's name list has a typo for bggr16: that one falls through as an unknown encoding)
        with pytest.raises(ValueError, match="valid pattern but is not supported"):
            p.process(frames[0], enc)
    p.set_debayer_allow_16bit(True)
    o.p.debayer_allow_16bit = True
    refs = [o.apply(frames[i], enc)[0] for i in range(3)]
    for i in range(3):
        assert_same(p.process(frames[i], enc), refs[i], f"{enc} flip {flip} frame {i}")
    batch = p.process_batch(frames, enc)
    for i in range(3):
        assert_same(batch[i], refs[i], f"{enc} flip {flip} batch frame {i}")
    assert p.output_shape((rows, cols), enc)[2] == 3


def test_pinned_and_pageable_buffers(oracle_built):
    """rip_apply with page-locked caller buffers (copy engines address them directly) and with pageable ones (staged through
    the pipeline's pinned buffers) gives the same bytes; results of process() come from the pinned pool and go back to it."""
    rows, cols = 480, 640
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (6, rows, cols), dtype=np.uint8)
    p, o = make_pair(rows, cols, **FULL)
    refs = [o.apply(f, "bayer_bggr8")[0] for f in frames]
    pin_in = p.pinned_empty((rows, cols))
    assert pin_in is not None
    held = []
    for mode in ("pageable-out", "pinned-out", "pinned-in", "registered"):
        p.use_pinned_results = mode not in ("pageable-out", "registered")
        p._set_bool("apply/register_caller_buffers", mode == "registered")  # opt-in: the caller's ordinary buffers get page-locked once
        for i, f in enumerate(frames):
            src = f
            if mode == "pinned-in":
                pin_in[...] = f
                src = pin_in
            out = p.process(src, "bayer_bggr8")
            assert_same(out, refs[i], f"{mode} frame {i}")
            held.append(out)   # results stay valid while the caller holds them, whatever later calls do
    for k, out in enumerate(held):
        assert_same(out, refs[k % len(frames)], f"held result {k}")
    p._set_bool("apply/register_caller_buffers", False)
    p.use_pinned_results = True
    # the pool has a bound: holding more results than it may pin falls back to pageable arrays, transparently
    many = [p.process(frames[0], "bayer_bggr8") for _ in range(p._pool.MAX_SLOTS + 4)]
    for out in many:
        assert_same(out, refs[0], "beyond the pool")
    import gc
    del many
    gc.collect()
    assert len(p._pool._free) >= 1
    del p          # the pipeline goes first: buffers it registered are still mapped when it unregisters them
    gc.collect()
    del held
