"""GPU parity of the BENCHMARKED kernel sequence at the BENCHMARKED sizes: the batch entry points keep the pre-undistortion
image in the 4-byte intermediate (strip kernel, rip_strip.cuh) and undistort with the TMA-staged tile kernel
(rip_fast.cu k_remap_tile) on the real tile table of a 12 MP / 4K map -- the pair bench.py times.  Every frame of a small
batch against the cv2 oracle, plus the alternative kernels for the same frames (global-memory gather; round-1 tile fused
kernel).  BASELINE configs 3 (64 x 4032x3040, pca) and 5 (4K, ccc + undistort)."""
import numpy as np
import pytest

from raw_image_pipeline_b200 import synth
from test_gpu_parity import FULL, assert_same, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dist", ["U", "N"])
def test_config3_batch_kernels_at_4032x3040(oracle_built, dist):
    rows, cols, n, enc = 3040, 4032, 2, "bayer_rggb8"
    frames = synth.bayer_batch(n, rows, cols, enc, 3100, dist)
    p, o = make_pair(rows, cols, **FULL)
    refs = [o.apply(frames[i], enc)[0] for i in range(n)]
    launches0 = p.kernel_launches()
    got = p.process_batch(frames, enc)  # fused kernel -> 4-byte intermediate -> tile undistortion kernel
    launches = p.kernel_launches() - launches0
    assert launches > 0 and launches % 4 == 0  # stats, lut, fused, remap per chunk of the host batch: four of OUR kernels, nothing else
    for i in range(n):
        assert_same(got[i], refs[i], f"default fused + tile remap kernels, 12 MP {dist} frame {i}")
    p._set_int("debug/fused_kernel", 2)  # strip kernel
    got = p.process_batch(frames, enc)
    for i in range(n):
        assert_same(got[i], refs[i], f"strip fused + tile remap kernels, 12 MP {dist} frame {i}")
    p._set_bool("debug/force_gather_remap", True)
    got = p.process_batch(frames, enc)
    for i in range(n):
        assert_same(got[i], refs[i], f"strip fused + gather remap kernels, 12 MP {dist} frame {i}")
    p._set_bool("debug/force_gather_remap", False)
    p._set_int("debug/fused_kernel", 1)  # round-1 tile fused kernel, same intermediate
    got = p.process_batch(frames, enc)
    for i in range(n):
        assert_same(got[i], refs[i], f"tile fused + tile remap kernels, 12 MP {dist} frame {i}")


@pytest.mark.parametrize("dist", ["U", "N"])
def test_config5_batch_kernels_at_4k_ccc(oracle_built, dist):
    rows, cols, n, enc = 2160, 3840, 2, "bayer_rggb8"
    frames = synth.bayer_batch(n, rows, cols, enc, 5100, dist)
    kw = dict(FULL); kw["wb"] = "ccc"
    p, o = make_pair(rows, cols, **kw)
    refs, uvs = [], []
    for i in range(n):
        refs.append(o.apply(frames[i], enc)[0])
        uvs.append(tuple(int(v) for v in o.ccc.uv_pos))
    got = p.process_batch(frames, enc)
    for i in range(n):
        assert_same(got[i], refs[i], f"4K ccc {dist} frame {i}")
    # the batch entry point reports the estimate of its last frame (it used to leave the constructor's (128, 128))
    assert p.ccc_uv() == uvs[-1], (p.ccc_uv(), uvs)
    if dist == "N":
        assert p.ccc_uv() != (128, 128)
    p._set_bool("debug/force_gather_remap", True)
    got = p.process_batch(frames, enc)
    for i in range(n):
        assert_same(got[i], refs[i], f"4K ccc {dist} frame {i}, gather kernel")


def test_device_batch_reports_ccc_estimate(oracle_built):
    """rip_apply_batch_device is asynchronous: the estimate is fetched when the getter is called."""
    import torch
    rows, cols, n, enc = 540, 720, 3, "bayer_bggr8"
    frames = synth.bayer_batch(n, rows, cols, enc, 5200, "N")
    p, o = make_pair(rows, cols, wb="ccc", gamma=0.8)
    for i in range(n):
        ref, _ = o.apply(frames[i], enc)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=s.cuda_stream)
    assert p.ccc_uv() == tuple(int(v) for v in o.ccc.uv_pos)
    s.synchronize()
    assert_same(d_out[n - 1].cpu().numpy(), ref, "last frame of the device batch")


def test_strip_and_tile_fused_kernels_agree_on_every_stage_set(oracle_built):
    """All 32 stage sets x {0, 180 degrees} through the strip kernel and the round-1 tile kernel (debug/fused_kernel = 1),
    BGR8 output (no undistortion) and 4-byte intermediate (with undistortion), on a shape with partial strips and a ragged
    last row segment; pca and ccc white balance (the strip kernel has separate instantiations for a G table)."""
    rows, cols, enc = 166, 400, "bayer_grbg8"
    frames = synth.bayer_batch(2, rows, cols, enc, 7700, "U")
    for stages in range(32):
        for flip in (0, 180):
            kw = dict(flip=flip)
            if stages & 1: kw["wb"] = "ccc" if (stages & 6) == 2 else "pca"
            if stages & 2: kw["cc"] = True
            if stages & 4: kw["gamma"] = 0.8
            if stages & 8: kw["vig"] = (1.5, 1e-3, 1e-6)
            if stages & 16: kw["enh"] = (1.0, 1.2, 1.0)
            if stages % 3 == 0: kw["undistort"] = (0.0, 0.8)
            p, o = make_pair(rows, cols, **kw)
            p._set_int("debug/fused_kernel", 2)
            strip = p.process_batch(frames, enc)
            p._set_int("debug/fused_kernel", 1)
            tile = p.process_batch(frames, enc)
            assert_same(strip, tile, f"stages {stages} flip {flip}: strip vs tile kernel")
            if stages in (0, 5, 21, 27, 31):
                ref, _ = o.apply(frames[1], enc)
                assert_same(strip[1], ref, f"stages {stages} flip {flip}: strip kernel vs oracle")
