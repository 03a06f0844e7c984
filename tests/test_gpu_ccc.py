"""GPU parity of the convolutional-colour-constancy white balance (SURVEY 8a row a5 + the Kalman
tracker of 8f-1) against the cv2 replay of convolutional_color_constancy.cpp."""
import numpy as np
import pytest

from oracle import cv2_oracle as O
from raw_image_pipeline_b200 import synth
from test_gpu_parity import FULL, assert_same, make_pair

pytestmark = pytest.mark.gpu

CFA_RGGB = {(0, 0): 2, (0, 1): 1, (1, 0): 1, (1, 1): 0}  # (row%2, col%2) -> channel index in (B, G, R)


def flat_bayer(rows, cols, bgr, noise=0, seed=0):
    """A Bayer (rggb) frame whose demosaiced interior is the constant colour `bgr`."""
    raw = np.empty((rows, cols), np.uint8)
    for (py, px), c in CFA_RGGB.items():
        raw[py::2, px::2] = bgr[c]
    if noise:
        rng = np.random.default_rng(seed)
        raw = np.clip(raw.astype(np.int16) + rng.integers(-noise, noise + 1, raw.shape), 0, 255).astype(np.uint8)
    return raw


def check_frame(p, o, raw, enc, what):
    ref, _ = o.apply(raw, enc)
    got = p.process(raw, enc)
    uv = (p._get_int("stats/ccc_u"), p._get_int("stats/ccc_v"))
    assert uv == tuple(int(x) for x in o.ccc.uv_pos), (what, uv, o.ccc.uv_pos)
    gains = [np.float32(g) for g in p._get_doubles("stats/ccc_gains")]
    assert gains == [np.float32(g) for g in o.ccc.last_gains], (what, gains, o.ccc.last_gains)
    assert_same(got, ref, what)


@pytest.mark.parametrize("dist", ["U", "N"])
@pytest.mark.parametrize("shape", [(540, 720), (1080, 1920), (271, 361), (135, 180)])
def test_ccc_white_balance_only(oracle_built, shape, dist):
    raw = synth.bayer_frame(shape[0], shape[1], "bayer_bggr8", 31, dist)
    p, o = make_pair(*shape, wb="ccc")
    check_frame(p, o, raw, "bayer_bggr8", f"ccc {shape} {dist}")


@pytest.mark.parametrize("angle", [90, 180, 270])
def test_ccc_after_flip(oracle_built, angle):
    raw = synth.bayer_frame(270, 362, "bayer_grbg8", 33, "N")
    p, o = make_pair(270, 362, wb="ccc", flip=angle)
    check_frame(p, o, raw, "bayer_grbg8", f"ccc flip {angle}")


def test_ccc_response_map_and_argmax_on_peaked_histograms(oracle_built):
    """Flat-colour frames concentrate the histogram in one bin, so the learned filter (not only the
    bias) decides the arg-max: checks the FFT path against cv2's fp32 response."""
    p, o = make_pair(540, 720, wb="ccc")
    seen = set()
    for i, bgr in enumerate([(204, 204, 204), (60, 120, 200), (200, 120, 60), (90, 200, 90), (180, 70, 180), (120, 119, 118),
                             (52, 230, 140), (201, 64, 77)]):
        raw = flat_bayer(540, 720, bgr, noise=2, seed=i)
        check_frame(p, o, raw, "bayer_rggb8", f"flat {bgr}")
        seen.add(tuple(o.ccc.uv_pos))
        resp = np.frombuffer(p.debug_table("ccc_response"), np.float64).reshape(256, 256) + o.ccc.bias.astype(np.float64)
        ref = o.ccc.response_map.astype(np.float64) / 65536.0
        assert np.abs(resp - ref).max() <= 2e-6 * np.abs(ref).max()
    assert len(seen) >= 3, seen  # the test really moves the illuminant estimate


def test_ccc_colour_input(oracle_built):
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    for enc in ("bgr8", "rgb8"):
        p, o = make_pair(300, 400, wb="ccc", gamma=0.8)
        check_frame(p, o, img, enc, enc)


def sequence():
    colours = [(204, 204, 204), (60, 120, 200), (60, 120, 200), (200, 120, 60), (90, 200, 90), (90, 200, 90), (90, 200, 90),
               (180, 70, 180), (204, 204, 204), (52, 230, 140), (52, 230, 140), (201, 64, 77), (201, 64, 77), (201, 64, 77)]
    return [flat_bayer(270, 360, c, noise=3, seed=100 + i) for i, c in enumerate(colours)]


def test_ccc_temporal_consistency_and_reset(oracle_built):
    """cv::KalmanFilter tracker (ccc.cpp:300-340): frame-by-frame equality, a reset in the middle,
    and switching the tracker off and on again."""
    p, o = make_pair(270, 360, wb="ccc", gamma=0.8)
    p.set_white_balance_temporal_consistency(True); o.p.wb_temporal_consistency = True
    frames = sequence()
    for i, raw in enumerate(frames):
        if i == 6:
            p.reset_white_balance_temporal_consistency(); o.ccc.first_frame = True
        if i == 9:
            p.set_white_balance_temporal_consistency(False); o.p.wb_temporal_consistency = False
        if i == 11:
            p.set_white_balance_temporal_consistency(True); o.p.wb_temporal_consistency = True
        check_frame(p, o, raw, "bayer_rggb8", f"temporal frame {i}")


def test_ccc_temporal_batch_equals_frame_by_frame(oracle_built):
    import torch
    frames = np.stack(sequence())
    n, rows, cols = frames.shape
    p, o = make_pair(rows, cols, wb="ccc", gamma=0.8)
    p.set_white_balance_temporal_consistency(True); o.p.wb_temporal_consistency = True
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, "bayer_rggb8", d_out.data_ptr(), host=False,
                        stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for i in range(n):
        ref, _ = o.apply(frames[i], "bayer_rggb8")
        assert_same(got[i], ref, f"temporal batch frame {i}")


def test_config5_4k_full_chain_ccc(oracle_built):
    """BASELINE configs[4], one frame: 3840x2160 rggb8, every module on, ccc white balance, undistortion."""
    rows, cols = 2160, 3840
    raw = synth.bayer_frame(rows, cols, "bayer_rggb8", 5000, "N")
    kw = dict(FULL); kw["wb"] = "ccc"
    p, o = make_pair(rows, cols, **kw)
    check_frame(p, o, raw, "bayer_rggb8", "config5 4K ccc")


def test_ccc_temporal_host_batch_spanning_several_chunks(oracle_built):
    """rip_apply_batch_host splits long batches into chunks on several CUDA streams; the Kalman recurrence must still
    see the frames in order."""
    frames = np.stack(sequence() * 3)   # 42 frames -> three chunks of <= 16
    n, rows, cols = frames.shape
    p, o = make_pair(rows, cols, wb="ccc", gamma=0.8)
    p.set_white_balance_temporal_consistency(True); o.p.wb_temporal_consistency = True
    got = p.process_batch(frames, "bayer_rggb8")
    for i in range(n):
        ref, _ = o.apply(frames[i], "bayer_rggb8")
        assert_same(got[i], ref, f"temporal host batch frame {i}")
