"""The two oracles pin each other: the from-scratch numpy restatement of SURVEY Appendix A
(oracle/numpy_chain.py, no cv2 in the pixel path) against the cv2 call-for-call replay
(oracle/cv2_oracle.py), stage by stage on seeded frames."""
import cv2
import numpy as np
import pytest

from conftest import CC_EXAMPLE, scaled_calib
from oracle import cv2_oracle as O
from oracle import numpy_chain as N
from raw_image_pipeline_b200 import synth


@pytest.mark.parametrize("enc", sorted(N.CFA))
@pytest.mark.parametrize("shape", [(48, 64), (11, 13), (10, 12), (3, 3)])
def test_debayer(enc, shape):
    raw = synth.bayer_frame(shape[0], shape[1], enc, 3, "U")
    ref, _ = O.debayer(raw, enc)
    assert np.array_equal(N.debayer(raw, enc), ref)


@pytest.mark.parametrize("angle", [0, 90, 180, 270])
def test_flip(angle):
    img = np.random.default_rng(1).integers(0, 256, (13, 17, 3), dtype=np.uint8)
    assert np.array_equal(N.flip(img, angle), O.flip(img, angle))


@pytest.mark.parametrize("dist", ["U", "N"])
def test_full_chain_stage_by_stage(oracle_built, dist):
    rows, cols = 132, 200   # cols % 32 != 0: cv2's scalar row tail in HSV2BGR is exercised
    raw = synth.bayer_frame(rows, cols, "bayer_bggr8", 17, dist)
    c = scaled_calib(cols, rows)
    o = O.OraclePipeline(O.OracleParams(flip_enabled=True, flip_angle=180, wb_enabled=True, wb_method="pca", cc_enabled=True,
                                        cc_matrix=CC_EXAMPLE, cc_bias=[1.5, -2.25, 0.0], gamma_enabled=True, gamma_k=0.8,
                                        vig_enabled=True, enh_enabled=True, enh_hue_gain=1.0, enh_saturation_gain=1.2,
                                        enh_value_gain=1.1, und_enabled=True, und_K=c["K"], und_D=c["D"], und_width=cols,
                                        und_height=rows, und_balance=0.0, und_fov_scale=0.8))
    o.apply(raw, "bayer_bggr8", keep_stages=True)
    st = o.stages
    img = N.debayer(raw, "bayer_bggr8")
    assert np.array_equal(img, st["debayer"])
    img = N.flip(img, 180)
    assert np.array_equal(img, st["flip"])
    img = N.white_balance_pca(img)
    assert np.array_equal(img, st["white_balance"])
    img = N.color_calibration(img, CC_EXAMPLE, [1.5, -2.25, 0.0])
    assert np.array_equal(img, st["color_calibration"])
    img = N.gamma(img, 0.8)
    assert np.array_equal(img, st["gamma"])
    mask = N.vignetting_mask(rows, cols, 1.5, 1e-3, 1e-6)
    assert np.array_equal(mask, O.vignetting_mask(rows, cols, 1.5, 1e-3, 1e-6))  # numpy pow == this box's libm here
    img = N.vignetting(img, mask)
    assert np.array_equal(img, st["vignetting"])
    img = N.color_enhancer(img, 1.0, 1.2, 1.1)
    assert np.array_equal(img, st["color_enhancer"])
    _, mx, my = o.maps()
    img = N.remap(img, mx, my)
    assert np.array_equal(img, st["undistortion"])


def test_remap_random_maps_incl_out_of_bounds():
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    mx = rng.uniform(-3, 53, (60, 70)).astype(np.float32)
    my = rng.uniform(-3, 43, (60, 70)).astype(np.float32)
    mx[0, :3] = [np.inf, -np.inf, 1e9]
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    assert np.array_equal(N.remap(src, mx, my), ref)
