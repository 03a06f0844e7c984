"""Committed golden vectors (tests/golden, made by tools/make_golden.py from the cv2 replay of the
reference CPU path): the oracle must keep reproducing them (CPU), and the CUDA path must match
them byte for byte through the C ABI (GPU) -- no oracle involved in the GPU comparison."""
import json
import os

import cv2
import numpy as np
import pytest

from conftest import ROOT
from oracle import cv2_oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
META = json.load(open(os.path.join(GOLD, "chain_golden.json")))
DATA = np.load(os.path.join(GOLD, "chain_golden.npz"))
MODEL = os.path.join(ROOT, "raw_image_pipeline_b200", "config", "ccc_model.bin")
CASES = sorted(META["cases"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle_built, name):
    if cv2.__version__ != META["cv2"]:
        pytest.skip(f"golden vectors were made with cv2 {META['cv2']}")
    c = META["cases"][name]
    o = O.OraclePipeline(O.OracleParams(**c["params"]), MODEL)
    out, enc = o.apply(DATA[name + "__in"], c["encoding"])
    assert enc == c["out_encoding"]
    assert np.array_equal(out, DATA[name + "__out"])


def test_oracle_ccc_on_the_reference_sample_images():
    """Known answers recorded in SURVEY.md 8c: uv = (111, 139), gains (b, g, r) = (2.117, 1, 1.367)."""
    d = "/root/reference/raw_image_pipeline_white_balance/data"
    if not os.path.isdir(d):
        pytest.skip("reference checkout not present on this machine")
    for fn, want in META["reference_sample_images"].items():
        ccc = O.CCC(MODEL)
        ccc.balance_white(cv2.imread(os.path.join(d, fn), cv2.IMREAD_COLOR))
        assert [int(ccc.uv_pos[0]), int(ccc.uv_pos[1])] == want["uv"] == [111, 139]
        np.testing.assert_allclose([float(g) for g in ccc.last_gains], [2.117, 1.0, 1.367], atol=5e-4)


def configure(p, q, cols, rows):
    """Apply a golden case's OracleParams dict through the reference's setter API."""
    cfg = os.path.join(ROOT, "raw_image_pipeline_b200", "config")
    for name in ("white_balance", "color_calibration", "gamma_correction", "vignetting_correction", "color_enhancer",
                 "undistortion", "flip"):
        getattr(p, "set_" + name)(False)
    if q.get("flip_enabled"):
        p.set_flip(True); p.set_flip_angle(q["flip_angle"])
    if q.get("wb_enabled"):
        p.set_white_balance(True); p.set_white_balance_method(q["wb_method"])
        p.set_white_balance_saturation_threshold(q.get("wb_bright_thr", 0.8), q.get("wb_dark_thr", 0.1))
        p.set_white_balance_temporal_consistency(q.get("wb_temporal_consistency", True))
    if q.get("cc_enabled"):
        p.set_color_calibration(True); p.set_color_calibration_matrix(q["cc_matrix"])
        p.set_color_calibration_bias(q.get("cc_bias", [0, 0, 0]))
    if q.get("gamma_enabled"):
        p.set_gamma_correction(True); p.set_gamma_correction_method("custom"); p.set_gamma_correction_k(q["gamma_k"])
    if q.get("vig_enabled"):
        p.set_vignetting_correction(True)
        p.set_vignetting_correction_parameters(q.get("vig_scale", 1.5), q.get("vig_a2", 1e-3), q.get("vig_a4", 1e-6))
    if q.get("enh_enabled"):
        p.set_color_enhancer(True)
        p.set_color_enhancer_value_gain(q.get("enh_hue_gain", 1.0))   # cross-wired setters (color_enhancer.cpp:23-33)
        p.set_color_enhancer_saturation_gain(q.get("enh_saturation_gain", 1.0))
        p.set_color_enhancer_hue_gain(q.get("enh_value_gain", 1.0))
    if q.get("und_enabled"):
        p.load_camera_calibration(os.path.join(cfg, "alphasense_calib_example.yaml"))
        p.set_undistortion_image_size(q["und_width"], q["und_height"])
        p.set_undistortion_camera_matrix(q["und_K"]); p.set_undistortion_distortion_coeffs(q["und_D"])
        p.set_undistortion_balance(q["und_balance"]); p.set_undistortion_fov_scale(q["und_fov_scale"])
        p.set_undistortion(True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_golden(name):
    from raw_image_pipeline_b200 import RawImagePipeline
    c = META["cases"][name]
    p = RawImagePipeline(False, "", "", "")
    configure(p, c["params"], META["cols"], META["rows"])
    got = p.process(DATA[name + "__in"], c["encoding"])
    want = DATA[name + "__out"]
    assert got.shape == want.shape
    assert int(np.count_nonzero(got != want)) == 0
