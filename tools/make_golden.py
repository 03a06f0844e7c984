#!/usr/bin/env python3
"""Generates tests/golden/chain_golden.npz: seeded Bayer inputs and the outputs of the reference's CPU
path for them, produced by the cv2 call-for-call replay (oracle/cv2_oracle.py, cv2 version recorded).

The reference ships no golden vectors for this path; these pin the oracle
against drift of the OpenCV build and let the GPU tests run against committed bytes.

    python tools/make_golden.py        (needs cv2 + `make -C oracle`)

When /root/reference is present the script also records what the oracle's CCC estimator returns on
the reference's two sample images (raw_image_pipeline_white_balance/data/*.png).
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CC_EXAMPLE, scaled_calib  # noqa: E402
from oracle import cv2_oracle as O  # noqa: E402
from raw_image_pipeline_b200 import synth  # noqa: E402

MODEL = os.path.join(ROOT, "raw_image_pipeline_b200", "config", "ccc_model.bin")
ROWS, COLS = 66, 100   # ragged on purpose: cols % 32 != 0 exercises cv2's scalar row tails

# name -> (encoding, distribution, seed, OracleParams overrides)
CASES = {
    "debayer_rggb": ("bayer_rggb8", "U", 1, {}),
    "debayer_grbg": ("bayer_grbg8", "U", 2, {}),
    "debayer_gbrg": ("bayer_gbrg8", "U", 3, {}),
    "debayer_bggr": ("bayer_bggr8", "U", 4, {}),
    "config1_debayer_gamma": ("bayer_rggb8", "U", 5, dict(gamma_enabled=True, gamma_k=0.8)),
    "flip90_vignetting": ("bayer_grbg8", "N", 6, dict(flip_enabled=True, flip_angle=90, vig_enabled=True)),
    "flip270_enhancer": ("bayer_gbrg8", "N", 7, dict(flip_enabled=True, flip_angle=270, enh_enabled=True, enh_hue_gain=1.1,
                                                     enh_saturation_gain=0.8, enh_value_gain=1.3)),
    "pca_cc_bias": ("bayer_bggr8", "N", 8, dict(wb_enabled=True, wb_method="pca", cc_enabled=True, cc_matrix=CC_EXAMPLE,
                                                 cc_bias=[3.25, -7.5, 0.49])),
    "full_chain_pca_U": ("bayer_bggr8", "U", 9, "FULL_PCA"),
    "full_chain_pca_N": ("bayer_rggb8", "N", 10, "FULL_PCA"),
    "full_chain_ccc_N": ("bayer_rggb8", "N", 11, "FULL_CCC"),
}


def full(wb):
    c = scaled_calib(COLS, ROWS)
    return dict(flip_enabled=True, flip_angle=180, wb_enabled=True, wb_method=wb, wb_bright_thr=0.8, wb_dark_thr=0.2,
                wb_temporal_consistency=False, cc_enabled=True, cc_matrix=CC_EXAMPLE, gamma_enabled=True, gamma_k=0.8,
                vig_enabled=True, enh_enabled=True, enh_saturation_gain=1.2, und_enabled=True, und_K=c["K"], und_D=c["D"],
                und_width=COLS, und_height=ROWS, und_balance=0.0, und_fov_scale=0.8)


def case_params(spec):
    if spec == "FULL_PCA":
        return full("pca")
    if spec == "FULL_CCC":
        return full("ccc")
    return spec


def main():
    out = {}
    meta = {"cv2": cv2.__version__, "rows": ROWS, "cols": COLS, "cases": {}}
    for name, (enc, dist, seed, spec) in CASES.items():
        kw = case_params(spec)
        raw = synth.bayer_frame(ROWS, COLS, enc, seed, dist)
        o = O.OraclePipeline(O.OracleParams(**kw), MODEL)
        res, oenc = o.apply(raw, enc)
        out[name + "__in"] = raw
        out[name + "__out"] = res
        meta["cases"][name] = {"encoding": enc, "dist": dist, "seed": seed, "out_encoding": oenc,
                               "params": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in kw.items()}}
    ref_data = "/root/reference/raw_image_pipeline_white_balance/data"
    if os.path.isdir(ref_data):
        meta["reference_sample_images"] = {}
        for fn in sorted(os.listdir(ref_data)):
            img = cv2.imread(os.path.join(ref_data, fn), cv2.IMREAD_COLOR)
            if img is None:
                continue
            ccc = O.CCC(MODEL)
            ccc.balance_white(img)
            meta["reference_sample_images"][fn] = {"shape": list(img.shape), "uv": [int(ccc.uv_pos[0]), int(ccc.uv_pos[1])],
                                                   "gains_bgr": [float(g) for g in ccc.last_gains]}
    dst = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "chain_golden.npz"), **out)
    with open(os.path.join(dst, "chain_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(CASES), "cases;", os.path.getsize(os.path.join(dst, "chain_golden.npz")), "bytes")
    print(json.dumps(meta.get("reference_sample_images", {}), indent=1))


if __name__ == "__main__":
    main()
