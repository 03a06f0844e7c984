"""Randomised differential test on the GPU: seeded random shapes (incl. the smallest legal ones and widths that
qualify for the TMA fast path), CFA patterns, rotations, stage subsets and parameters -- CUDA path vs the cv2 oracle,
single frames and small batches, both kernel families."""
import os

import numpy as np
import pytest

from oracle import cv2_oracle as O
from raw_image_pipeline_b200 import synth
from test_gpu_parity import CC_EXAMPLE, ENCODINGS, assert_same, make_pair

pytestmark = pytest.mark.gpu


def random_case(rng):
    fast_w = [16, 32, 48, 112, 128, 144, 256, 272, 400]
    cols = int(rng.choice(fast_w)) if rng.random() < 0.6 else int(rng.integers(3, 300))
    rows = int(rng.integers(3, 75)) if rng.random() < 0.7 else int(rng.integers(75, 260))
    kw = {}
    if rng.random() < 0.6:
        kw["flip"] = int(rng.choice([90, 180, 270]))
    if rng.random() < 0.7:
        kw["wb"] = "pca" if rng.random() < 0.7 else "ccc"
    if rng.random() < 0.7:
        kw["cc"] = True
        m = np.array(CC_EXAMPLE) * rng.uniform(0.6, 1.3, 9)
        kw["cc_matrix"] = [float(x) for x in m]
        kw["cc_bias"] = tuple(float(x) for x in rng.uniform(-8, 8, 3)) if rng.random() < 0.5 else (0.0, 0.0, 0.0)
    if rng.random() < 0.7:
        kw["gamma"] = float(rng.uniform(0.3, 2.5))
    if rng.random() < 0.6:
        kw["vig"] = (float(rng.uniform(-0.5, 2.0)), float(rng.uniform(0, 3e-3)), float(rng.uniform(0, 3e-6)))
    if rng.random() < 0.6:
        kw["enh"] = tuple(float(x) for x in rng.uniform(0.5, 1.6, 3))
    if rng.random() < 0.5:
        kw["undistort"] = (float(rng.uniform(0, 1)), float(rng.uniform(0.6, 1.4)))
    return rows, cols, str(rng.choice(ENCODINGS)), str(rng.choice(["U", "N"])), kw


@pytest.mark.parametrize("seed", range(int(os.environ.get("RIP_FUZZ_N", "100"))))
def test_random_configuration(oracle_built, seed):
    rng = np.random.default_rng(9000 + seed)
    rows, cols, enc, dist, kw = random_case(rng)
    raw = synth.bayer_frame(rows, cols, enc, 7000 + seed, dist)
    p, o = make_pair(rows, cols, **kw)
    ref, _ = o.apply(raw, enc)
    what = f"seed {seed}: {rows}x{cols} {enc} {dist} {kw}"
    assert_same(p.process(raw, enc), ref, what)
    if kw.get("wb") != "ccc":  # per-frame state-free: a batch of copies must give the same frame (batch kernels / packed map)
        out = p.process_batch(np.stack([raw, raw, raw]), enc)
        assert_same(out[2], ref, what + " [host batch]")
    p._set_bool("debug/force_generic_kernels", True)
    assert_same(p.process(raw, enc), ref, what + " [generic kernels]")
