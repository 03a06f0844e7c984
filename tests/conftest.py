import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no librip_b200.so (built artefacts are git-ignored): build it once (nvcc cross-compiles without
    a GPU).  The package itself never builds or falls back on its own -- it raises when the library is missing."""
    lib = os.path.join(ROOT, "raw_image_pipeline_b200", "librip_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call([sys.executable, "-m", "raw_image_pipeline_b200.build"], cwd=ROOT)


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_built():
    """Make sure oracle/_build/libvignetting_mask.so exists (checker infrastructure)."""
    so = os.path.join(ROOT, "oracle", "_build", "libvignetting_mask.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return so


@pytest.fixture(scope="session")
def hostsim():
    """Host build of the kernels' per-pixel arithmetic (test infrastructure only)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "_hostsim.so")
    srcs = [os.path.join(d, "hostsim.cpp")] + [
        os.path.join(ROOT, "raw_image_pipeline_b200", "csrc", f)
        for f in ("pixel_math.cuh", "frame_math.cuh", "cv_tables.inc", "chain_tables.hpp", "ccc_math.cuh", "chain_quad.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-x", "c++", srcs[0], "-o", so])
    lib = ctypes.CDLL(so)
    return lib


def cube():
    """All 2^24 u8 triples as a 4096x4096x3 image."""
    v = np.arange(256, dtype=np.uint8)
    a, b, c = np.meshgrid(v, v, v, indexing="ij")
    return np.ascontiguousarray(np.stack([a, b, c], -1).reshape(4096, 4096, 3))


CC_EXAMPLE = [2.4276948, 0.21479778, -0.30818, 0.09277014, 1.1962607, -0.09772757,
              -0.24436986, -0.22239459, 2.099912]
CALIB_720 = dict(
    K=[347.548139773951, 0.0, 342.454373227748, 0.0, 347.434712422309, 271.368057185649, 0.0, 0.0, 1.0],
    D=[-0.0396482888762527, -0.00367688950406141, 0.00391742438164282, -0.00178738156007817],
    width=720, height=540)


def scaled_calib(width, height):
    """The 720x540 example calibration scaled to width x height."""
    sx, sy = width / 720.0, height / 540.0
    K = CALIB_720["K"]
    return dict(K=[K[0] * sx, 0.0, K[2] * sx, 0.0, K[4] * sy, K[5] * sy, 0.0, 0.0, 1.0],
                D=list(CALIB_720["D"]), width=width, height=height)
