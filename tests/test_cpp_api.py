"""The header-only C++ drop-in class (include/raw_image_pipeline/raw_image_pipeline.hpp) compiled with
g++ against librip_b200.so: host-side behaviour without a GPU, the full chain on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import CC_EXAMPLE, ROOT, scaled_calib

PKG = os.path.join(ROOT, "raw_image_pipeline_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "_test_cpp_api")


@pytest.fixture(scope="module")
def cpp_exe():
    src = os.path.join(ROOT, "tests", "cpp", "test_cpp_api.cpp")
    hdr = os.path.join(ROOT, "include", "raw_image_pipeline", "raw_image_pipeline.hpp")
    lib = os.path.join(PKG, "librip_b200.so")
    if not os.path.exists(EXE) or any(os.path.getmtime(f) > os.path.getmtime(EXE) for f in (src, hdr, lib)):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-DRIP_B200_FORCE_PLAIN_IMAGE", "-I", os.path.join(ROOT, "include"),
                               src, "-o", EXE, "-L", PKG, "-lrip_b200", "-Wl,-rpath," + PKG])
    return EXE


def test_cpp_class_host_side(cpp_exe):
    out = subprocess.run([cpp_exe, "host"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), (out.returncode, out.stdout, out.stderr)


@pytest.mark.gpu
def test_cpp_class_full_chain_on_gpu(cpp_exe, oracle_built, tmp_path):
    from oracle import cv2_oracle as O
    from raw_image_pipeline_b200 import synth
    rows, cols = 540, 720
    raw = synth.bayer_frame(rows, cols, "bayer_rggb8", 77, "N")
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    raw.tofile(fin)
    r = subprocess.run([cpp_exe, "run", fin, str(rows), str(cols), "bayer_rggb8", fout,
                        os.path.join(PKG, "config", "alphasense_calib_example.yaml")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    got = np.fromfile(fout, np.uint8).reshape(rows, cols, 3)
    c = scaled_calib(cols, rows)
    op = O.OracleParams(flip_enabled=True, flip_angle=180, wb_enabled=True, wb_method="pca", cc_enabled=True, cc_matrix=CC_EXAMPLE,
                        gamma_enabled=True, gamma_k=0.8, vig_enabled=True, enh_enabled=True, enh_saturation_gain=1.2,
                        und_enabled=True, und_K=c["K"], und_D=c["D"], und_width=cols, und_height=rows, und_balance=0.0,
                        und_fov_scale=0.8)
    ref, _ = O.OraclePipeline(op).apply(raw, "bayer_rggb8")
    assert int(np.count_nonzero(got != ref)) == 0


@pytest.fixture(scope="module")
def cvmat_exe():
    exe = os.path.join(ROOT, "tests", "cpp", "_test_cpp_api_cvmat")
    src = os.path.join(ROOT, "tests", "cpp", "test_cpp_api_cvmat.cpp")
    deps = [src, os.path.join(ROOT, "include", "raw_image_pipeline", "raw_image_pipeline.hpp"),
            os.path.join(ROOT, "tests", "cpp", "fake_opencv", "opencv2", "core.hpp"), os.path.join(PKG, "librip_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tests", "cpp", "fake_opencv"),
                               "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-lrip_b200", "-Wl,-rpath," + PKG])
    return exe


def test_cpp_class_cv_mat_branch_compiles_and_runs_host_side(cvmat_exe):
    """The `cv::Mat` flavour of the class (what raw_image_pipeline_ros.cpp would see) against a stand-in opencv2/core.hpp."""
    out = subprocess.run([cvmat_exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), (out.returncode, out.stdout, out.stderr)


@pytest.mark.gpu
def test_cpp_class_cv_mat_branch_on_gpu(cvmat_exe):
    out = subprocess.run([cvmat_exe, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), (out.returncode, out.stdout, out.stderr)


@pytest.fixture(scope="module")
def ros_compat_exe():
    exe = os.path.join(ROOT, "tests", "cpp", "_test_ros_caller_compat")
    src = os.path.join(ROOT, "tests", "cpp", "test_ros_caller_compat.cpp")
    deps = [src, os.path.join(ROOT, "include", "raw_image_pipeline", "raw_image_pipeline.hpp"),
            os.path.join(ROOT, "tests", "cpp", "fake_opencv", "opencv2", "core.hpp"), os.path.join(PKG, "librip_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "cpp", "fake_opencv"),
                               "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-lrip_b200", "-Wl,-rpath," + PKG])
    return exe


def test_every_call_of_the_reference_ros_wrapper_compiles_and_runs_host_side(ros_compat_exe):
    """the calls raw_image_pipeline_ros.cpp:52-291 makes (parameter setters, calibration loaders, CameraInfo
    getters, image getters), with its argument types, against the drop-in header."""
    for args in (["host", os.path.join(PKG, "config")], ["host"]):
        out = subprocess.run([ros_compat_exe] + args, capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip().endswith("OK"), (args, out.returncode, out.stdout, out.stderr)


@pytest.mark.gpu
def test_reference_ros_wrapper_call_sequence_on_gpu(ros_compat_exe):
    out = subprocess.run([ros_compat_exe, "gpu", os.path.join(PKG, "config")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), (out.returncode, out.stdout, out.stderr)
