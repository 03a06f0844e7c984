"""Builds librip_b200.so (the C-ABI product library) in-tree with nvcc for sm_100a.

    python -m raw_image_pipeline_b200.build        # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The library links the CUDA runtime statically and depends
on nothing else (no torch, no OpenCV).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librip_b200.so")
SOURCES = ["rip_api.cu", "rip_kernels.cu", "rip_fast.cu", "rip_strip.cu", "rip_strip_bgr8.cu", "ccc.cu", "host_state.cpp"]
HEADERS = ["pixel_math.cuh", "frame_math.cuh", "chain_tables.hpp", "kernels.hpp", "host_state.hpp", "ccc.hpp", "ccc_math.cuh", "devbuf.hpp", "tma.cuh", "bayer_window.cuh", "chain_quad.cuh", "rip_strip.cuh",
           "cv_tables.inc", os.path.join("..", "..", "include", "rip_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # every FMA in the pixel math is explicit (pixel_math.cuh)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "--cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(HERE, "build", s + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "static",
           "-Xcompiler", "-fPIC", "-o", LIB, *objs, "-ldl", "-lpthread"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
