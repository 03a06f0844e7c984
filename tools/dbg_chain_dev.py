#!/usr/bin/env python3
"""GPU debugging aid: chain_quad vs chain_pixel on the device (tests/gpu_chain/_chain_dev.so)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CC_EXAMPLE, cube
from oracle import cv2_oracle as O
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "gpu_chain", "_chain_dev.so"))
P = ctypes.c_void_p
img = cube()[:256]
n = img.shape[0] * img.shape[1]
rng = np.random.default_rng(3)
for stages in [8, 12, 16, 24, 31]:
    for use_mask in (False, True):
        mask = (rng.uniform(1.0, 2.6, n).astype(np.float32) if use_mask else np.ones(n, np.float32)) if stages & 8 else None
        cc = np.asarray(CC_EXAMPLE, np.float32); enh = np.asarray([1.0, 1.2, 1.0], np.float64)
        wb = np.tile(np.arange(256, dtype=np.uint8), 3); gamma = O.gamma_lut(0.8)
        q = np.empty_like(img); r = np.empty_like(img)
        rc = lib.chain_dev_run(ctypes.c_uint(stages), ctypes.c_long(n), P(img.ctypes.data), P(mask.ctypes.data) if mask is not None else None,
                               P(cc.ctypes.data), P(enh.ctypes.data), P(wb.ctypes.data), P(gamma.ctypes.data), P(q.ctypes.data), P(r.ctypes.data))
        d = (q != r).any(-1)
        print(f"stages {stages} mask {use_mask}: rc {rc}, differing pixels {int(d.sum())} of {n}")
        if d.any():
            idx = np.argwhere(d)[:6]
            for (y, x) in idx:
                print("   in", img[y, x], "mask", None if mask is None else mask[y * img.shape[1] + x], "quad", q[y, x], "ref", r[y, x])
