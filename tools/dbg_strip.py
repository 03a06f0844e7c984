#!/usr/bin/env python3
"""GPU debugging aid: strip kernel vs round-1 tile kernel per stage set; prints where they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from raw_image_pipeline_b200 import synth
from test_gpu_parity import make_pair

def describe(a, b, what):
    d = a != b
    n = int(d.sum())
    if n == 0:
        print(f"  OK   {what}"); return
    idx = np.argwhere(d)
    ys, xs, ch = idx[:, 1], idx[:, 2], idx[:, 3]
    print(f"  DIFF {what}: {n} values; rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()} per-channel {[int((ch == c).sum()) for c in range(3)]}"
          f" first {idx[0].tolist()} got {a[tuple(idx[0])]} want {b[tuple(idx[0])]}; distinct rows {len(set(ys.tolist()))} distinct cols {len(set(xs.tolist()))}")

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 166
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 400
enc = "bayer_grbg8"
frames = synth.bayer_batch(2, rows, cols, enc, 7700, "U")
for stages in [0, 4, 1, 2, 3, 8, 16, 12, 24, 31]:
    for flip in (0, 180):
        for und in (False, True):
            kw = dict(flip=flip)
            if stages & 1: kw["wb"] = "pca"
            if stages & 2: kw["cc"] = True
            if stages & 4: kw["gamma"] = 0.8
            if stages & 8: kw["vig"] = (1.5, 1e-3, 1e-6)
            if stages & 16: kw["enh"] = (1.0, 1.2, 1.0)
            if und: kw["undistort"] = (0.0, 0.8)
            p, o = make_pair(rows, cols, **kw)
            p._set_int("debug/fused_kernel", 2)
            strip = p.process_batch(frames, enc)
            p._set_int("debug/fused_kernel", 1)
            tile = p.process_batch(frames, enc)
            describe(strip, tile, f"stages {stages:2d} flip {flip:3d} und {und}: strip vs tile")
