#!/usr/bin/env python3
"""Fused-kernel time per stage set, tile kernel against strip kernel (debug/fused_kernel = 1 / 2), 64 x 4032x3040 frames
device-resident, with and without undistortion (4-byte intermediate / BGR8 output): the measurement behind
strip_kernel_preferred().  Every combination is also checked against the oracle on one small frame."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from test_gpu_parity import make_pair

rows, cols, enc, n = 3040, 4032, "bayer_rggb8", 64
dev = torch.device("cuda:0")
frames = bench.make_frames(16, rows, cols, 77, enc=enc)
d_in = torch.from_numpy(np.concatenate([frames] * 4)).to(dev)
d_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
SETS = {"none": {}, "gamma": dict(gamma=0.8), "wb": dict(wb="pca"), "wb+gamma": dict(wb="pca", gamma=0.8), "cc": dict(cc=True),
        "cc+gamma": dict(cc=True, gamma=0.8), "wb+cc": dict(wb="pca", cc=True), "wb+cc+gamma": dict(wb="pca", cc=True, gamma=0.8),
        "wb+cc+gamma+vig": dict(wb="pca", cc=True, gamma=0.8, vig=(1.5, 1e-3, 1e-6)), "full": dict(wb="pca", cc=True, gamma=0.8, vig=(1.5, 1e-3, 1e-6), enh=(1.0, 1.2, 1.0))}
out = []
for name, kw in SETS.items():
    for und in (False, True):
        kw2 = dict(kw)
        if und: kw2["undistort"] = (0.0, 0.8)
        row = {"stage_set": name, "undistortion": und}
        for fk, label in ((1, "tile_ms"), (2, "strip_ms")):
            p, _ = make_pair(rows, cols, **kw2)
            p._set_int("debug/fused_kernel", fk)
            for _ in range(3):
                p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=stream)
            torch.cuda.synchronize()
            p._set_bool("profile/kernel_events", True)
            for _ in range(5):
                p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=stream)
            torch.cuda.synchronize()
            k = p._get_doubles("stats/kernel_ms")
            row[label] = round(k[2] / max(1.0, k[6]), 4)
            del p
        out.append(row); print(json.dumps(row), flush=True)
