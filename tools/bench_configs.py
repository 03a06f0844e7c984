#!/usr/bin/env python3
"""Secondary measurements for the BASELINE.json configs that are not the bench.py headline (one GPU):
config 1 (640x480 debayer+gamma), config 2 / 4 (1920x1080 full chain, per-frame latency and stream throughput),
config 5 (3840x2160, ccc white balance + undistortion, 64 frames per GPU).  Prints one JSON object per config.

    python tools/bench_configs.py [--frames 64] [--iters 20]
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from raw_image_pipeline_b200 import RawImagePipeline, synth  # noqa: E402


def full_pipeline(rows, cols, wb="pca"):
    p = bench.make_pipeline(rows, cols, device=0)
    p.set_white_balance_method(wb)
    p.set_white_balance_saturation_threshold(0.8, 0.2)
    p.set_white_balance_temporal_consistency(False)
    return p


def device_throughput(p, frames, enc, iters):
    n, rows, cols = frames.shape
    d_in = torch.from_numpy(frames).cuda()
    orows, ocols, och = p.output_shape((rows, cols), enc)
    d_out = torch.empty((n, orows, ocols, och), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    run = lambda: p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=stream)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    p._set_bool("profile/kernel_events", True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    k = p._get_doubles("stats/kernel_ms")
    p._set_bool("profile/kernel_events", False)
    ms = e0.elapsed_time(e1) / iters
    return {"frames": n, "ms_per_batch": ms, "us_per_frame": ms * 1e3 / n, "mpix_per_s": n * rows * cols / (ms * 1e-3) / 1e6,
            "kernel_ms_per_batch": {name: k[i] / iters for i, name in enumerate(["stats", "lut", "fused", "remap"])}}


def latency(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e6)
    ts.sort()
    return {"p50_us": statistics.median(ts), "p99_us": ts[min(len(ts) - 1, int(0.99 * len(ts)))], "min_us": ts[0], "n": len(ts)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    out = []

    # config 1: 640x480 rggb8, debayer + gamma
    p1 = bench.make_witness_pipeline(device=0)
    f1 = synth.bayer_batch(min(a.frames, 16), 480, 640, "bayer_rggb8", 1000, "U")
    f1 = np.concatenate([f1] * ((a.frames + len(f1) - 1) // len(f1)))[:a.frames]
    r = {"config": "1: 640x480 bayer_rggb8, debayer + gamma", "device_batch": device_throughput(p1, f1, "bayer_rggb8", a.iters)}
    one = np.ascontiguousarray(f1[0])
    r["rip_apply_pageable_host"] = latency(lambda: p1.process(one, "bayer_rggb8"), 200)
    out.append(r)

    # config 2 / 4: 1920x1080 bggr8 full chain (pca), per frame and as a stream
    rows, cols, enc = 1080, 1920, "bayer_bggr8"
    p2 = full_pipeline(rows, cols)
    f2 = synth.bayer_batch(8, rows, cols, enc, 2000, "N")
    f2 = np.concatenate([f2] * ((a.frames + 7) // 8))[:a.frames]
    r = {"config": "2/4: 1920x1080 bayer_bggr8, full chain (pca WB, undistortion), one stream on one GPU",
         "device_batch": device_throughput(p2, f2, enc, a.iters)}
    one = np.ascontiguousarray(f2[0])
    r["rip_apply_pageable_host"] = latency(lambda: p2.process(one, enc), 200)
    h_in = torch.from_numpy(f2).pin_memory()
    h_out = torch.empty((a.frames, rows, cols, 3), dtype=torch.uint8).pin_memory()
    r["batch_host_pinned_1_frame"] = latency(
        lambda: p2.process_batch_ptr(h_in.data_ptr(), 1, rows, cols, 1, enc, h_out.data_ptr(), host=True), 256)
    p2.process_batch_ptr(h_in.data_ptr(), a.frames, rows, cols, 1, enc, h_out.data_ptr(), host=True)  # warm-up: allocations
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        p2.process_batch_ptr(h_in.data_ptr(), a.frames, rows, cols, 1, enc, h_out.data_ptr(), host=True)
    dt = time.perf_counter() - t0
    r["batch_host_pinned_stream"] = {"frames": a.frames * reps, "fps": a.frames * reps / dt,
                                     "mpix_per_s": a.frames * reps * rows * cols / dt / 1e6}
    out.append(r)

    # config 5 (one GPU's share): 64 x 3840x2160 rggb8, all modules, ccc WB, undistortion
    rows, cols, enc = 2160, 3840, "bayer_rggb8"
    p5 = full_pipeline(rows, cols, wb="ccc")
    f5 = synth.bayer_batch(8, rows, cols, enc, 5000, "N")
    f5 = np.concatenate([f5] * ((a.frames + 7) // 8))[:a.frames]
    r = {"config": "5 (per GPU): 64 x 3840x2160 bayer_rggb8, full chain, ccc WB (0.8/0.2, no temporal consistency), undistortion",
         "device_batch": device_throughput(p5, f5, enc, max(3, a.iters // 4))}
    r["ccc_last_uv"] = [p5._get_int("stats/ccc_u"), p5._get_int("stats/ccc_v")]
    out.append(r)

    for r in out:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
