#!/usr/bin/env python3
"""Where the time of one RawImagePipeline::apply() call goes (1080p full chain): copy-in / launch / device wait / copy-out
as measured inside rip_apply ("stats/apply_us"), and the Python-visible latency, for a few copy-pool sizes."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import bench
    from raw_image_pipeline_b200 import synth
    rows, cols, enc = 1080, 1920, "bayer_bggr8"
    p = bench.make_pipeline(rows, cols)
    if os.environ.get("NOGRAPH"): p._set_bool("apply/cuda_graph", False)
    frames = synth.bayer_batch(8, rows, cols, enc, 1, "N")
    if os.environ.get("PAGEABLE_OUT"): p.use_pinned_results = False   # results in fresh pageable arrays (staged copy-out)
    if os.environ.get("PINNED_IN"):                                     # the caller's frames are page-locked too
        pinned = [p.pinned_empty((rows, cols)) for _ in range(8)]
        for a, f in zip(pinned, frames): a[...] = f
        frames = pinned
    for i in range(10): p.process(frames[i % 8], enc)
    lat, parts = [], []
    for i in range(200):
        t0 = time.perf_counter(); p.process(frames[i % 8], enc); lat.append((time.perf_counter() - t0) * 1e6)
        parts.append(p._get_doubles("stats/apply_us"))
    lat = np.sort(lat); parts = np.median(np.asarray(parts), axis=0)
    print(json.dumps({"copy_threads": os.environ.get("RIP_B200_COPY_THREADS"), "p50_us": float(lat[100]), "p99_us": float(lat[198]),
                      "serial_in": os.environ.get("RIP_B200_SERIAL_COPY_IN"), "graph": os.environ.get("NOGRAPH") is None,
                      "pinned_in": bool(os.environ.get("PINNED_IN")), "pageable_out": bool(os.environ.get("PAGEABLE_OUT")),
                      "inside_rip_apply_us": {"copy_in": parts[0], "launch": parts[1], "device_a": parts[2], "device_b": parts[3], "copy_out": parts[4]}}))
else:
    # device_a + device_b = waiting for H2D + kernels + D2H (the read-back of the PCA coefficients into pageable memory
    # blocks first, the stream synchronisation after it finds the stream idle)
    for n, extra in (("1", {}), ("2", {}), ("4", {}), ("8", {}), ("8", {"PAGEABLE_OUT": "1"}), ("8", {"PINNED_IN": "1"}), ("8", {"NOGRAPH": "1"})):
        env = dict(os.environ, RIP_B200_COPY_THREADS=n, **extra)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
