#!/usr/bin/env python3
"""BASELINE config 5 through ONE call of the product's multi-GPU entry point (MultiGpuPipeline.process_batch ->
rip_apply_batch_host_multi): n_gpus x 64 frames of 3840x2160, full chain with CCC white balance + undistortion, pinned
host buffers; prints Mpix/s at 1, 2, 4, ... GPUs of the box and checks two frames against the oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from raw_image_pipeline_b200 import MultiGpuPipeline

cfg = bench.CONFIGS[5]
rows, cols, enc = cfg["rows"], cfg["cols"], cfg["enc"]
n_dev = torch.cuda.device_count()
per_gpu = int(os.environ.get("FRAMES_PER_GPU", "64"))
base = bench.make_frames(16, rows, cols, 5000, enc=enc)
results = []
only = [int(x) for x in os.environ.get("ONLY_GPUS", "").split(",") if x]
g = 1
while g <= n_dev:
    if only and g not in only:
        g *= 2
        continue
    n = per_gpu * g
    h_in = torch.from_numpy(np.concatenate([base] * ((n + 15) // 16))[:n]).pin_memory()
    h_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8).pin_memory()
    mp = MultiGpuPipeline(n_gpus=g, use_gpu=False, params_path="", calibration_path=os.path.join(ROOT, "raw_image_pipeline_b200", "config", "alphasense_calib_example.yaml"))
    # configure every replica like bench.make_pipeline
    ref_p = bench.make_pipeline  # noqa
    mp.set_flip(True); mp.set_flip_angle(180)
    mp.set_white_balance(True); mp.set_white_balance_method("ccc"); mp.set_white_balance_saturation_threshold(0.8, 0.2); mp.set_white_balance_temporal_consistency(False)
    mp.set_color_calibration(True); mp.set_color_calibration_matrix(bench.CC)
    mp.set_gamma_correction(True); mp.set_gamma_correction_method("custom"); mp.set_gamma_correction_k(0.8)
    mp.set_vignetting_correction(True); mp.set_vignetting_correction_parameters(1.5, 1e-3, 1e-6)
    mp.set_color_enhancer(True); mp.set_color_enhancer_saturation_gain(1.2)
    mp.set_undistortion_image_size(cols, rows); mp.set_undistortion_camera_matrix(bench.calib_K(rows, cols))
    mp.set_undistortion_distortion_coeffs(bench.CALIB_D); mp.set_undistortion_balance(0.0); mp.set_undistortion_fov_scale(0.8)
    mp.set_undistortion(True)
    for _ in range(2):
        mp.process_batch_ptr(h_in.data_ptr(), n, rows, cols, 1, enc, h_out.data_ptr())
    launches0 = [r.kernel_launches() for r in mp.replicas]
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        mp.process_batch_ptr(h_in.data_ptr(), n, rows, cols, 1, enc, h_out.data_ptr())
    dt = (time.perf_counter() - t0) / reps
    o = bench.make_oracle(rows, cols, "ccc")
    bad = 0
    for i in (0, n - 1):
        ref, _ = o.apply(h_in[i].numpy(), enc)
        bad += int(np.count_nonzero(h_out[i].numpy() != ref))
    launches = [r.kernel_launches() - l0 for r, l0 in zip(mp.replicas, launches0)]
    share = [round(l / max(1, sum(launches)), 3) for l in launches]   # kernel launches are proportional to the chunks a GPU took
    results.append({"gpus": g, "frames": n, "mpix_per_s": n * rows * cols / dt / 1e6, "ms": dt * 1e3, "differing_values_vs_oracle": bad,
                    "share_of_work_per_gpu": share})
    del mp, h_in, h_out
    g *= 2
print(json.dumps({"entry": "rip_apply_batch_host_multi (one process, one host thread per GPU, chunks of <= 16 frames claimed on demand)", "workload": cfg["workload"], "results": results}))
