#!/usr/bin/env python3
"""Benchmark of the RAW hot path (BASELINE.json metric: Mpix/s end-to-end at 12 MP full chain;
fused-kernel HBM GB/s vs peak).

    python bench.py --gpus 1 --steps 5 --warmup 3            # this framework (B200 kernels), BASELINE configs[2]
    python bench.py --config 2 | 4 | 5                        # the other BASELINE configurations
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # the reference's CPU path (cv2 oracle)
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU, frames sharded (weak scaling)

A "step" is one pass of the hot path over one batch of synthetic Bayer frames.  Default: BASELINE.json configs[2] =
64 x 4032x3040 bayer_rggb8, full chain (flip 180, pca WB, colour calibration, gamma, vignetting, enhancer, undistortion).
`value` = DEVICE-RESIDENT throughput (inputs and outputs in HBM, CUDA events), `e2e` = the same through the host-buffer
C-ABI entry point with H2D/D2H copies inside the timed region (the end-to-end number).  Prints ONE JSON line on rank 0.

  --config 2   1 x 1920x1080 bayer_bggr8, full chain: latency-bound -- a step = `--frames` single-frame calls;
               `value` = device-resident one-frame launches, `e2e` = RawImagePipeline::apply() from/to pageable numpy
               arrays; `latency_us` reports p50/p99 per frame with and without the CUDA-graph replay.
  --config 4   one 1080p camera stream per GPU (rank), >= 256 frames each, frame by frame host to host;
               `streams` reports per-stream p50/p99 latency and Mpix/s, `value`/`e2e` the aggregate.
  --config 5   64 x 3840x2160 per GPU, full chain with CCC white balance + undistortion, distribution N.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CC = [2.4276948, 0.21479778, -0.30818, 0.09277014, 1.1962607, -0.09772757, -0.24436986, -0.22239459, 2.099912]
CALIB_D = [-0.0396482888762527, -0.00367688950406141, 0.00391742438164282, -0.00178738156007817]
METRIC = "Mpix/s end-to-end at 12MP full chain"
CHAIN = ("debayer + flip 180 + {wb} white balance + colour calibration (example matrix) + gamma 0.8 + vignetting "
         "(1.5,1e-3,1e-6) + HSV enhancer (sat 1.2) + fisheye undistortion (balance 0, fov 0.8)")
CONFIGS = {
    2: dict(rows=1080, cols=1920, enc="bayer_bggr8", wb="pca", frames=64, mode="latency",
            workload="BASELINE configs[1]: 1920x1080 bayer_bggr8 frames one at a time, full chain = " + CHAIN.format(wb="pca")),
    3: dict(rows=3040, cols=4032, enc="bayer_rggb8", wb="pca", frames=64, mode="batch",
            workload="BASELINE configs[2]: batch of 4032x3040 bayer_rggb8 frames, full chain = " + CHAIN.format(wb="pca")),
    4: dict(rows=1080, cols=1920, enc="bayer_bggr8", wb="pca", frames=256, mode="latency",
            workload="BASELINE configs[3]: one 1920x1080 bayer_bggr8 camera stream per GPU, frame by frame, full chain = " + CHAIN.format(wb="pca")),
    5: dict(rows=2160, cols=3840, enc="bayer_rggb8", wb="ccc", frames=64, mode="batch",
            workload="BASELINE configs[4]: batch of 3840x2160 bayer_rggb8 frames per GPU, full chain = " + CHAIN.format(wb="ccc (bright 0.8, dark 0.2, no temporal consistency)")),
}
ROWS, COLS, ENC = 3040, 4032, "bayer_rggb8"  # config 3 (module-level names kept for tools/)


def calib_K(rows, cols):
    sx, sy = cols / 720.0, rows / 540.0
    return [347.548139773951 * sx, 0.0, 342.454373227748 * sx, 0.0, 347.434712422309 * sy, 271.368057185649 * sy, 0.0, 0.0, 1.0]


def make_pipeline(rows, cols, device=None, wb="pca"):
    from raw_image_pipeline_b200 import RawImagePipeline
    cfg = os.path.join(ROOT, "raw_image_pipeline_b200", "config")
    p = RawImagePipeline(False, "", os.path.join(cfg, "alphasense_calib_example.yaml"), "", device=device)
    p.set_flip(True); p.set_flip_angle(180)
    p.set_white_balance(True); p.set_white_balance_method(wb)
    p.set_white_balance_saturation_threshold(0.8, 0.2); p.set_white_balance_temporal_consistency(False)
    p.set_color_calibration(True); p.set_color_calibration_matrix(CC)
    p.set_gamma_correction(True); p.set_gamma_correction_method("custom"); p.set_gamma_correction_k(0.8)
    p.set_vignetting_correction(True); p.set_vignetting_correction_parameters(1.5, 1e-3, 1e-6)
    p.set_color_enhancer(True); p.set_color_enhancer_saturation_gain(1.2)
    p.set_undistortion_image_size(cols, rows); p.set_undistortion_camera_matrix(calib_K(rows, cols))
    p.set_undistortion_distortion_coeffs(CALIB_D); p.set_undistortion_balance(0.0); p.set_undistortion_fov_scale(0.8)
    p.set_undistortion(True)
    return p


def make_witness_pipeline(device=None):
    """SURVEY 8d 'HBM-roofline witness': the same fused kernel with only debayer + gamma enabled (BASELINE configs[0]'s
    module set) -- what the tile machinery sustains when the per-pixel arithmetic is light."""
    from raw_image_pipeline_b200 import RawImagePipeline
    p = RawImagePipeline(False, "", "", "", device=device)
    for name in ("white_balance", "color_calibration", "vignetting_correction", "color_enhancer", "undistortion", "flip"):
        getattr(p, "set_" + name)(False)
    p.set_gamma_correction(True); p.set_gamma_correction_method("custom"); p.set_gamma_correction_k(0.8)
    return p


def make_oracle(rows, cols, wb="pca"):
    """The reference's CPU path (call-for-call cv2 replay): the timed CPU baseline and the checker of the timed outputs."""
    from oracle import cv2_oracle as O
    op = O.OracleParams(flip_enabled=True, flip_angle=180, wb_enabled=True, wb_method=wb, wb_bright_thr=0.8, wb_dark_thr=0.2,
                        wb_temporal_consistency=False, cc_enabled=True, cc_matrix=CC,
                        gamma_enabled=True, gamma_k=0.8, vig_enabled=True, enh_enabled=True, enh_saturation_gain=1.2,
                        und_enabled=True, und_K=calib_K(rows, cols), und_D=CALIB_D, und_width=cols, und_height=rows,
                        und_balance=0.0, und_fov_scale=0.8)
    return O.OraclePipeline(op, os.path.join(ROOT, "raw_image_pipeline_b200", "config", "ccc_model.bin") if wb == "ccc" else None)


def make_frames(n, rows, cols, seed0, distinct=16, enc=ENC):
    """n frames, of which min(n, 16) are distinct (generated on the host with Philox) and the rest repeats of them."""
    from raw_image_pipeline_b200 import synth
    d = min(n, distinct)
    base = synth.bayer_batch(d, rows, cols, enc, seed0, "N")
    if d == n:
        return base
    return np.concatenate([base] * ((n + d - 1) // d))[:n]


class ClockSampler:
    """SM clock and clock-event reasons sampled through NVML (in-process thread, every ~5 ms)
    DURING the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def start(self):
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.nv = nv
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            time.sleep(0.005)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        reasons = set()
        for _, r, _ in self.samples:
            for bit, name in names.items():
                if r & bit:
                    reasons.add(name)
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max((s[2] for s in self.samples), default=None)}


def bind_to_gpu_numa_node(local):
    """Multi-GPU runs: pin this rank's host thread (and, by first touch, its pinned staging buffers) to the NUMA node its
    GPU hangs off, so that the e2e leg's H2D/D2H traffic of eight ranks does not cross the socket interconnect."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def time_cpu_reference(cfg, n_sample, warm, threads=None):
    import cv2
    if threads:
        cv2.setNumThreads(threads)
    rows, cols, enc = cfg["rows"], cfg["cols"], cfg["enc"]
    o = make_oracle(rows, cols, cfg["wb"])
    frames = make_frames(n_sample, rows, cols, 3000, enc=enc)
    for i in range(warm):
        o.apply(frames[i % n_sample], enc)
    t0 = time.perf_counter()
    for i in range(n_sample):
        o.apply(frames[i], enc)
    dt = time.perf_counter() - t0
    return n_sample * rows * cols / dt / 1e6, cv2.getNumThreads(), dt


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    rows, cols, enc = cfg["rows"], cfg["cols"], cfg["enc"]
    n_sample = args.ref_frames
    o = make_oracle(rows, cols, cfg["wb"])
    frames = make_frames(n_sample, rows, cols, 3000, enc=enc)
    for _ in range(args.warmup):
        o.apply(frames[0], enc)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(n_sample):
            o.apply(frames[i], enc)
    dt = time.perf_counter() - t0
    value = args.steps * n_sample * rows * cols / dt / 1e6
    cores = cv2.getNumThreads()
    sample = (f"{n_sample} frames of {cols}x{rows} per step (bounded sample of the workload's frames), cv2 {cv2.__version__}, "
              f"{cores} threads; vignetting mask cached across frames (flatters the reference, SURVEY B-7)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["workload"], "bench_config": args.config, "frames_per_step": n_sample, "distribution": "N"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def copy_ceiling(world):
    """Measured host<->device copy ceiling of the box at `world` GPUs (tools/pcie_probe, profiles/pcie_ceiling.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "pcie_ceiling.json")) as f:
            return json.load(f)["d2h_gbs_with_concurrent_h2d"].get(str(world))
    except Exception:
        return None


def percentiles(us):
    a = np.sort(np.asarray(us, np.float64))
    return {"p50": float(a[len(a) // 2]), "p99": float(a[min(len(a) - 1, int(len(a) * 0.99))]), "mean": float(a.mean()), "n": int(len(a))}


class Job:
    """rank / device / barrier plumbing shared by the modes"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        from raw_image_pipeline_b200 import sharding
        return sharding.max_over_ranks(x, device=self.dev)

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def check_parity(p, cfg, frames, outputs, which):
    """outputs[i] (numpy) against the oracle for the frame indices `which`"""
    o = make_oracle(cfg["rows"], cfg["cols"], cfg["wb"])
    nbad, maxd = 0, 0
    for i in which:
        ref, _ = o.apply(frames[i], cfg["enc"])
        d = np.abs(outputs[i].astype(np.int16) - ref.astype(np.int16))
        nbad += int(np.count_nonzero(d)); maxd = max(maxd, int(d.max()))
    return {"frames_checked": len(which), "frames": list(which), "max_abs_diff": maxd, "differing_values": nbad,
            "against": "oracle/cv2_oracle.py (cv2 replay of the reference CPU path), outputs of the timed call"}


def run_batch(args, cfg):
    """configs 3 and 5: one batch of frames per step and GPU"""
    job = Job()
    torch = job.torch
    rank, world, local, dev = job.rank, job.world, job.local, job.dev
    n, rows, cols, enc = args.frames, cfg["rows"], cfg["cols"], cfg["enc"]
    p = make_pipeline(rows, cols, device=local, wb=cfg["wb"])
    frames = make_frames(n, rows, cols, 3000 + 100 * rank, enc=enc)
    h_in = torch.from_numpy(frames).pin_memory()
    d_in = h_in.to(dev, non_blocking=True)
    d_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    def step_device():
        p.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=stream)

    for _ in range(args.warmup):
        step_device()
    job.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = p.kernel_launches()
    p._set_bool("profile/kernel_events", True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    job.barrier()
    ms_total = job.max_over_ranks(ev0.elapsed_time(ev1))
    p._set_bool("profile/kernel_events", False)
    kernel_ms = p._get_doubles("stats/kernel_ms")
    launches = p.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * n * rows * cols * args.steps / (ms_total * 1e-3) / 1e6

    # ---- HBM-roofline witness: debayer + gamma only, same buffers -----------------------------------
    witness = None
    if not args.no_witness and args.config == 3:
        pw = make_witness_pipeline(device=local)
        def step_w():
            pw.process_batch_ptr(d_in.data_ptr(), n, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=stream)
        for _ in range(3):
            step_w()
        torch.cuda.synchronize()
        pw._set_bool("profile/kernel_events", True)
        for _ in range(args.steps):
            step_w()
        torch.cuda.synchronize()
        kw = pw._get_doubles("stats/kernel_ms")
        pw._set_bool("profile/kernel_events", False)
        w_ms = kw[2] / max(kw[4 + 2], 1.0)
        witness = {"modules": "debayer + gamma (k=0.8)", "kernel": "k_fused_strip<gamma, BGR8>", "avg_launch_ms": w_ms,
                   "achieved_gbs": 4.0 * n * rows * cols / (w_ms * 1e-3) / 1e9 if w_ms > 0 else None,
                   "mpix_per_s": n * rows * cols / (w_ms * 1e-3) / 1e6 if w_ms > 0 else None}
        del pw

    # ---- end to end through the host-buffer entry point (pinned host memory, copies inside) ------
    h_out = torch.empty((n, rows, cols, 3), dtype=torch.uint8).pin_memory()

    def step_host():
        p.process_batch_ptr(h_in.data_ptr(), n, rows, cols, 1, enc, h_out.data_ptr(), host=True)

    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, args.e2e_steps))
    e2e, dt_e2e = None, 0.0
    if e2e_steps:
        for _ in range(min(args.warmup, 2)):
            step_host()
        job.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        dt_e2e = job.max_over_ranks(time.perf_counter() - t0)
        job.barrier()
        e2e = world * n * rows * cols * e2e_steps / dt_e2e / 1e6
    ccc = None
    if cfg["wb"] == "ccc":  # the estimate of the last frame of the last call: shows the white balance did real work
        ccc = {"uv_last_frame": list(p.ccc_uv()), "gains_bgr_last_frame": [float(g) for g in p.ccc_gains()],
               "constructor_value": [128, 128]}
        try:
            resp = np.frombuffer(p.debug_table("ccc_response"), np.float64)
            top = np.sort(resp)[-2:]
            ccc["response_top1_minus_top2"] = float(top[1] - top[0])
            ccc["response_range"] = float(resp.max() - resp.min())
        except Exception as e:  # pragma: no cover
            ccc["response_error"] = repr(e)
    if rank != 0:
        job.finish()
        return

    # ---- parity of the timed outputs: first and last frame of the step against the cv2 oracle (checker only) ----
    parity = None
    if not args.no_parity:
        step_device()
        torch.cuda.synchronize()
        which = sorted({0, n - 1})
        parity = check_parity(p, cfg, frames, {i: d_out[i].cpu().numpy() for i in which}, which)
        if e2e_steps:
            parity["host_path_equals_device_path"] = bool(torch.equal(d_out[0].cpu(), h_out[0]) and torch.equal(d_out[n - 1].cpu(), h_out[n - 1]))

    peak, peak_src = hbm_peak()
    n_fused = max(kernel_ms[4 + 2], 1.0)
    fused_ms = kernel_ms[2] / n_fused
    algo_bytes = 4.0 * n * rows * cols  # 1 B Bayer read + 3 B BGR8 write per pixel
    achieved = algo_bytes / (fused_ms * 1e-3) / 1e9 if fused_ms > 0 else 0.0
    step_kernel_ms = {k: kernel_ms[i] / args.steps for i, k in enumerate(["pca_stats", "pca_lut", "fused", "remap"])}
    def per_launch(i):
        return kernel_ms[i] / max(kernel_ms[4 + i], 1.0)
    px = float(n) * rows * cols
    other = {
        # SURVEY 8d counts 14 B/px (3 gather + 8 fp32 map + 3 write); the kernel is DESIGNED to move 11 (4-byte intermediate,
        # packed 4-byte map, 3 B out): `moved_frac_of_peak` is the honest figure, `frac_of_peak` charges bytes it never moves
        # moved by design: 4-byte intermediate read + 3-byte output + the 4-byte packed map once per group of 8 frames
        "k_remap_tile": {"algorithmic_bytes_per_px": 14, "moved_bytes_per_px": 7.5, "avg_launch_ms": per_launch(3),
                         "achieved_gbs": 14 * px / (per_launch(3) * 1e-3) / 1e9 if per_launch(3) > 0 else None,
                         "moved_gbs": 7.5 * px / (per_launch(3) * 1e-3) / 1e9 if per_launch(3) > 0 else None},
        "whole_step": {"algorithmic_bytes_per_px": 19, "ms": ms_total / args.steps,
                       "achieved_gbs": 19 * px / (ms_total / args.steps * 1e-3) / 1e9},
    }
    if cfg["wb"] == "pca":
        other["k_pca_stats"] = {"algorithmic_bytes_per_px": 1, "avg_launch_ms": per_launch(0),
                                "achieved_gbs": px / (per_launch(0) * 1e-3) / 1e9 if per_launch(0) > 0 else None}
    for v in other.values():
        if v.get("achieved_gbs"):
            v["frac_of_peak"] = v["achieved_gbs"] / peak
        if v.get("moved_gbs"):
            v["moved_frac_of_peak"] = v["moved_gbs"] / peak
    if witness and witness.get("achieved_gbs"):
        witness["frac_of_peak"] = witness["achieved_gbs"] / peak
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "fused_traffic.json")) as f:
            tj = json.load(f)
        if args.config == 3 and n == 64:
            traffic = tj.get("dram_bytes_per_launch_at_bench_size"); traffic_src = tj.get("source")
            wt = tj.get("other_kernels", {}).get("k_fused_strip<gamma, BGR8> (witness)")
            if witness and wt:
                witness["traffic"] = wt.get("dram_bytes_per_launch")
                witness["algorithmic_bytes_per_launch"] = wt.get("algorithmic_bytes_per_launch")
    except Exception:
        pass
    ceiling = copy_ceiling(world)
    d2h_gbs = 3.0 * world * n * rows * cols * e2e_steps / dt_e2e / 1e9 if e2e_steps else None
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["workload"], "bench_config": args.config, "frames_per_step_per_gpu": n,
                   "value_is": "device-resident throughput (inputs and outputs in HBM, CUDA events on the launching stream); "
                               "the end-to-end number of the metric's name is e2e.value",
                   "distribution": "N (natural-ish) + sensor cast; min(frames, 16) distinct frames per GPU, repeated to fill the batch",
                   "l2": f"inputs {n * rows * cols / 1e6:.0f} MB + outputs {3 * n * rows * cols / 1e6:.0f} MB per step >> 126 MB L2 "
                         "(no flush needed)", "parallelism": f"frames sharded over {world} GPU(s), no collective", "host_numa_binding_rank0": job.numa,
                   "kernel_ms_per_step": step_kernel_ms},
        "roofline": {"bound": "hbm", "kernel": "k_fused<all stages, Bayer -> 4-byte intermediate>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": fused_ms,
                     "note": "full chain at bit-exact parity is bound by instruction issue (80 %) and the shared-memory pipe of its table lookups (78 %), not by HBM (DESIGN.md section 8)",
                     "other_kernels": other, "witness_debayer_gamma": witness},
        "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": int(n * rows * cols) * world,
                "d2h_bytes_per_step": int(3 * n * rows * cols) * world, "steps": e2e_steps,
                "ms_per_step": dt_e2e / max(e2e_steps, 1) * 1e3, "d2h_gbs": d2h_gbs, "copy_ceiling_d2h_gbs": ceiling,
                "frac_of_copy_ceiling": (d2h_gbs / ceiling) if (d2h_gbs and ceiling) else None,
                "api": "rip_apply_batch_host (pinned host buffers, H2D + kernels + D2H inside the timed region, host wall clock)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity": parity,
    }
    if ccc:
        line["ccc"] = ccc
    if world == 1 and not args.no_cpu_baseline:
        v, cores, dt = time_cpu_reference(cfg, args.cpu_frames, 1)
        import cv2
        line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_frames} frames of {cols}x{rows} through the cv2 call-for-call replay of the "
                                          f"reference CPU path (cv2 {cv2.__version__}, {cores} threads, {dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    job.finish()
    if parity and parity["differing_values"] != 0:
        raise SystemExit("bench.py: the timed outputs differ from the oracle: " + json.dumps(parity))


def run_latency(args, cfg):
    """configs 2 and 4: frames one at a time.  A step = `--frames` consecutive single-frame calls of one camera stream."""
    job = Job()
    torch = job.torch
    rank, world, local, dev = job.rank, job.world, job.local, job.dev
    n, rows, cols, enc = args.frames, cfg["rows"], cfg["cols"], cfg["enc"]
    p = make_pipeline(rows, cols, device=local, wb=cfg["wb"])
    frames = make_frames(n, rows, cols, 3000 + 100 * rank, enc=enc)  # pageable numpy, like a caller's images
    px = rows * cols
    d_in = torch.from_numpy(frames).to(dev)
    d_out = torch.empty((rows, cols, 3), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    def step_device():  # one launch sequence (stats, lut, fused, remap) per frame, device-resident
        for i in range(n):
            p.process_batch_ptr(d_in[i].data_ptr(), 1, rows, cols, 1, enc, d_out.data_ptr(), host=False, stream=stream)

    for _ in range(args.warmup):
        step_device()
    job.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = p.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    job.barrier()
    ms_total = job.max_over_ranks(ev0.elapsed_time(ev1))
    launches = p.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * n * px * args.steps / (ms_total * 1e-3) / 1e6
    # kernel time alone (CUDA events around every kernel)
    p._set_bool("profile/kernel_events", True)
    step_device()
    torch.cuda.synchronize()
    kms = p._get_doubles("stats/kernel_ms")
    p._set_bool("profile/kernel_events", False)
    kernel_us_per_frame = float(sum(kms[:4])) / n * 1e3

    # ---- end to end: RawImagePipeline::apply() frame by frame, pageable numpy in, numpy out -------
    def timed_apply(steps, src=None):
        src = frames if src is None else src
        lat = []
        outs = {}
        t_begin = time.perf_counter()
        for s in range(steps):
            for i in range(n):
                t0 = time.perf_counter()
                out = p.process(src[i % len(src)], enc)
                lat.append((time.perf_counter() - t0) * 1e6)
                if s == steps - 1 and i in (0, n - 1):
                    outs[i] = out
        return lat, outs, time.perf_counter() - t_begin

    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, args.e2e_steps))
    e2e, lat_graph, lat_plain, lat_pinned, lat_registered, outs, dt_e2e = None, None, None, None, None, {}, 0.0
    if e2e_steps:
        timed_apply(1)  # warm-up: buffers, lazy tables, graph capture
        job.barrier()
        lat, outs, dt = timed_apply(e2e_steps)
        dt_e2e = job.max_over_ranks(dt)
        e2e = world * n * px * e2e_steps / dt_e2e / 1e6
        lat_graph = percentiles(lat)
        replays = p._get_int("stats/graph_replays")
        if args.config == 2:  # the same without the CUDA-graph replay
            p._set_bool("apply/cuda_graph", False)
            timed_apply(1)
            lat_plain = percentiles(timed_apply(e2e_steps)[0])
            p._set_bool("apply/cuda_graph", True)
        # SURVEY 8d config 4 "pinned double buffers": the camera frames arrive in page-locked buffers (two, used in turn)
        pinned = [p.pinned_empty((rows, cols)) for _ in range(2)]
        if all(a is not None for a in pinned):
            def pinned_frames():
                for i in range(n):
                    pinned[i & 1][...] = frames[i]
                    yield pinned[i & 1]
            lat_p = []
            for rep in range(1 + e2e_steps):  # pass 0 warms up (graph capture for the new buffer kind) and is not counted
                for a in pinned_frames():
                    t0 = time.perf_counter()
                    p.process(a, enc)
                    if rep:
                        lat_p.append((time.perf_counter() - t0) * 1e6)
            lat_pinned = percentiles(lat_p)
        # opt-in: the library page-locks the caller's ordinary buffers the first time it sees them (a camera ring that is reused)
        p._set_bool("apply/register_caller_buffers", True)
        ring = [np.empty((rows, cols), np.uint8) for _ in range(8)]   # a camera driver's ring of ordinary buffers
        lat_r = []
        for rep in range(1 + e2e_steps):  # pass 0 registers the ring and is not counted
            for i in range(n):
                ring[i & 7][...] = frames[i]
                t0 = time.perf_counter()
                p.process(ring[i & 7], enc)
                if rep:
                    lat_r.append((time.perf_counter() - t0) * 1e6)
        lat_registered = percentiles(lat_r)
        p._set_bool("apply/register_caller_buffers", False)
    per_stream = job.gather({"rank": rank, "latency_us": lat_graph, "mpix_per_s": (n * px * e2e_steps / dt_e2e / 1e6) if e2e_steps else None})
    if rank != 0:
        job.finish()
        return
    parity = None
    if not args.no_parity and outs:
        parity = check_parity(p, cfg, frames, outs, sorted(outs))
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["workload"], "bench_config": args.config, "frames_per_step_per_gpu": n,
                   "value_is": "device-resident throughput of one-frame launch sequences (frames in HBM, CUDA events); "
                               "the end-to-end number is e2e.value (RawImagePipeline::apply, pageable host arrays, one frame per call)",
                   "distribution": "N (natural-ish) + sensor cast", "l2": "latency-bound: one 2 MP frame per call (fits L2)",
                   "parallelism": f"one camera stream per GPU, {world} GPU(s), no collective", "host_numa_binding_rank0": job.numa},
        "latency_us": {"kernels_only_per_frame": kernel_us_per_frame,
                       "device_resident_call_per_frame": ms_total / args.steps / n * 1e3,
                       "apply_host_to_host_cuda_graph": lat_graph, "apply_host_to_host_no_graph": lat_plain,
                       "apply_page_locked_input_buffers": lat_pinned,
                       "apply_registered_caller_buffers_opt_in": lat_registered,
                       "graph_replays": replays if e2e_steps else None},
        "streams": per_stream,
        "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": int(n * px) * world, "d2h_bytes_per_step": int(3 * n * px) * world,
                "steps": e2e_steps, "ms_per_step": dt_e2e / max(e2e_steps, 1) * 1e3,
                "api": "rip_apply (pageable host arrays -> pinned staging -> one CUDA-graph launch -> pinned staging -> host array)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity": parity,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, dt = time_cpu_reference(cfg, min(args.cpu_frames * 4, n), 1)
        import cv2
        line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
                                "sample": f"{min(args.cpu_frames * 4, n)} frames of {cols}x{rows} through the cv2 call-for-call replay of the "
                                          f"reference CPU path (cv2 {cv2.__version__}, {cores} threads, {dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    job.finish()
    if parity and parity["differing_values"] != 0:
        raise SystemExit("bench.py: the timed outputs differ from the oracle: " + json.dumps(parity))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE configuration (SURVEY 8d); default 3 = configs[2]")
    ap.add_argument("--frames", type=int, default=None, help="frames per step per GPU (default: the configuration's)")
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames timed for cpu_baseline")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames per step for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-witness", action="store_true", help="skip the debayer+gamma-only roofline witness")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed outputs")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.frames is None:
        args.frames = cfg["frames"]
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, cfg)
    elif cfg["mode"] == "batch":
        run_batch(args, cfg)
    else:
        if args.steps > 5 and args.config == 4:
            pass
        run_latency(args, cfg)


if __name__ == "__main__":
    main()
