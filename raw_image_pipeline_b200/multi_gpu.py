"""One box, several B200s: a `MultiGpuPipeline` owns one `RawImagePipeline` per GPU, configured identically, and
shards work over them with no collective (frames are independent; the only cross-frame state, the CCC
Kalman tracker, belongs to one camera stream and stays on one device).

    mp = MultiGpuPipeline(n_gpus=8, use_gpu=True)
    mp.set_white_balance_method("ccc")           # every setter of RawImagePipeline, applied to all replicas
    out = mp.process_batch(frames, "bayer_rggb8")   # BASELINE config 5: one call, frames cut into n_gpus chunks
    outs = mp.process_streams([s0, s1, ...], "bayer_bggr8")   # config 4: camera stream i lives on GPU i % n_gpus

The batch call goes through the C ABI's rip_apply_batch_host_multi (one host thread per GPU inside the library); the
stream call uses one Python thread per GPU (the C calls release the GIL).  The reference has no multi-GPU path
(raw_image_pipeline.cpp:193-196)."""
from __future__ import annotations

import ctypes
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .pipeline import RawImagePipeline, _raise
from .sharding import shard_range, stream_owner


class MultiGpuPipeline:
    def __init__(self, n_gpus: Optional[int] = None, use_gpu: bool = False, params_path: Optional[str] = None,
                 calibration_path: str = "", color_calibration_path: str = "", devices: Optional[Sequence[int]] = None):
        lib = L.load()
        if devices is None:
            if n_gpus is None:
                n_gpus = lib.rip_device_count()
                if n_gpus <= 0:
                    raise RuntimeError("no CUDA device available (raw_image_pipeline_b200 has no CPU fallback)")
            devices = list(range(n_gpus))
        if not devices:
            raise ValueError("MultiGpuPipeline needs at least one device")
        self.devices = list(devices)
        self.replicas: List[RawImagePipeline] = [
            RawImagePipeline(use_gpu, params_path, calibration_path, color_calibration_path, device=d) for d in self.devices]
        self._lib = lib

    @property
    def n_gpus(self) -> int:
        return len(self.replicas)

    def __getattr__(self, name):
        """set_* / load_* / reset_* / _set_* are applied to every replica; is_* / get_* / output_shape answer from replica 0
        (the replicas are configured identically)."""
        if name.startswith(("set_", "load_", "reset_", "init_", "_set_")):
            def fan_out(*args, **kw):
                for r in self.replicas:
                    getattr(r, name)(*args, **kw)
            return fan_out
        if name.startswith(("is_", "get_", "_get_")) or name in ("output_shape", "log", "debug_table"):
            return getattr(self.replicas[0], name)
        raise AttributeError(name)

    def kernel_launches(self) -> int:
        return sum(r.kernel_launches() for r in self.replicas)

    def shard(self, n_frames: int):
        """[(begin, end)] per GPU: contiguous chunks whose sizes differ by at most one (sharding.shard_range) -- the split
        process_streams and the torchrun bench use; process_batch deals chunks on demand instead."""
        return [shard_range(n_frames, i, self.n_gpus) for i in range(self.n_gpus)]

    def process_batch(self, frames: np.ndarray, encoding: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        """n frames host -> host over all GPUs (rip_apply_batch_host_multi: chunks of <= 16 frames claimed on demand by the
        pipelines' worker threads, so GPUs behind a slower host link take fewer frames)."""
        if frames.dtype != np.uint8 or frames.ndim not in (3, 4) or not frames.flags.c_contiguous:
            raise ValueError("frames must be a C-contiguous uint8 array (n, rows, cols[, channels])")
        n, rows, cols = frames.shape[:3]
        ch = frames.shape[3] if frames.ndim == 4 else 1
        orows, ocols, och = self.replicas[0].output_shape(frames.shape[1:], encoding)
        if out is None:
            out = np.empty((n, orows, ocols, och), np.uint8)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.uint8 or not out.flags.c_contiguous or not out.flags.writeable
              or out.size != n * orows * ocols * och):
            raise ValueError(f"out must be a writable C-contiguous uint8 array of {n}x{orows}x{ocols}x{och} values")
        self.process_batch_ptr(frames.ctypes.data, n, rows, cols, ch, encoding, out.ctypes.data)
        return out

    def process_batch_ptr(self, in_ptr: int, n: int, rows: int, cols: int, channels: int, encoding: str, out_ptr: int):
        orows, ocols, och = self.replicas[0].output_shape((rows, cols, channels), encoding)
        handles = (ctypes.c_void_p * self.n_gpus)(*[r._h for r in self.replicas])
        rc = self._lib.rip_apply_batch_host_multi(handles, self.n_gpus, in_ptr, rows * cols * channels, n, rows, cols, channels,
                                                  encoding.encode(), out_ptr, orows * ocols * och)
        if rc != L.RIP_OK:
            _raise(rc, (self._lib.rip_last_error(self.replicas[0]._h) or b"").decode())

    def process_streams(self, streams: Sequence[np.ndarray], encoding: str) -> List[np.ndarray]:
        """Each element of `streams` is the frame sequence (n_i, rows, cols[, ch]) of one camera; stream i is processed, in
        order, by GPU stream_owner(i) -- so a stream's CCC tracker state never leaves its device.  Streams that share a
        GPU run one after the other."""
        results: List[Optional[np.ndarray]] = [None] * len(streams)
        errors = []

        def worker(gpu: int):
            try:
                for i, frames in enumerate(streams):
                    if stream_owner(i, self.n_gpus) == gpu:
                        results[i] = self.replicas[gpu].process_batch(np.ascontiguousarray(frames), encoding)
            except BaseException as e:  # surfaced to the caller below
                errors.append((gpu, e))

        threads = [threading.Thread(target=worker, args=(g,)) for g in range(self.n_gpus)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0][1]
        return results  # type: ignore[return-value]
