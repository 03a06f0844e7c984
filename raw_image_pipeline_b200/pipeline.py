"""Python mirror of the reference's pybind class (raw_image_pipeline_python/src/
raw_image_pipeline_python.cpp:16-73): same class name, same snake_case method names, same
argument meaning -- implemented over the C ABI of librip_b200.so.

Extra (not bound by the reference's pybind module but present on its C++ class,
raw_image_pipeline.hpp:53-56,134-137): load_camera_calibration, load_color_calibration,
init_undistortion, get_dist_debayered_image, get_dist_color_image, get_rect_mask,
get_processed_image; and the batch entry points process_batch / process_batch_device.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


class RawImagePipelineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _raise(code: int, msg: str):
    if code == L.RIP_ERR_INVALID_ARGUMENT:
        # the reference throws std::invalid_argument, which pybind11 maps to ValueError
        raise ValueError(msg)
    raise RawImagePipelineError(code, msg)


def _is_16bit(encoding: str) -> bool:
    return encoding in ("bayer_rggb16", "bayer_bggr16", "bayer_gbrg16", "bayer_grbg16")


class _PinnedPool:
    """Result arrays of process() backed by page-locked memory (rip_pinned_alloc): the GPU's copy engine writes the frame
    straight into the array the caller receives, instead of into a staging buffer that is then copied (and page-faulted)
    into a fresh pageable array.  A buffer goes back to the pool when the last array viewing it is garbage-collected;
    when the caller holds on to more results than the pool is allowed to pin, process() falls back to pageable arrays."""

    MAX_BYTES = 512 << 20
    MAX_SLOTS = 16

    def __init__(self, lib):
        self._lib = lib
        self._free = []   # (ptr, capacity)
        self._total = 0
        self._slots = 0
        self._closed = False

    def take(self, nbytes: int):
        for i, (ptr, cap) in enumerate(self._free):
            if cap >= nbytes:
                del self._free[i]
                return self._wrap(ptr, cap, nbytes)
        if self._slots >= self.MAX_SLOTS or self._total + nbytes > self.MAX_BYTES:
            if not self._free:
                return None
            ptr, cap = self._free.pop()   # too small for this frame size: replace it
            self._lib.rip_pinned_free(ptr)
            self._total -= cap
            self._slots -= 1
            if self._total + nbytes > self.MAX_BYTES:
                return None
        ptr = ctypes.c_void_p()
        if self._lib.rip_pinned_alloc(nbytes, ctypes.byref(ptr)) != L.RIP_OK or not ptr.value:
            return None
        self._total += nbytes
        self._slots += 1
        return self._wrap(ptr.value, nbytes, nbytes)

    def _wrap(self, ptr: int, cap: int, nbytes: int) -> np.ndarray:
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, ptr, cap)   # runs when the last array viewing `buf` is gone
        return np.frombuffer(buf, np.uint8)

    def _release(self, ptr: int, cap: int):
        if self._closed:
            self._lib.rip_pinned_free(ptr)
        else:
            self._free.append((ptr, cap))

    def close(self):
        self._closed = True
        for ptr, _ in self._free:
            self._lib.rip_pinned_free(ptr)
        self._free = []


class RawImagePipeline:
    def __init__(self, use_gpu: bool = False, params_path: Optional[str] = None, calibration_path: str = "",
                 color_calibration_path: str = "", device: Optional[int] = None):
        """``RawImagePipeline(use_gpu)`` == the reference's 1-argument constructor (loads the default
        params, camera calibration and colour calibration); with ``params_path`` given (may be "")
        it is the 4-argument constructor (raw_image_pipeline.cpp:16-40)."""
        self._lib = L.load()
        self._h = ctypes.c_void_p()
        if params_path is None and not calibration_path and not color_calibration_path:
            rc = self._lib.rip_create_default(int(bool(use_gpu)), ctypes.byref(self._h))
        else:
            rc = self._lib.rip_create(int(bool(use_gpu)), (params_path or "").encode(), calibration_path.encode(),
                                      color_calibration_path.encode(), ctypes.byref(self._h))
        if rc != L.RIP_OK:
            _raise(rc, (self._lib.rip_last_error(None) or b"").decode())
        if device is not None:
            self._check(self._lib.rip_set_device(self._h, int(device)))
        self._pool = _PinnedPool(self._lib)
        self.use_pinned_results = True   # process() returns arrays backed by page-locked memory (see _PinnedPool)

    def __del__(self):
        pool = getattr(self, "_pool", None)
        if pool is not None:
            pool.close()
        h = getattr(self, "_h", None)
        if h:
            self._lib.rip_destroy(h)
            self._h = None

    # ---- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != L.RIP_OK:
            _raise(rc, (self._lib.rip_last_error(self._h) or b"").decode())

    def _set_bool(self, key, v): self._check(self._lib.rip_set_bool(self._h, key.encode(), int(bool(v))))
    def _set_int(self, key, v): self._check(self._lib.rip_set_int(self._h, key.encode(), int(v)))
    def _set_double(self, key, v): self._check(self._lib.rip_set_double(self._h, key.encode(), float(v)))
    def _set_string(self, key, v): self._check(self._lib.rip_set_string(self._h, key.encode(), str(v).encode()))

    def _set_doubles(self, key, values: Sequence[float]):
        arr = (ctypes.c_double * len(values))(*[float(v) for v in values])
        self._check(self._lib.rip_set_doubles(self._h, key.encode(), arr, len(values)))

    def _get_bool(self, key) -> bool:
        v = ctypes.c_int()
        self._check(self._lib.rip_get_bool(self._h, key.encode(), ctypes.byref(v)))
        return bool(v.value)

    def _get_int(self, key) -> int:
        v = ctypes.c_int()
        self._check(self._lib.rip_get_int(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def _get_double(self, key) -> float:
        v = ctypes.c_double()
        self._check(self._lib.rip_get_double(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def _get_string(self, key) -> str:
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self._lib.rip_get_string(self._h, key.encode(), buf, len(buf)))
        return buf.value.decode()

    def _get_doubles(self, key, shape=None) -> np.ndarray:
        arr = (ctypes.c_double * 16)()
        n = ctypes.c_int()
        self._check(self._lib.rip_get_doubles(self._h, key.encode(), arr, 16, ctypes.byref(n)))
        out = np.array(arr[:n.value], dtype=np.float64)
        return out.reshape(shape) if shape else out

    def set_register_caller_buffers(self, enabled: bool):
        """Opt-in ("apply/register_caller_buffers"): process()/apply() page-lock the image buffers they are handed the first
        time they see them and copy straight from them afterwards -- for callers that cycle through a fixed set of arrays
        (a camera ring).  Those arrays must stay alive (mapped) as long as this pipeline does."""
        self._set_bool("apply/register_caller_buffers", enabled)

    def pinned_empty(self, shape, dtype=np.uint8) -> Optional[np.ndarray]:
        """An uninitialised array in page-locked memory from the pipeline's pool (None when the pool is exhausted): frames
        placed in such an array are uploaded by the copy engine directly, without the staging copy of pageable images."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        flat = self._pool.take(n)
        return None if flat is None else flat.view(dt).reshape(shape)

    # ---- main interfaces ---------------------------------------------------------------------
    def output_shape(self, image_shape: Tuple[int, ...], encoding: str) -> Tuple[int, int, int]:
        rows, cols = image_shape[0], image_shape[1]
        ch = image_shape[2] if len(image_shape) == 3 else 1
        r, c, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.rip_output_shape(self._h, rows, cols, ch, encoding.encode(), ctypes.byref(r),
                                               ctypes.byref(c), ctypes.byref(k)))
        return r.value, c.value, k.value

    def _run(self, image: np.ndarray, encoding: str) -> Tuple[np.ndarray, str]:
        want = np.uint16 if self._reads_16bit(encoding) else np.uint8
        if _is_16bit(encoding) and want is np.uint8 and image.ndim in (2, 3):
            self.output_shape(image.shape, encoding)  # extension off: the reference's "valid pattern but is not supported" comes first
        if image.dtype != want or image.ndim not in (2, 3):
            raise ValueError(f"image must be a {np.dtype(want).name} array of shape (rows, cols) or (rows, cols, channels)")
        item = image.dtype.itemsize
        img = image if image.strides[-1] == item and (image.ndim == 2 or image.strides[1] == image.shape[2] * item) \
            else np.ascontiguousarray(image)
        rows, cols = img.shape[0], img.shape[1]
        ch = img.shape[2] if img.ndim == 3 else 1
        orows, ocols, och = self.output_shape(img.shape, encoding)
        # page-locked when the pool has room: no staging copy on the way out
        flat = self._pool.take(orows * ocols * och) if self.use_pinned_results else None
        out = flat.reshape(orows, ocols, och) if flat is not None else np.empty((orows, ocols, och), np.uint8)
        enc = ctypes.create_string_buffer(encoding.encode(), 64)
        r, c, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.rip_apply(self._h, img.ctypes.data, rows, cols, ch, img.strides[0], enc, 64,
                                        out.ctypes.data, out.nbytes, ctypes.byref(r), ctypes.byref(c), ctypes.byref(k)))
        if och == 1:
            out = out[:, :, 0]  # a 1-channel cv::Mat is a 2-D array (cvnp convention of the reference binding)
        return out, enc.value.decode()

    def process(self, image: np.ndarray, encoding: str) -> np.ndarray:
        """RawImagePipeline::process (raw_image_pipeline.cpp:182-188): returns the processed copy."""
        out, _ = self._run(image, encoding)
        return out

    def apply(self, image: np.ndarray, encoding: str) -> bool:
        """RawImagePipeline::apply (raw_image_pipeline.cpp:190-205).  Like the reference's binding
        it mutates ``image`` in place when the result has the same shape (e.g. a ``bgr8`` input)
        and always returns True; when the shape changes (Bayer 1ch -> BGR 3ch) a numpy array cannot
        be re-seated -- fetch the result with get_processed_image() or use process()."""
        out, _ = self._run(image, encoding)
        if out.shape == image.shape:
            image[...] = out
        return True

    def process_batch(self, frames: np.ndarray, encoding: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        """n frames host -> host through rip_apply_batch_host (pinned memory gives full PCIe rate)."""
        want = np.uint16 if self._reads_16bit(encoding) else np.uint8
        if _is_16bit(encoding) and want is np.uint8 and frames.ndim in (3, 4):
            self.output_shape(frames.shape[1:], encoding)
        if frames.dtype != want or frames.ndim not in (3, 4) or not frames.flags.c_contiguous:
            raise ValueError(f"frames must be a C-contiguous {np.dtype(want).name} array (n, rows, cols[, channels])")
        n, rows, cols = frames.shape[:3]
        ch = frames.shape[3] if frames.ndim == 4 else 1
        orows, ocols, och = self.output_shape(frames.shape[1:], encoding)
        if out is None:
            out = np.empty((n, orows, ocols, och), np.uint8)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.uint8 or not out.flags.c_contiguous or not out.flags.writeable
              or out.size != n * orows * ocols * och):
            # the C entry point writes n tightly packed frames through the raw pointer: anything else corrupts memory
            raise ValueError(f"out must be a writable C-contiguous uint8 array of {n}x{orows}x{ocols}x{och} values")
        self.process_batch_ptr(frames.ctypes.data, n, rows, cols, ch, encoding, out.ctypes.data, host=True,
                               in_frame_stride=rows * cols * ch * frames.dtype.itemsize)
        return out

    def process_batch_ptr(self, in_ptr: int, n: int, rows: int, cols: int, channels: int, encoding: str, out_ptr: int,
                          host: bool, dist_color_ptr: int = 0, stream: int = 0,
                          in_frame_stride: Optional[int] = None, out_frame_stride: Optional[int] = None):
        """Raw-pointer batch entry (device or host memory); see rip_apply_batch_device/_host."""
        orows, ocols, och = self.output_shape((rows, cols, channels), encoding)
        bps = 2 if self._reads_16bit(encoding) else 1
        ins = in_frame_stride if in_frame_stride is not None else rows * cols * channels * bps
        outs = out_frame_stride if out_frame_stride is not None else orows * ocols * och
        if n <= 0 or ins < rows * cols * channels * bps or outs < orows * ocols * och:
            raise ValueError("process_batch_ptr: n must be positive and the frame strides at least one frame")
        if host:
            self._check(self._lib.rip_apply_batch_host(self._h, in_ptr, ins, n, rows, cols, channels, encoding.encode(),
                                                       out_ptr, outs))
        else:
            self._check(self._lib.rip_apply_batch_device(self._h, in_ptr, ins, n, rows, cols, channels,
                                                         encoding.encode(), out_ptr, outs, dist_color_ptr or None,
                                                         stream or None))

    # ---- loaders -----------------------------------------------------------------------------
    def load_params(self, file_path: str): self._check(self._lib.rip_load_params(self._h, file_path.encode()))
    def load_camera_calibration(self, file_path: str): self._check(self._lib.rip_load_camera_calibration(self._h, file_path.encode()))
    def load_color_calibration(self, file_path: str): self._check(self._lib.rip_load_color_calibration(self._h, file_path.encode()))
    def init_undistortion(self): self._check(self._lib.rip_init_undistortion(self._h))
    def reset_white_balance_temporal_consistency(self): self._check(self._lib.rip_reset_white_balance_temporal_consistency(self._h))

    # ---- setters (names = raw_image_pipeline_python.cpp:25-57) ---------------------------------
    def set_gpu(self, use_gpu): self._set_bool("gpu", use_gpu)
    def set_debug(self, debug): self._set_bool("debug", debug)
    def set_debayer(self, enabled): self._set_bool("debayer/enabled", enabled)
    def set_debayer_encoding(self, encoding): self._set_string("debayer/encoding", encoding)
    def _reads_16bit(self, encoding: str) -> bool:
        """16-bit samples are read only with the opt-in extension on; otherwise the 16-bit Bayer names throw the
        reference's "valid pattern but is not supported" before any pixel is touched (debayer.cpp:76-78)."""
        return _is_16bit(encoding) and self._get_bool("debayer/allow_16bit")

    def set_debayer_allow_16bit(self, enabled):
        """EXTENSION (not in the reference, which throws for bayer_*16): accept 16-bit Bayer frames (uint16 arrays)."""
        self._set_bool("debayer/allow_16bit", enabled)
    def set_flip(self, enabled): self._set_bool("flip/enabled", enabled)
    def set_flip_angle(self, angle): self._set_int("flip/angle", angle)
    def set_white_balance(self, enabled): self._set_bool("white_balance/enabled", enabled)
    def set_white_balance_method(self, method): self._set_string("white_balance/method", method)
    def set_white_balance_percentile(self, percentile): self._set_double("white_balance/clipping_percentile", percentile)
    def set_white_balance_saturation_threshold(self, bright_thr, dark_thr): self._set_doubles("white_balance/saturation_threshold", [bright_thr, dark_thr])
    def set_white_balance_temporal_consistency(self, enabled): self._set_bool("white_balance/temporal_consistency", enabled)
    def set_gamma_correction(self, enabled): self._set_bool("gamma_correction/enabled", enabled)
    def set_gamma_correction_method(self, method): self._set_string("gamma_correction/method", method)
    def set_gamma_correction_k(self, k): self._set_double("gamma_correction/k", k)
    def set_vignetting_correction(self, enabled): self._set_bool("vignetting_correction/enabled", enabled)
    def set_vignetting_correction_parameters(self, scale, a2, a4): self._set_doubles("vignetting_correction/parameters", [scale, a2, a4])
    def set_color_enhancer(self, enabled): self._set_bool("color_enhancer/enabled", enabled)
    def set_color_enhancer_hue_gain(self, gain): self._set_double("color_enhancer/hue_gain", gain)
    def set_color_enhancer_saturation_gain(self, gain): self._set_double("color_enhancer/saturation_gain", gain)
    def set_color_enhancer_value_gain(self, gain): self._set_double("color_enhancer/value_gain", gain)
    def set_color_calibration(self, enabled): self._set_bool("color_calibration/enabled", enabled)
    def set_color_calibration_matrix(self, m): self._set_doubles("color_calibration/matrix", list(m))
    def set_color_calibration_bias(self, b): self._set_doubles("color_calibration/bias", list(b))
    def set_undistortion(self, enabled): self._set_bool("undistortion/enabled", enabled)
    def set_undistortion_image_size(self, width, height): self._set_doubles("undistortion/image_size", [width, height])
    def set_undistortion_new_image_size(self, width, height): self._set_doubles("undistortion/new_image_size", [width, height])
    def set_undistortion_balance(self, balance): self._set_double("undistortion/balance", balance)
    def set_undistortion_fov_scale(self, fov_scale): self._set_double("undistortion/fov_scale", fov_scale)
    def set_undistortion_camera_matrix(self, m): self._set_doubles("undistortion/camera_matrix", list(m))
    def set_undistortion_distortion_coeffs(self, c): self._set_doubles("undistortion/distortion_coefficients", list(c))
    def set_undistortion_distortion_model(self, model): self._set_string("undistortion/distortion_model", model)
    def set_undistortion_rectification_matrix(self, m): self._set_doubles("undistortion/rectification_matrix", list(m))
    def set_undistortion_projection_matrix(self, m): self._set_doubles("undistortion/projection_matrix", list(m))

    # ---- getters (raw_image_pipeline_python.cpp:58-72 + raw_image_pipeline.hpp:109-137) -------
    def is_debayer_enabled(self): return self._get_bool("debayer/enabled")
    def is_flip_enabled(self): return self._get_bool("flip/enabled")
    def is_white_balance_enabled(self): return self._get_bool("white_balance/enabled")
    def is_color_calibration_enabled(self): return self._get_bool("color_calibration/enabled")
    def is_gamma_correction_enabled(self): return self._get_bool("gamma_correction/enabled")
    def is_vignetting_correction_enabled(self): return self._get_bool("vignetting_correction/enabled")
    def is_color_enhancer_enabled(self): return self._get_bool("color_enhancer/enabled")
    def is_undistortion_enabled(self): return self._get_bool("undistortion/enabled")

    def get_color_calibration_matrix(self): return self._get_doubles("color_calibration/matrix", (3, 3)).astype(np.float32)
    def get_color_calibration_bias(self): return self._get_doubles("color_calibration/bias", (4, 1))

    def get_dist_image_height(self): return self._get_int("dist/image_height")
    def get_dist_image_width(self): return self._get_int("dist/image_width")
    def get_dist_distortion_model(self): return self._get_string("dist/distortion_model")
    def get_dist_camera_matrix(self): return self._get_doubles("dist/camera_matrix", (3, 3))
    def get_dist_distortion_coefficients(self): return self._get_doubles("dist/distortion_coefficients", (1, 4))
    def get_dist_rectification_matrix(self): return self._get_doubles("dist/rectification_matrix", (3, 3))
    def get_dist_projection_matrix(self): return self._get_doubles("dist/projection_matrix", (3, 4))
    def get_rect_image_height(self): return self._get_int("rect/image_height")
    def get_rect_image_width(self): return self._get_int("rect/image_width")
    def get_rect_distortion_model(self): return self._get_string("rect/distortion_model")
    def get_rect_camera_matrix(self): return self._get_doubles("rect/camera_matrix", (3, 3))
    def get_rect_distortion_coefficients(self): return self._get_doubles("rect/distortion_coefficients", (1, 4))
    def get_rect_rectification_matrix(self): return self._get_doubles("rect/rectification_matrix", (3, 3))
    def get_rect_projection_matrix(self): return self._get_doubles("rect/projection_matrix", (3, 4))

    def _get_image(self, which: int) -> np.ndarray:
        r, c, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = self._lib.rip_get_image(self._h, which, None, 0, ctypes.byref(r), ctypes.byref(c), ctypes.byref(k))
        if rc not in (L.RIP_OK, L.RIP_ERR_BUFFER_TOO_SMALL):
            self._check(rc)
        if r.value == 0 or c.value == 0:
            return np.empty((0, 0), np.uint8)
        out = np.empty((r.value, c.value, k.value), np.uint8)
        self._check(self._lib.rip_get_image(self._h, which, out.ctypes.data, out.nbytes, ctypes.byref(r), ctypes.byref(c),
                                            ctypes.byref(k)))
        return out[:, :, 0] if k.value == 1 else out

    def get_dist_debayered_image(self): return self._get_image(L.RIP_IMAGE_DIST_DEBAYERED)
    def get_dist_color_image(self): return self._get_image(L.RIP_IMAGE_DIST_COLOR)
    def get_rect_mask(self): return self._get_image(L.RIP_IMAGE_RECT_MASK)
    def get_processed_image(self): return self._get_image(L.RIP_IMAGE_PROCESSED)

    # ---- inspection --------------------------------------------------------------------------
    def kernel_launches(self) -> int: return self._get_int("stats/kernel_launches")
    def ccc_uv(self) -> Tuple[int, int]:
        """(u, v) = arg-max of the CCC response for the last frame processed (single frame or last of a batch)."""
        return self._get_int("stats/ccc_u"), self._get_int("stats/ccc_v")
    def ccc_gains(self) -> np.ndarray: return self._get_doubles("stats/ccc_gains")
    def log(self) -> str: return self._get_string("log")

    def debug_table(self, name: str, rows: int = 0, cols: int = 0) -> bytes:
        n = ctypes.c_size_t()
        rc = self._lib.rip_debug_table(self._h, name.encode(), rows, cols, None, 0, ctypes.byref(n))
        if rc not in (L.RIP_OK, L.RIP_ERR_BUFFER_TOO_SMALL):
            self._check(rc)
        buf = ctypes.create_string_buffer(n.value)
        self._check(self._lib.rip_debug_table(self._h, name.encode(), rows, cols, buf, n.value, ctypes.byref(n)))
        return buf.raw
