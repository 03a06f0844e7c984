"""ctypes binding of librip_b200.so (the C ABI declared in include/rip_b200.h).

The product path fails loudly when the CUDA library is missing: there is no Python/CPU
fallback for pixel work anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librip_b200.so")

RIP_OK = 0
RIP_ERR_INVALID_ARGUMENT = 1
RIP_ERR_CUDA = 2
RIP_ERR_IO = 3
RIP_ERR_BUFFER_TOO_SMALL = 4
RIP_ERR_UNKNOWN_KEY = 5
RIP_ERR_UNSUPPORTED = 6

RIP_IMAGE_DIST_DEBAYERED = 0
RIP_IMAGE_DIST_COLOR = 1
RIP_IMAGE_PROCESSED = 2
RIP_IMAGE_RECT_MASK = 3

# every symbol include/rip_b200.h declares: name -> (restype, argtypes)
_H = c_void_p
SIGNATURES = {
    "rip_create": (c_int, [c_int, c_char_p, c_char_p, c_char_p, POINTER(_H)]),
    "rip_create_default": (c_int, [c_int, POINTER(_H)]),
    "rip_destroy": (None, [_H]),
    "rip_last_error": (c_char_p, [_H]),
    "rip_load_params": (c_int, [_H, c_char_p]),
    "rip_load_camera_calibration": (c_int, [_H, c_char_p]),
    "rip_load_color_calibration": (c_int, [_H, c_char_p]),
    "rip_init_undistortion": (c_int, [_H]),
    "rip_reset_white_balance_temporal_consistency": (c_int, [_H]),
    "rip_set_bool": (c_int, [_H, c_char_p, c_int]),
    "rip_set_int": (c_int, [_H, c_char_p, c_int]),
    "rip_set_double": (c_int, [_H, c_char_p, c_double]),
    "rip_set_string": (c_int, [_H, c_char_p, c_char_p]),
    "rip_set_doubles": (c_int, [_H, c_char_p, POINTER(c_double), c_int]),
    "rip_get_bool": (c_int, [_H, c_char_p, POINTER(c_int)]),
    "rip_get_int": (c_int, [_H, c_char_p, POINTER(c_int)]),
    "rip_get_double": (c_int, [_H, c_char_p, POINTER(c_double)]),
    "rip_get_string": (c_int, [_H, c_char_p, c_char_p, c_size_t]),
    "rip_get_doubles": (c_int, [_H, c_char_p, POINTER(c_double), c_int, POINTER(c_int)]),
    "rip_output_shape": (c_int, [_H, c_int, c_int, c_int, c_char_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rip_apply": (c_int, [_H, c_void_p, c_int, c_int, c_int, c_size_t, c_char_p, c_size_t, c_void_p, c_size_t,
                          POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rip_get_image": (c_int, [_H, c_int, c_void_p, c_size_t, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rip_apply_batch_device": (c_int, [_H, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_char_p, c_void_p,
                                       c_size_t, c_void_p, c_void_p]),
    "rip_apply_batch_host": (c_int, [_H, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_char_p, c_void_p,
                                     c_size_t]),
    "rip_pinned_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "rip_pinned_free": (c_int, [c_void_p]),
    "rip_apply_batch_host_multi": (c_int, [POINTER(_H), c_int, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_char_p, c_void_p,
                                           c_size_t]),
    "rip_debug_table": (c_int, [_H, c_char_p, c_int, c_int, c_void_p, c_size_t, POINTER(c_size_t)]),
    "rip_device_count": (c_int, []),
    "rip_set_device": (c_int, [_H, c_int]),
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m raw_image_pipeline_b200.build` "
                "(nvcc, sm_100a). raw_image_pipeline_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
