#!/usr/bin/env python3
"""Converts the reference's CCC/FFCC model (raw_image_pipeline_white_balance/model/default.bin:
int32 width, int32 height, width*height fp32 filter, width*height fp32 bias) into this
library's layout: magic "RIPCCC1\\0", int32 width, int32 height, then filter and bias
*already transposed* (the reference transposes both right after reading them,
convolutional_color_constancy.cpp:131-132), so the loader needs no transpose.

Usage: python tools/convert_ccc_model.py <default.bin> raw_image_pipeline_b200/config/ccc_model.bin
"""
import struct
import sys

import numpy as np

src, dst = sys.argv[1], sys.argv[2]
d = open(src, "rb").read()
w, h = struct.unpack("ii", d[:8])
a = np.frombuffer(d[8:8 + 8 * w * h], dtype=np.float32)
filt = np.ascontiguousarray(a[:w * h].reshape(h, w).T)
bias = np.ascontiguousarray(a[w * h:].reshape(h, w).T)
with open(dst, "wb") as f:
    f.write(b"RIPCCC1\0")
    f.write(struct.pack("ii", filt.shape[1], filt.shape[0]))
    f.write(filt.tobytes())
    f.write(bias.tobytes())
print("wrote", dst, filt.shape)
