#!/usr/bin/env python3
"""Small workload for compute-sanitizer (tools/sanitize.sh): every kernel family once, on shapes with partial strips /
tiles, both rotations, checked against the oracle so that a sanitizer-clean run is also a correct run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from raw_image_pipeline_b200 import synth  # noqa: E402
from test_gpu_parity import FULL, make_pair  # noqa: E402


def main():
    rows, cols, enc = 166, 400, "bayer_grbg8"
    frames = synth.bayer_batch(2, rows, cols, enc, 9100, "N")
    bad = 0
    for flip in (0, 180):
        for wb in ("pca", "ccc"):
            kw = dict(FULL); kw["flip"] = flip; kw["wb"] = wb
            p, o = make_pair(rows, cols, **kw)
            refs = [o.apply(frames[i], enc)[0] for i in range(2)]
            for fused in (1, 2):                       # tile kernel, strip kernel
                p._set_int("debug/fused_kernel", fused)
                out = p.process_batch(frames, enc)     # 4-byte intermediate + tile undistortion kernel
                bad += int(np.count_nonzero(out[0] != refs[0])) + int(np.count_nonzero(out[1] != refs[1]))
            p._set_int("debug/fused_kernel", 0)
            p._set_bool("debug/force_gather_remap", True)
            bad += int(np.count_nonzero(p.process_batch(frames, enc)[1] != refs[1]))
            p._set_bool("debug/force_generic_kernels", True)
            bad += int(np.count_nonzero(p.process_batch(frames, enc)[1] != refs[1]))
            p._set_bool("debug/force_generic_kernels", False); p._set_bool("debug/force_gather_remap", False)
            for i in range(3):                          # apply(): staging + CUDA-graph replay
                bad += int(np.count_nonzero(p.process(frames[i % 2], enc) != refs[i % 2]))
            p._set_bool("undistortion/rect_mask", True)
            p.process(frames[0], enc)
            assert p.get_rect_mask().shape == (rows, cols)
    # light stage sets (strip kernel, BGR8 staging + TMA stores), ragged height
    for kw in (dict(gamma=0.8), dict(wb="pca", gamma=0.8), dict(flip=180), dict(wb="ccc", gamma=0.8, flip=180), dict(wb="pca", cc=True, gamma=0.8)):
        p, o = make_pair(163, 400, **kw)
        raw = synth.bayer_frame(163, 400, enc, 9200, "U")
        bad += int(np.count_nonzero(p.process_batch(raw[None], enc)[0] != o.apply(raw, enc)[0]))
    # 16-bit Bayer extension (k_bayer16_to_bgr8 + the 3-channel chain)
    p, o = make_pair(166, 400, **FULL)
    p.set_debayer_allow_16bit(True); o.p.debayer_allow_16bit = True
    raw16 = (np.random.default_rng(9300).integers(0, 4096, (166, 400), dtype=np.uint16) << 4)
    bad += int(np.count_nonzero(p.process(raw16, "bayer_rggb16") != o.apply(raw16, "bayer_rggb16")[0]))
    print("sanitize workload: values differing from the oracle =", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
