import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raw_image_pipeline_b200 import RawImagePipeline, synth
def run(rows, cols, flip, pca, gamma):
    p = RawImagePipeline(False, "", "", "")
    for name in ("white_balance", "color_calibration", "gamma_correction", "vignetting_correction", "color_enhancer", "undistortion", "flip"):
        getattr(p, "set_" + name)(False)
    if pca: p.set_white_balance(True); p.set_white_balance_method("pca")
    if flip: p.set_flip(True); p.set_flip_angle(flip)
    if gamma: p.set_gamma_correction(True); p.set_gamma_correction_k(0.8)
    raw = synth.bayer_frame(rows, cols, "bayer_rggb8", 1, "U")
    try:
        out = p.process(raw, "bayer_rggb8")
        print("OK  ", rows, cols, flip, pca, gamma, out.shape, flush=True)
    except Exception as e:
        print("FAIL", rows, cols, flip, pca, gamma, str(e)[:120], flush=True)
        sys.exit(1)
a = sys.argv[1:]
run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]))
