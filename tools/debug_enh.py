import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cv2
from test_gpu_parity import make_pair
from raw_image_pipeline_b200 import synth
from oracle import cv2_oracle as O
v = np.arange(256, dtype=np.uint8)
a, b, c = np.meshgrid(v, v, v, indexing="ij")
cube = np.ascontiguousarray(np.stack([a, b, c], -1).reshape(4096, 4096, 3))
for kw in (dict(enh=(1.0, 1.2, 1.0)), dict(enh=(1.0, 1.0, 1.0)), dict(gamma=0.8, enh=(1.0,1.2,1.0)), dict(vig=(1.5,1e-3,1e-6), enh=(1.0,1.2,1.0))):
    p, o = make_pair(4096, 4096, **kw)
    ref, _ = o.apply(cube, "bgr8")
    got = p.process(cube, "bgr8")
    bad = (got != ref).any(-1)
    print(kw, "cube mismatching px:", int(bad.sum()))
    if bad.sum():
        idx = np.argwhere(bad)[:8]
        for (y, x) in idx:
            print("  in", cube[y, x], "got", got[y, x], "ref", ref[y, x], "hsv", cv2.cvtColor(cube[y:y+1, x:x+1], cv2.COLOR_BGR2HSV)[0,0])
raw = synth.bayer_frame(540, 720, "bayer_bggr8", 21, "U")
p, o = make_pair(540, 720, enh=(1.0, 1.2, 1.0))
ref, _ = o.apply(raw, "bayer_bggr8", keep_stages=True)
got = p.process(raw, "bayer_bggr8")
bad = (got != ref).any(-1)
print("bayer enh-only mismatching px:", int(bad.sum()), "rows hist", np.bincount(np.argwhere(bad)[:, 0] % 32, minlength=32), "col%128 hist", np.bincount(np.argwhere(bad)[:, 1] % 4, minlength=4))
for (y, x) in np.argwhere(bad)[:8]:
    print("  at", y, x, "in", o.stages["flip"][y, x], "got", got[y, x], "ref", ref[y, x])
img = np.random.default_rng(5).integers(0, 256, (270, 362, 3), dtype=np.uint8)
for kw in (dict(), dict(gamma=0.8), dict(cc=True), dict(wb="pca"), dict(vig=(1.5,1e-3,1e-6)), dict(flip=180)):
    p, o = make_pair(270, 362, **kw)
    ref, _ = o.apply(img, "bgr8"); got = p.process(img, "bgr8")
    bad = (got != ref).any(-1)
    print("colour input", kw, "mismatching px:", int(bad.sum()), np.argwhere(bad)[:5].tolist())
