"""Generates tests/golden/cv2_golden.npz with python cv2 (OpenCV 4.13, the only OpenCV in this image).
Run in the build container:  python tests/golden/make_golden.py
The fixtures pin the OpenCV-owned arithmetic the oracle restates (cv::integral incl. tilted, INTER_LINEAR_EXACT
resize, cartToPolar) and the .vec reader against the reference's own fixture (first samples of
traincascade/res/barcode.vec; data, not source)."""
import os
import struct

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261018)
out = {}
for k, (h, w) in enumerate([(24, 24), (32, 75), (13, 9)]):
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    s, q, t = cv2.integral3(img)
    out["int%d_img" % k], out["int%d_sum" % k], out["int%d_sq" % k], out["int%d_tilt" % k] = img, s, q, t
for k, (sw, sh, dw, dh) in enumerate([(1920, 1080, 43, 24), (640, 480, 341, 256), (300, 200, 299, 199), (100, 80, 75, 32), (97, 61, 24, 24)]):
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    if k == 0:
        src = src[:270, :480].copy()      # keep the fixture small: 480x270 -> 43x24
    out["rs%d_src" % k] = src
    out["rs%d_dst" % k] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
dx = rng.integers(-255, 256, 4096).astype(np.float32)
dy = rng.integers(-255, 256, 4096).astype(np.float32)
mag, ang = cv2.cartToPolar(dx, dy, angleInDegrees=False)
out["c2p_dx"], out["c2p_dy"], out["c2p_mag"], out["c2p_ang"] = dx, dy, mag[:, 0], ang[:, 0]
ref_vec = "/root/reference/traincascade/res/barcode.vec"
if os.path.exists(ref_vec):
    raw = open(ref_vec, "rb").read()
    count, vsz, _, _ = struct.unpack("<iihh", raw[:12])
    rec = 1 + 2 * vsz
    out["vec_header"] = np.array([count, vsz, len(raw)], np.int64)
    out["vec_first3_raw"] = np.frombuffer(raw[12:12 + 3 * rec], np.uint8).copy()
np.savez_compressed(os.path.join(HERE, "cv2_golden.npz"), **out)
print("wrote", os.path.join(HERE, "cv2_golden.npz"), {k: v.shape for k, v in out.items()})
