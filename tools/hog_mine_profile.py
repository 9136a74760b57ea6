"""HOG mining sweep for ncu / timing: a small trained HOG cascade (48x96 windows, BASELINE configs[3] geometry) over every window
of `n` generated 960x720 backgrounds (cc_mine_scan), by the window miner (level gradients + one CTA per window, mode 0) and,
with `slots`, by the literal per-window slot path (mode 2).  Usage: hog_mine_profile.py [n_images] [sweeps] [slots]"""
import ctypes as C
import io
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth, _lib
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
modes = [0, 2] if "slots" in sys.argv[3:] else [0]
W, H = 48, 96
pos = synth.make_positives(700, W, H, seed=12345)
bgs = [synth.make_background(960, 720, 4000 + i, clutter=1.5) for i in range(6)]       # registered round-robin up to n_images entries
ctx = ccb.Context(0)
lib, h = ctx.lib, ctx.h
cc = CvCascadeClassifier(ctx, out=io.StringIO())
cc.train(tempfile.mkdtemp(), "pos.vec", "bg.txt", 500, 1000, 1024, 1024, 5, CvCascadeParams(0, 2, (W, H)), ccb.CvFeatureParams.create(2),
         ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs[:6])
print("cascade: %d stages, weak per stage %s" % (len(cc.stageClassifiers), [len(s.trees) for s in cc.stageClassifiers]), flush=True)
_lib.check(lib.cc_bg_clear(h), "cc_bg_clear")
for i in range(nimg):
    im = np.ascontiguousarray(bgs[i % 6])
    _lib.check(lib.cc_bg_register(h, im.ctypes.data_as(C.c_void_p), im.shape[1], im.shape[0]), "cc_bg_register")
cc._sync_cascade()
win, surv, dig = C.c_int64(0), C.c_int64(0), C.c_uint64(0)
for mode in modes:
    _lib.check(lib.cc_mine_set_mode(h, mode), "cc_mine_set_mode")
    for s in range(sweeps):
        ctx.timing_reset()
        t0 = time.time()
        _lib.check(lib.cc_mine_scan_digest(h, 0, nimg, C.byref(win), C.byref(surv), C.byref(dig)), "cc_mine_scan_digest")
        ctx.synchronize()
        dt = time.time() - t0
        ms, k = ctx.timing("mine")
        ims, ik = ctx.timing("integral")
        print("mode %d sweep %d: %d images, %d windows, %d survivors (digest %016x), %.2f ms wall, %.2f ms mining kernels (%d launches) + %.2f ms K1 -> %.3g windows/s"
              % (mode, s, nimg, win.value, surv.value, dig.value, dt * 1e3, ms, k, ims, win.value / dt), flush=True)
_lib.check(lib.cc_mine_set_mode(h, 0), "cc_mine_set_mode")
