"""Mining throughput probe (BASELINE configs[4] shape): cascade evaluation over every window of generated
1920x1080 backgrounds (13584 windows / image at round 0).  Usage: mine_bench.py [n_images] [n_stages] [mode]"""
import sys, os, time, io, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth, _lib
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nstages = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t0 = time.time()
pos = synth.make_positives(1400, 24, 24)
small = synth.make_backgrounds(12, 480, 360, seed0=500, clutter=2.5)
ctx = ccb.Context(0)
cc = CvCascadeClassifier(ctx, out=io.StringIO())
import tempfile
ok = cc.train(tempfile.mkdtemp(prefix="mine_bench_cascade_"), "pos.vec", "bg.txt", 1000, 2000, 1024, 1024, nstages, CvCascadeParams(0, 0, (24, 24)),
              ccb.CvHaarFeatureParams(mode), ccb.CvCascadeBoostParams(), positives=pos, backgrounds=small)
print("trained %d stages in %.1fs: weak per stage %s" % (len(cc.stageClassifiers), time.time() - t0, [len(s.trees) for s in cc.stageClassifiers]))
for t in cc.timings:
    print("   stage %d: mine %.3fs (%d windows consumed) train %.3fs weak %d" % (t["stage"], t["mine_s"], t["neg_consumed"], t["train_s"], t["weak"]))
big = [synth.make_background(1920, 1080, 9000 + i, clutter=0.3) for i in range(min(nimg, 8))]
lib, h = ctx.lib, ctx.h
_lib.check(lib.cc_bg_clear(h))
for i in range(nimg):
    b = big[i % len(big)]
    _lib.check(lib.cc_bg_register(h, b.ctypes.data_as(C.c_void_p), b.shape[1], b.shape[0]))
cc._sync_cascade()
win, surv = C.c_int64(0), C.c_int64(0)
_lib.check(lib.cc_mine_scan(h, 0, nimg, C.byref(win), C.byref(surv)))      # warm-up (same size: allocations)
ctx.timing_reset()
ctx.synchronize()
t0 = time.time()
_lib.check(lib.cc_mine_scan(h, 0, nimg, C.byref(win), C.byref(surv)))
ctx.synchronize()
dt = time.time() - t0
print("scan: %d images, %d windows, %d survivors, %.3f s -> %.3g windows/s" % (nimg, win.value, surv.value, dt, win.value / dt))
for name in ["mine", "integral"]:
    ms, k = ctx.timing(name)
    print("   %-10s %9.3f ms  launches %d" % (name, ms, k))
