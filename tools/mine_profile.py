"""Mining sweep for ncu / timing: the fixture cascade of BASELINE configs[2] over every window of `n` generated 1920x1080
backgrounds (BASELINE configs[4] shape, cc_mine_scan: every window evaluated).  Usage: mine_profile.py [n_images] [sweeps]"""
import ctypes as C
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth, _lib
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 320
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
ctx = ccb.Context(0)
lib, h = ctx.lib, ctx.h
stages, note = bench.load_fixture_cascade()
cc = CvCascadeClassifier(ctx, out=io.StringIO())
cc.cascadeParams = CvCascadeParams(0, 0, (24, 24))
cc.featureEvaluator = ccb.CvFeatureEvaluator.create(ccb.CvFeatureParams.HAAR, ctx)
cc.featureEvaluator.init(ccb.CvHaarFeatureParams(2), 64, (24, 24))
cc.stageClassifiers = stages
block = synth.make_backgrounds_device(nimg, 1920, 1080, 9000, dev)
torch.cuda.synchronize()
for i in range(nimg):
    _lib.check(lib.cc_bg_register_device(h, C.c_void_p(block[i].data_ptr()), 1920, 1080, 1920), "cc_bg_register_device")
ctx.synchronize()
cc._sync_cascade()
win, surv = C.c_int64(0), C.c_int64(0)
for s in range(sweeps):
    ctx.timing_reset()
    t0 = time.time()
    _lib.check(lib.cc_mine_scan(h, 0, nimg, C.byref(win), C.byref(surv)), "cc_mine_scan")
    ctx.synchronize()
    dt = time.time() - t0
    ms, k = ctx.timing("mine")
    print("sweep %d: %d images, %d windows, %d survivors, %.2f ms wall, %.2f ms in mining kernels over %d launches -> %.3g windows/s"
          % (s, nimg, win.value, surv.value, dt * 1e3, ms, k, win.value / dt))
