"""Race / memory check driver for the mining kernels (and, optionally, one tiny stage training).

  python tools/mine_race.py train  gpurun_out/race_cascade.pkl         # no sanitizer: trains a small Haar ALL cascade, saves its stages
  compute-sanitizer --tool racecheck python tools/mine_race.py scan gpurun_out/race_cascade.pkl [nimg] [w] [h]
  compute-sanitizer --tool racecheck python tools/mine_race.py stage   # one tiny stage training (12x12 windows) under the sanitizer

`scan` evaluates the saved cascade over every window of `nimg` generated backgrounds with every miner (mode 0 fused pyramid
builder, 1 separate kernels, 3 fused tile miner, 2 per-window slots) and requires identical survivor counts."""
import ctypes as C
import io
import os
import pickle
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth, _lib
from cascadeclassifier_b200.boost import CvCascadeBoost
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

what = sys.argv[1] if len(sys.argv) > 1 else "scan"
ctx = ccb.Context(0)
if what == "train":
    pos = synth.make_positives(500, 24, 24, seed=3)
    small = synth.make_backgrounds(8, 480, 360, seed0=900, clutter=2.0)
    cc = CvCascadeClassifier(ctx, out=io.StringIO())
    cc.train(tempfile.mkdtemp(prefix="race_"), "pos.vec", "bg.txt", 350, 600, 1024, 1024, 4, CvCascadeParams(0, 0, (24, 24)),
             ccb.CvHaarFeatureParams(2), ccb.CvCascadeBoostParams(), positives=pos, backgrounds=small)
    with open(sys.argv[2], "wb") as f:
        pickle.dump([(s.threshold, s.trees) for s in cc.stageClassifiers], f)
    print("trained %d stages, weak %s" % (len(cc.stageClassifiers), [len(s.trees) for s in cc.stageClassifiers]))
elif what == "scan":
    with open(sys.argv[2], "rb") as f:
        stages = [CvCascadeBoost.from_trees(t, tr) for t, tr in pickle.load(f)]
    nimg = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    w = int(sys.argv[4]) if len(sys.argv) > 4 else 1920
    h = int(sys.argv[5]) if len(sys.argv) > 5 else 1080
    ev = ccb.CvFeatureEvaluator.create(ccb.CvFeatureParams.HAAR, ctx)
    ev.init(ccb.CvHaarFeatureParams(2), 64, (24, 24))
    lib, hd = ctx.lib, ctx.h
    for i in range(nimg):
        b = synth.make_background(w, h, 7000 + i, clutter=1.0)
        _lib.check(lib.cc_bg_register(hd, b.ctypes.data_as(C.c_void_p), b.shape[1], b.shape[0]))
    arr = (_lib.Stage * len(stages))()
    for i, s in enumerate(stages):
        arr[i] = s.c_stage()
    _lib.check(lib.cc_cascade_set_stages(hd, arr, len(stages)))
    res = {}
    for mode in (0, 1, 3, 2):
        _lib.check(lib.cc_mine_set_mode(hd, mode))
        win, surv = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.cc_mine_scan(hd, 0, nimg, C.byref(win), C.byref(surv)))
        res[mode] = (win.value, surv.value)
        print("mode %d: %d windows, %d survivors" % (mode, win.value, surv.value), flush=True)
    assert len(set(res.values())) == 1, res
    print("all miners agree")
else:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from conftest import make_samples
    imgs, labels = make_samples(150, 250, 12, 12)
    ev = ccb.CvFeatureEvaluator.create(ccb.CvFeatureParams.HAAR, ctx)
    ev.init(ccb.CvHaarFeatureParams(2), len(imgs), (12, 12))
    for sort_mode in (0, 1):
        ev.setImages(imgs, labels)          # also invalidates the previous precalc
        _lib.check(ctx.lib.cc_boost_set_sort_mode(ctx.h, sort_mode))
        st = CvCascadeBoost()
        st.train(ev, len(imgs), 1024, 1024, ccb.CvCascadeBoostParams(maxFalseAlarm=0.2, maxWeakCount=3))
        print("sort mode %d: %d weak trees, first split on %d" % (sort_mode, len(st.trees), st.trees[0][0]["var_idx"]), flush=True)
