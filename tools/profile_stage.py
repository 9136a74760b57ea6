"""Short cfg-3-shaped stage for ncu: one setImages + one 2-weak stage (precalc, sort, 2 split searches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cascadeclassifier_b200 as ccb
from conftest import make_samples
npos, nneg = int(sys.argv[1]), int(sys.argv[2])
imgs, labels = make_samples(npos, nneg)
ctx = ccb.Context(0)
ev = ccb.CvFeatureEvaluator.create(0, ctx)
ev.init(ccb.CvHaarFeatureParams(2), len(imgs), (24, 24))
ev.setImages(imgs, labels)
st = ccb.CvCascadeBoost()
st.train(ev, len(imgs), 1024, 1024, ccb.CvCascadeBoostParams(maxFalseAlarm=0.5, maxWeakCount=2))
print("weak", len(st.trees), st.log)
