"""Ad-hoc kernel timing on the GPU box (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cascadeclassifier_b200 as ccb
from conftest import make_samples

npos, nneg = int(sys.argv[1]) if len(sys.argv) > 1 else 5000, int(sys.argv[2]) if len(sys.argv) > 2 else 10000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
maxfa = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
t0 = time.time()
imgs, labels = make_samples(npos, nneg)
print("data gen %.1fs" % (time.time() - t0), flush=True)
ctx = ccb.Context(0)
ev = ccb.CvFeatureEvaluator.create(0, ctx)
fp = ccb.CvHaarFeatureParams(mode)
n = len(imgs)
ev.init(fp, n, (24, 24))
if os.environ.get("CC_SORT_MODE"):
    ctx.set_sort_mode(int(os.environ["CC_SORT_MODE"]))
for rep in range(3):
    ctx.timing_reset()
    st0 = ctx.search_stats() if rep else dict(searches=0, seed_rows=0, flagged_rows=0)
    t0 = time.time()
    ev.setImages(imgs, labels)
    ctx.synchronize()
    t1 = time.time()
    st = ccb.CvCascadeBoost()
    ok = st.train(ev, n, 1024, 1024, ccb.CvCascadeBoostParams(maxFalseAlarm=maxfa, maxWeakCount=100))
    t2 = time.time()
    F = ev.getNumFeatures()
    print("rep %d: setImages %.2f ms, train %.1f ms, weak %d, F=%d N=%d" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(st.trees), F, n))
    for name in ["integral", "precalc", "sort", "split_search", "boost_small"]:
        ms, k = ctx.timing(name)
        extra = ""
        if name == "precalc":
            extra = " -> %.3g evals/s, %.1f%% of 6547.8 GB/s" % (F * n / (ms * 1e-3), F * n * 4 / (ms * 1e-3) / 6547.8e9 * 100)
        if name == "sort":
            extra = " -> %.1f%% of HBM roofline (6 B/elem), %d rows deferred to the radix kernel" % (F * n * 6 / (ms * 1e-3) / 6547.8e9 * 100, ctx.sort_deferred_rows())
        if name == "split_search" and k:
            st1 = ctx.search_stats()
            ns = max(1, st1["searches"] - st0["searches"])
            extra0 = " searches %d flagged/search %.0f seeds/search %.0f" % (ns, (st1["flagged_rows"] - st0["flagged_rows"]) / ns, (st1["seed_rows"] - st0["seed_rows"]) / ns)
            print("  ", extra0)
            per = ms / ns
            extra = " -> %.2f ms/node, %.1f%% of HBM roofline (6 B/elem)" % (per, F * n * 6 / (per * 1e-3) / 6547.8e9 * 100)
        print("   %-13s %9.3f ms  launches %d%s" % (name, ms, k, extra))
    print("   log", st.log[:3], "...", st.log[-1] if st.log else None, "thr", st.threshold)
