"""GPU parity: K4 stage training (CvCascadeBoost::train) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import orc
from conftest import make_samples

pytestmark = pytest.mark.gpu


def prove_tie(ev, labels, weights, mask, g_root, r_root, rel=2e-6):
    """A divergence between the GPU and the oracle is accepted only as a float-level tie that is re-derived here, independently
    of both: the squared-error criterion of find_split_ord_reg (src/o_cvboostree.cpp:361-426) is evaluated in numpy for the two
    chosen (variable, threshold) pairs on the boosting state in force before the tree (oracle snapshot), and must agree to `rel`."""
    act = np.flatnonzero(mask)
    y = 2.0 * labels[act].astype(np.float64) - 1.0
    w = weights[act]
    quals = []
    for root in (g_root, r_root):
        f = ev.values([root["var_idx"]], 0, len(labels))[0][act]
        left = f <= np.float32(root["c"])
        L, R = w[left].sum(), w[~left].sum()
        ls, rs = (w * y)[left].sum(), (w * y)[~left].sum()
        quals.append((ls * ls * R + rs * rs * L) / (L * R))
    assert abs(quals[0] - quals[1]) <= rel * abs(quals[1]), "not a tie: GPU %s -> %.12g, oracle %s -> %.12g" % (g_root, quals[0], r_root, quals[1])


def compare_stage(gpu_stage, ref, tie=None):
    """Trees must agree in (feature, threshold, leaf values, structure) -- all of them.  The one tolerated divergence (north_star:
    "wherever split criteria are not tied within float tolerance") has to be PROVEN by the caller-supplied `tie(t, g_root, r_root)`
    check; without it any difference fails.  Returns the number of identical trees."""
    n = min(len(gpu_stage.trees), len(ref.trees))
    for t in range(n):
        g, r = gpu_stage.trees[t], ref.trees[t]
        if g[0]["var_idx"] != r[0]["var_idx"] or g[0]["c"] != r[0]["c"]:
            assert tie is not None, "tree %d differs: %s vs %s" % (t, g[0], r[0])
            tie(t, g[0], r[0])
            return t
        assert len(g) == len(r)
        for a, b in zip(g, r):
            assert a["left"] == b["left"] and a["right"] == b["right"]
            if a["left"] >= 0:
                assert a["var_idx"] == b["var_idx"]
                assert a["c"] == b["c"] and a["subset"] == b["subset"]
                assert a["quality"] == pytest.approx(b["quality"], rel=1e-5)
            assert a["value"] == pytest.approx(b["value"], rel=1e-9, abs=1e-12)
    assert len(gpu_stage.trees) == len(ref.trees)
    return n


def assert_full_stage(stage, ref):
    """every tree, the stage threshold and every (N, HR, FA) row of the stdout table"""
    assert compare_stage(stage, ref) == len(ref.trees)
    assert stage.threshold == pytest.approx(ref.threshold, rel=1e-6, abs=1e-7)
    assert [r[0] for r in stage.log] == [r[0] for r in ref.log]
    for a, b in zip(stage.log, ref.log):
        assert a[1] == pytest.approx(b[1]) and a[2] == pytest.approx(b[2])


def run_pair(gpu_ctx, ftype, mode, W, H, npos, nneg, **bp):
    import cascadeclassifier_b200 as ccb
    imgs, labels = make_samples(npos, nneg, W, H)
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(ftype, gpu_ctx)
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    ev.init(fp, n, (W, H))
    ev.setImages(imgs, labels)
    params = ccb.CvCascadeBoostParams(boostType=bp.get("boost_type", 3), minHitRate=bp.get("min_hit_rate", 0.995),
                                      maxFalseAlarm=bp.get("max_false_alarm", 0.5),
                                      weightTrimRate=bp.get("weight_trim_rate", 0.95), maxDepth=bp.get("max_depth", 1),
                                      maxWeakCount=bp.get("weak_count", 100))
    stage = ccb.CvCascadeBoost()
    ok = stage.train(ev, n, 1024, 1024, params)
    o = orc.Evaluator(ftype, mode, W, H, n)
    o.set_images(imgs, labels)
    td = orc.TrainData(o, n, orc.SORT_STABLE, 8)
    ref = orc.boost_train(td, orc.boost_params(**bp), cat_sort=orc.SORT_STABLE)
    return ok, stage, ref, ev


@pytest.mark.parametrize("ftype,mode", [(orc.HAAR, orc.BASIC), (orc.HAAR, orc.ALL), (orc.LBP, 0)])
def test_gab_stump_stage_matches_oracle(gpu_ctx, ftype, mode):
    ok, stage, ref, ev = run_pair(gpu_ctx, ftype, mode, 24, 24, 200, 400, max_false_alarm=0.05, weak_count=12)
    assert ok == ref.ok
    assert_full_stage(stage, ref)


def test_hog_stage_matches_oracle(gpu_ctx):
    ok, stage, ref, ev = run_pair(gpu_ctx, orc.HOG, 0, 48, 96, 60, 120, max_false_alarm=0.05, weak_count=6)
    assert ok == ref.ok
    assert_full_stage(stage, ref)


@pytest.mark.parametrize("bt", [orc.DAB, orc.RAB, orc.LB])
def test_other_boost_types(gpu_ctx, bt):
    ok, stage, ref, ev = run_pair(gpu_ctx, orc.HAAR, orc.BASIC, 24, 24, 120, 240, boost_type=bt, max_false_alarm=0.05,
                                  weak_count=6)
    assert ok == ref.ok
    assert_full_stage(stage, ref)


def test_depth2_trees(gpu_ctx):
    ok, stage, ref, ev = run_pair(gpu_ctx, orc.HAAR, orc.BASIC, 24, 24, 150, 300, max_depth=2, max_false_alarm=0.05,
                                  weak_count=5)
    assert ok == ref.ok
    assert_full_stage(stage, ref)


def test_stage_predict_matches_training_decision(gpu_ctx):
    ok, stage, ref, ev = run_pair(gpu_ctx, orc.HAAR, orc.BASIC, 24, 24, 100, 200, max_false_alarm=0.2, weak_count=8)
    n = 300
    sums = stage.predict(ev, 0, returnSum=True, count=n)
    dec = stage.predict(ev, 0, returnSum=False, count=n)
    thr = np.float32(stage.threshold) - np.float32(0.00001)
    # the decision is taken on the f64 stage sum (src/boost.cpp:461-477); the returned f32 sum can only disagree at the threshold itself
    bad = (sums >= thr) != (dec == 1.0)
    assert (np.abs(sums[bad] - thr) <= 4e-7 * max(1.0, abs(float(thr)))).all(), (sums[bad], thr)
    assert bad.sum() <= 2
    pos = sums[:100]
    assert (dec[:100] == 1).mean() >= 0.99
    assert np.isfinite(pos).all()


def test_feature_sharded_search_single_gpu(gpu_ctx):
    """Feature sharding without NCCL: train the first weak tree on each of 2 slices (one context each, run one after
    the other on the same GPU), merge with the library's rule -> must equal the unsharded first tree."""
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200._lib import check
    from cascadeclassifier_b200.dist import select_winner, shard_range
    imgs, labels = make_samples(150, 250)
    n = len(imgs)
    params = ccb.CvCascadeBoostParams(maxWeakCount=1)
    full = None
    cands = []
    for rank, world in [(0, 1), (0, 2), (1, 2)]:
        ctx = ccb.Context(0)
        ev = ccb.CvFeatureEvaluator.create(0, ctx)
        ev.init(ccb.CvHaarFeatureParams(2), n, (24, 24))
        check(ctx.lib.cc_boost_set_shard(ctx.h, rank, world), "set_shard")
        ev.setImages(imgs, labels)
        st = ccb.CvCascadeBoost()
        assert st.train(ev, n, 1024, 1024, params)
        root = st.trees[0][0]
        if world == 1:
            full = root
        else:
            lo, hi = shard_range(ev.getNumFeatures(), rank, world)
            assert lo <= root["var_idx"] < hi
            cands.append(dict(valid=True, quality=root["quality"], var_idx=root["var_idx"], c=root["c"]))
        ctx.close()
    win = select_winner(cands)
    assert win["var_idx"] == full["var_idx"] and win["c"] == full["c"] and win["quality"] == full["quality"]


@pytest.mark.parametrize("bt,depth", [(orc.GAB, 1), (orc.GAB, 2), (orc.RAB, 1), (orc.DAB, 1)])
def test_pruned_search_equals_exhaustive(gpu_ctx, bt, depth):
    """The bound pass (k_ss_bound) may only drop variables that cannot win: the pruned search must grow exactly the
    trees of the exhaustive search (bitwise: same feature, threshold, quality, leaf values), and it must actually prune."""
    import cascadeclassifier_b200 as ccb
    imgs, labels = make_samples(300, 500)
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(orc.HAAR, gpu_ctx)
    ev.init(ccb.CvHaarFeatureParams(orc.ALL), n, (24, 24))
    ev.setImages(imgs, labels)
    params = ccb.CvCascadeBoostParams(boostType=bt, maxFalseAlarm=0.02, maxDepth=depth, maxWeakCount=10)
    stages = []
    for mode in (1, 0):
        gpu_ctx.set_search_mode(mode)
        st = ccb.CvCascadeBoost()
        assert st.train(ev, n, 1024, 1024, params)
        stages.append(st)
    gpu_ctx.set_search_mode(0)
    ex, pr = stages
    assert len(ex.trees) == len(pr.trees) and ex.threshold == pr.threshold
    for a, b in zip(ex.trees, pr.trees):
        assert a == b
    stats = gpu_ctx.search_stats()
    assert stats["searches"] >= len(pr.trees)
    F = ev.getNumFeatures()
    # far fewer rows searched exactly than an exhaustive search would touch
    assert stats["flagged_rows"] + stats["seed_rows"] < 0.25 * F * stats["searches"], stats


_MATRIX = [
    # ftype, mode, W, H, npos, nneg, boost parameters                                                what the case adds
    (orc.LBP, 0, 24, 24, 150, 250, dict(boost_type=orc.DAB, weak_count=5, max_false_alarm=0.05)),   # categorical x classification (misclass)
    (orc.LBP, 0, 24, 24, 150, 250, dict(boost_type=orc.RAB, weak_count=5, max_false_alarm=0.05)),   # categorical x classification (Gini)
    (orc.LBP, 0, 24, 24, 150, 250, dict(boost_type=orc.LB, weak_count=4, max_false_alarm=0.05)),    # categorical x LogitBoost targets
    (orc.LBP, 0, 32, 20, 120, 200, dict(max_depth=2, weak_count=4, max_false_alarm=0.05)),          # categorical, depth 2, non-square window
    (orc.HAAR, orc.CORE, 20, 20, 120, 260, dict(max_depth=3, weak_count=3, max_false_alarm=0.05)),  # depth 3
    (orc.HAAR, orc.BASIC, 24, 24, 130, 270, dict(weight_trim_rate=1.0, weak_count=6, max_false_alarm=0.02)),   # trimming off
    (orc.HAAR, orc.BASIC, 24, 24, 130, 270, dict(weight_trim_rate=0.5, weak_count=6, max_false_alarm=0.02)),   # aggressive trimming
    (orc.HAAR, orc.ALL, 24, 24, 130, 270, dict(min_hit_rate=0.9, max_false_alarm=0.3, weak_count=8)),          # loose hit rate: threshold index > 0
    (orc.HAAR, orc.BASIC, 64, 48, 40, 70, dict(weak_count=3, max_false_alarm=0.05)),                # big window: tiles beyond shared memory
    (orc.HOG, 0, 64, 64, 50, 90, dict(boost_type=orc.RAB, weak_count=4, max_false_alarm=0.05)),     # HOG x classification, square window
    (orc.HAAR, orc.BASIC, 24, 24, 9, 14, dict(weak_count=3)),                                       # fewer samples than one 32-sample tile
]


@pytest.mark.parametrize("case", range(len(_MATRIX)))
def test_stage_configuration_matrix(gpu_ctx, case):
    """Flag combinations the CLI accepts (-bt, -maxDepth, -weightTrimRate, -minHitRate, -w / -h, feature types) that the
    BASELINE configs do not exercise: one stage each, trees / thresholds / log rows against the oracle (trees are compared
    in full: every tree, threshold and log row)."""
    ftype, mode, W, H, npos, nneg, bp = _MATRIX[case]
    ok, stage, ref, ev = run_pair(gpu_ctx, ftype, mode, W, H, npos, nneg, **bp)
    assert ok == ref.ok
    if not ref.ok:
        return
    assert_full_stage(stage, ref)


@pytest.mark.parametrize("bt", [orc.GAB, orc.DAB, orc.RAB, orc.LB])
@pytest.mark.parametrize("ftype", [orc.HAAR, orc.LBP])
def test_weights_and_mask_after_training_match_oracle(gpu_ctx, bt, ftype):
    """update_weights (src/boost.cpp:168-407) + trim_weights (src/o_cvboost.cpp:101-139), checked directly: the AdaBoost
    weights and the subsample mask the library holds after the last weak tree against the oracle's snapshot after the same
    tree.  Weights: f64, relative 1e-12 (cvExp / reassociated sums); mask: identical."""
    import ctypes as C
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200._lib import check
    k = 4
    imgs, labels = make_samples(140, 260)
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(ftype, gpu_ctx)
    fp = ccb.CvFeatureParams.create(ftype)
    ev.init(fp, n, (24, 24))
    ev.setImages(imgs, labels)
    stage = ccb.CvCascadeBoost()
    # maxFalseAlarm ~ 0: the stage cannot stop before weak_count trees unless the training set is separated
    assert stage.train(ev, n, 1024, 1024, ccb.CvCascadeBoostParams(boostType=bt, maxFalseAlarm=1e-6, minHitRate=0.999999, maxWeakCount=k))
    w = np.zeros(n, np.float64)
    m = np.zeros(n, np.uint8)
    check(gpu_ctx.lib.cc_boost_get_weights(gpu_ctx.h, w.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p)), "cc_boost_get_weights")
    o = orc.Evaluator(ftype, 0, 24, 24, n)
    o.set_images(imgs, labels)
    td = orc.TrainData(o, n, orc.SORT_STABLE, 8)
    ref = orc.boost_train(td, orc.boost_params(boost_type=bt, weak_count=k, max_false_alarm=1e-6, min_hit_rate=0.999999), cat_sort=orc.SORT_STABLE, record=True)
    assert compare_stage(stage, ref) == len(ref.trees)
    ow, om = ref.snapshot(len(ref.trees) - 1)
    assert abs(w.sum() - 1.0) < 1e-9 and abs(ow.sum() - 1.0) < 1e-9
    np.testing.assert_allclose(w, ow, rtol=1e-12, atol=1e-300)
    assert ((m != 0) == (om != 0)).all()
    assert 0 < int((m != 0).sum()) <= n


def test_gpu_stumps_match_cv2_boost(gpu_ctx):
    """Independent pin of the CUDA trainer itself (not through the oracle): cv2.ml.Boost GENTLE stumps on the fval matrix the
    library computed -- variable, f32 threshold, quality, leaf values of every weak tree (tests/cv2_boost_ref.py)."""
    import cascadeclassifier_b200 as ccb
    from cv2_boost_ref import assert_stumps_match, cv2_boost_stumps
    imgs, labels = make_samples(120, 200, 12, 12)
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(orc.HAAR, gpu_ctx)
    ev.init(ccb.CvHaarFeatureParams(orc.ALL), n, (12, 12))
    ev.setImages(imgs, labels)
    F = ev.getNumFeatures()
    X = np.concatenate([ev.values(list(range(a, min(F, a + 4096))), 0, n) for a in range(0, F, 4096)]).T.copy()
    for bt, name in ((orc.GAB, "GAB"), (orc.LB, "LB"), (orc.RAB, "RAB")):
        stage = ccb.CvCascadeBoost()
        assert stage.train(ev, n, 1024, 1024, ccb.CvCascadeBoostParams(boostType=bt, maxFalseAlarm=1e-6, minHitRate=0.999999, maxWeakCount=8))
        k = assert_stumps_match(stage.trees, cv2_boost_stumps(X, labels, name, 8, 0.95), 3, check_quality=(name != "RAB"), what="gpu " + name)
        assert k >= 3


@pytest.mark.parametrize("ftype,mode,bt", [(orc.HAAR, orc.ALL, orc.GAB), (orc.HAAR, orc.BASIC, orc.DAB), (orc.HAAR, orc.CORE, orc.RAB),
                                           (orc.HAAR, orc.BASIC, orc.LB), (orc.LBP, 0, orc.GAB), (orc.LBP, 0, orc.RAB)])
def test_stump_path_equals_general_path(gpu_ctx, ftype, mode, bt):
    """cc_boost_set_tree_path: stumps grown with one host read per tree (speculative launches, tree assembled on the device) must be
    the trees of the general node-by-node path, bit for bit -- trees, leaf values, threshold, (N, HR, FA) rows, final weights and mask."""
    import ctypes as C
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200._lib import check
    imgs, labels = make_samples(180, 320)
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(ftype, gpu_ctx)
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    ev.init(fp, n, (24, 24))
    ev.setImages(imgs, labels)
    params = ccb.CvCascadeBoostParams(boostType=bt, maxFalseAlarm=0.02, maxWeakCount=9)
    runs = []
    try:
        for path in (1, 0):
            check(gpu_ctx.lib.cc_boost_set_tree_path(gpu_ctx.h, path), "cc_boost_set_tree_path")
            st = ccb.CvCascadeBoost()
            assert st.train(ev, n, 1024, 1024, params)
            w = np.zeros(n, np.float64)
            m = np.zeros(n, np.uint8)
            check(gpu_ctx.lib.cc_boost_get_weights(gpu_ctx.h, w.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p)), "cc_boost_get_weights")
            runs.append((st.trees, st.threshold, st.log, w, m))
    finally:
        check(gpu_ctx.lib.cc_boost_set_tree_path(gpu_ctx.h, 0), "cc_boost_set_tree_path")
    a, b = runs
    assert len(a[0]) >= 2
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])


# ---- wide path: numSamples beyond one CTA's shared memory (a9: CvCascadeBoostTrainData::setData takes any N) ---------------------
def _train_gpu(gpu_ctx, ftype, mode, W, H, imgs, labels, **kw):
    import cascadeclassifier_b200 as ccb
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(ftype, gpu_ctx)
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    ev.init(fp, n, (W, H))
    ev.setImages(imgs, labels)
    st = ccb.CvCascadeBoost()
    ok = st.train(ev, n, 1024, 1024, ccb.CvCascadeBoostParams(**kw))
    return ok, st, ev


@pytest.mark.parametrize("ftype,mode,bt,depth", [(orc.HAAR, orc.ALL, orc.GAB, 1), (orc.HAAR, orc.BASIC, orc.RAB, 1), (orc.HAAR, orc.BASIC, orc.DAB, 2),
                                                 (orc.HAAR, orc.CORE, orc.LB, 1), (orc.HOG, 0, orc.GAB, 1), (orc.LBP, 0, orc.GAB, 2)])
def test_wide_kernels_equal_shared_memory_kernels(gpu_ctx, ftype, mode, bt, depth):
    """cc_boost_set_search_mode(ctx, 2) forces the kernels that stages with more than 15360 samples take (global-memory radix sort
    into 32-bit index rows, operands and values gathered from global memory, categorical operands from global memory) on a small
    stage: sorted rows and trees must be bit-identical to the shared-memory kernels'."""
    import ctypes as C
    from cascadeclassifier_b200._lib import check
    W, H = (32, 32) if ftype == orc.HOG else (16, 16)
    imgs, labels = make_samples(170, 290, W, H)
    n = len(imgs)
    runs = []
    try:
        for sm in (1, 2):
            gpu_ctx.set_search_mode(sm)
            ok, st, ev = _train_gpu(gpu_ctx, ftype, mode, W, H, imgs, labels, boostType=bt, maxDepth=depth, maxFalseAlarm=0.05, maxWeakCount=6)
            assert ok
            rows = None
            if ftype != orc.LBP:
                nr = min(64, ev.getNumFeatures() * ev.getFeatureSize())
                rows = np.zeros((nr, n), np.uint32)
                check(gpu_ctx.lib.cc_boost_get_sorted_idx32(gpu_ctx.h, 0, nr, rows.ctypes.data_as(C.c_void_p)), "cc_boost_get_sorted_idx32")
                vals = np.zeros((nr, n), np.float32)
                check(gpu_ctx.lib.cc_boost_get_sorted_val(gpu_ctx.h, 0, nr, vals.ctypes.data_as(C.c_void_p)), "cc_boost_get_sorted_val")
                rows = (rows, vals)
            runs.append((st.trees, st.threshold, st.log, rows))
    finally:
        gpu_ctx.set_search_mode(0)
    a, b = runs
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    if a[3] is not None:
        assert np.array_equal(a[3][0], b[3][0]) and np.array_equal(a[3][1].view(np.uint32), b[3][1].view(np.uint32))


@pytest.mark.parametrize("ftype,mode,W,npos,nneg,bt", [(orc.HAAR, orc.ALL, 12, 6000, 14000, orc.GAB),       # 20000 > 15360: u16 range, wide kernels
                                                       (orc.HAAR, orc.BASIC, 10, 16000, 50000, orc.GAB),    # 66000 >= 65536: is_buf_16u == false in the reference
                                                       (orc.HAAR, orc.BASIC, 10, 9000, 21000, orc.RAB),
                                                       (orc.LBP, 0, 12, 9000, 21000, orc.GAB)])              # categorical operands no longer fit in shared memory
def test_large_sample_counts_match_oracle(gpu_ctx, ftype, mode, W, npos, nneg, bt):
    """CvCascadeBoostTrainData::setData accepts any sample count (32-bit index buffers from 65536 on,
    src/o_cvcascadeboosttraindata.cpp:250-251,283-309): stages with 20000 / 30000 / 66000 samples against the oracle -- sorted index
    rows (stable order) and every weak tree."""
    import ctypes as C
    from cascadeclassifier_b200._lib import check
    imgs, labels = make_samples(npos, nneg, W, W)
    n = len(imgs)
    kw = dict(boostType=bt, maxFalseAlarm=0.3, maxWeakCount=4)
    ok, st, ev = _train_gpu(gpu_ctx, ftype, mode, W, W, imgs, labels, **kw)
    o = orc.Evaluator(ftype, mode, W, W, n)
    o.set_images(imgs, labels)
    td = orc.TrainData(o, n, orc.SORT_STABLE, 16)
    ref = orc.boost_train(td, orc.boost_params(boost_type=bt, max_false_alarm=0.3, weak_count=4), cat_sort=orc.SORT_STABLE)
    assert ok == ref.ok
    assert_full_stage(st, ref)
    if ftype != orc.LBP:
        nr = 48
        rows = np.zeros((nr, n), np.uint32)
        check(gpu_ctx.lib.cc_boost_get_sorted_idx32(gpu_ctx.h, 0, nr, rows.ctypes.data_as(C.c_void_p)), "cc_boost_get_sorted_idx32")
        assert np.array_equal(rows, td.sidx()[:nr].astype(np.uint32))
