"""GPU parity: K5 mining (fillPassedSamples / NegReader / PosReader) and the whole cascade loop vs the oracle."""
import ctypes as C
import io
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _data(n_pos_vec, n_bg, W=24, H=24, bw=160, bh=120, seed=3):
    from cascadeclassifier_b200 import synth
    pos = synth.make_positives(n_pos_vec, W, H, seed=seed)
    bgs = synth.make_backgrounds(n_bg, bw, bh, seed0=seed * 50)
    return pos, bgs


def _state_dict(st):
    return dict(last=st.last, round=st.round, point=(st.point_x, st.point_y), offset=(st.offset_x, st.offset_y),
                img_size=(st.level_w, st.level_h), scale=st.scale)


@pytest.mark.parametrize("ftype,mode", [(orc.HAAR, orc.ALL), (orc.LBP, 0)])
def test_stage0_mining_matches_reader_stream(gpu_ctx, ftype, mode):
    """No stages yet: fillPassedSamples accepts the first `count` windows of the NegReader stream."""
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import _lib
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(50, 3, bw=100, bh=70)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    cc.cascadeParams = CvCascadeParams(0, ftype, (24, 24))
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    cc.featureEvaluator = ccb.CvFeatureEvaluator.create(ftype, gpu_ctx)
    cc.featureEvaluator.init(fp, 40 + 700, (24, 24))
    cc._create_readers(pos, bgs)
    cc.numPos, cc.numNeg = 40, 700
    cc.stageClassifiers = []
    cc._sync_cascade()
    got, consumed = cc.fillPassedSamples(0, 40, True, 0)
    assert (got, consumed) == (40, 40)
    got, consumed = cc.fillPassedSamples(40, 333, False, 1e-9)
    assert (got, consumed) == (333, 333)
    got2, consumed2 = cc.fillPassedSamples(373, 367, False, 1e-9)     # resumes mid-level, crosses images
    assert (got2, consumed2) == (367, 367)
    rd = orc.NegReader(24, 24, bgs)
    ref = np.stack([rd.get() for _ in range(700)])
    imgs = cc.featureEvaluator.images(0, 740)
    assert np.array_equal(imgs[:40], pos[:40])
    assert np.array_equal(imgs[40:], ref)
    assert _state_dict(cc._reader) == rd.state()
    # integrals of mined slots are those of the window
    o = orc.Evaluator(ftype, mode, 24, 24, 1)
    for si in (40, 41, 372, 373, 739):
        o.set_images(imgs[si:si + 1], np.zeros(1, np.uint8))
        s, t, nf = cc.featureEvaluator.integrals(si)
        os_, ot, onf = o.integrals(0)
        assert (s == os_).all()
        if ftype == orc.HAAR:
            assert (t == ot).all() and np.float32(nf) == np.float32(onf)
    labels = cc.featureEvaluator.getCls()
    assert (labels[:40] == 1).all() and (labels[40:740] == 0).all()


def _run_cascade(gpu_ctx, tmp_path, ftype, mode, W, H, numPos, numNeg, numStages, pos, bgs, **kw):
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    out = io.StringIO()
    cc = CvCascadeClassifier(gpu_ctx, out=out)
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    sp = ccb.CvCascadeBoostParams(boostType=kw.get("boost_type", 3), maxFalseAlarm=kw.get("max_false_alarm", 0.5),
                                  minHitRate=kw.get("min_hit_rate", 0.995), maxDepth=kw.get("max_depth", 1),
                                  maxWeakCount=kw.get("weak_count", 100))
    ok = cc.train(str(tmp_path), "pos.vec", "bg.txt", numPos, numNeg, 1024, 1024, numStages,
                  CvCascadeParams(0, ftype, (W, H)), fp, sp, positives=pos, backgrounds=bgs)
    o = orc.Cascade(ftype, mode, W, H, numPos, numNeg, numStages,
                    orc.boost_params(boost_type=kw.get("boost_type", orc.GAB), max_false_alarm=kw.get("max_false_alarm", 0.5),
                                     min_hit_rate=kw.get("min_hit_rate", 0.995), max_depth=kw.get("max_depth", 1),
                                     weak_count=kw.get("weak_count", 100)), sort_mode=orc.SORT_STABLE, threads=8)
    o.set_pos(pos)
    for b in bgs:
        o.add_bg(b)
    o.train()
    return ok, cc, o, out.getvalue()


@pytest.mark.parametrize("ftype,mode", [(orc.HAAR, orc.BASIC), (orc.LBP, 0)])
def test_cascade_training_matches_oracle(gpu_ctx, tmp_path, ftype, mode):
    pos, bgs = _data(260, 6)
    ok, cc, o, log = _run_cascade(gpu_ctx, tmp_path, ftype, mode, 24, 24, 200, 300, 4, pos, bgs)
    assert ok
    ref_stages = o.stages()
    ref_reports = o.reports()
    assert len(cc.stageClassifiers) == len(ref_stages), log
    from test_gpu_boost import compare_stage

    class _R:
        pass
    for s, (g, r) in enumerate(zip(cc.stageClassifiers, ref_stages)):
        rr = _R()
        rr.trees = r["trees"]
        agreed = compare_stage(g, rr)
        assert agreed == len(r["trees"]), "stage %d diverged at tree %d (tie)" % (s, agreed)
        assert g.threshold == pytest.approx(r["threshold"], rel=1e-6, abs=1e-7)
    for g, r in zip(cc.reports, ref_reports):
        for k in ("pos_count", "pos_consumed", "neg_count", "neg_consumed"):
            assert g[k] == r[k], (k, g, r)
        assert g["acceptance_ratio"] == pytest.approx(r["acceptance_ratio"])
    assert _state_dict(cc._reader) == o.neg_state()
    # files exist and OpenCV's own detector can load the result (test_integration.cpp:125-127)
    assert os.path.exists(os.path.join(str(tmp_path), "params.xml"))
    assert os.path.exists(os.path.join(str(tmp_path), "stage0.xml"))
    import cv2
    det = cv2.CascadeClassifier(os.path.join(str(tmp_path), "cascade.xml"))
    if ftype != orc.HOG:
        assert not det.empty()
    assert "===== TRAINING 0-stage =====" in log and "POS count : consumed" in log


@pytest.mark.parametrize("ftype,mode", [(orc.HAAR, orc.ALL), (orc.LBP, 0)])
def test_mining_paths_agree(gpu_ctx, tmp_path, ftype, mode):
    """The fused tile miner and the whole-level integral planes (fused builder, and the separate row/diagonal/column
    kernels) must select exactly the windows the literal per-window path (setImage + predict on a scratch slot) selects:
    same consumed count, same reader state, same accepted pixels, same integrals."""
    import copy
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import _lib
    from cascadeclassifier_b200._lib import check
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(300, 5, bw=230, bh=170)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    ok = cc.train(str(tmp_path), "pos.vec", "bg.txt", 200, 300, 1024, 1024, 3, CvCascadeParams(0, ftype, (24, 24)), fp,
                  ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
    assert len(cc.stageClassifiers) >= 1
    if ftype == orc.HAAR:
        # the point of this case: tilted features evaluated from a whole-level tilted integral.  The trained stages may
        # not use one, so a hand-made stage of three tilted stumps (one per tilted shape family) is appended.
        tilted = [fi for fi in range(0, cc.featureEvaluator.getNumFeatures(), 997) if cc.featureEvaluator.feature(fi)[2]][:3]
        assert len(tilted) == 3
        trees = [[dict(left=1, right=2, var_idx=fi, c=0.01 * (k - 1), value=0.0), dict(left=-1, right=-1, var_idx=-1, value=1.0),
                  dict(left=-1, right=-1, var_idx=-1, value=-1.0)] for k, fi in enumerate(tilted)]
        cc.stageClassifiers.append(ccb.CvCascadeBoost.from_trees(-0.5, trees))
    cc._sync_cascade()
    lib, h = gpu_ctx.lib, gpu_ctx.h
    start = copy.copy(cc._reader)
    results = []
    for m in (2, 1, 3, 0):      # slots (literal), level planes by separate kernels, fused tile miner, level planes by the fused builder
        check(lib.cc_mine_set_mode(h, m), "cc_mine_set_mode")
        st = _lib.ReaderState.from_buffer_copy(bytes(start))
        consumed, got = C.c_int64(0), C.c_int32(0)
        check(lib.cc_mine_negatives(h, 200, 250, 1e-7, C.byref(st), C.byref(consumed), C.byref(got)), "cc_mine_negatives")
        imgs = cc.featureEvaluator.images(200, got.value)
        ints = [cc.featureEvaluator.integrals(200 + k) for k in (0, got.value // 2, got.value - 1)]
        results.append((got.value, consumed.value, _state_dict(st), imgs, ints))
    check(lib.cc_mine_set_mode(h, 0), "cc_mine_set_mode")
    ref = results[0]
    assert ref[0] == 250 and ref[1] > 250
    for r in results[1:]:
        assert r[0] == ref[0] and r[1] == ref[1] and r[2] == ref[2]
        assert np.array_equal(r[3], ref[3])
        for a, b in zip(r[4], ref[4]):
            assert (a[0] == b[0]).all() and np.float32(a[2]) == np.float32(b[2])
            if a[1] is not None and b[1] is not None:
                assert (np.asarray(a[1]) == np.asarray(b[1])).all()
    # images whose width is not a multiple of the 16-byte row pitch, and one wider than 3072 pixels (the fused builder stages
    # source rows in shared memory; wide rows take the opt-in shared-memory size): literal path vs planes vs tiles again
    from cascadeclassifier_b200 import synth
    odd = [synth.make_background(3500, 60, 77, clutter=1.0), synth.make_background(333, 77, 78, clutter=2.0),
           synth.make_background(57, 33, 79), synth.make_background(1001, 99, 80, clutter=2.0)]
    check(lib.cc_bg_clear(h), "cc_bg_clear")
    for im in odd:
        im = np.ascontiguousarray(im)
        check(lib.cc_bg_register(h, im.ctypes.data_as(C.c_void_p), im.shape[1], im.shape[0]), "cc_bg_register")
    results = []
    for m in (2, 3, 0):
        check(lib.cc_mine_set_mode(h, m), "cc_mine_set_mode")
        st = _lib.ReaderState()
        consumed, got = C.c_int64(0), C.c_int32(0)
        check(lib.cc_mine_negatives(h, 200, 120, 1e-7, C.byref(st), C.byref(consumed), C.byref(got)), "cc_mine_negatives")
        results.append((got.value, consumed.value, _state_dict(st), cc.featureEvaluator.images(200, got.value)))
    check(lib.cc_mine_set_mode(h, 0), "cc_mine_set_mode")
    assert results[0][0] > 0
    for r in results[1:]:
        assert r[:3] == results[0][:3]
        assert np.array_equal(r[3], results[0][3])


@pytest.mark.parametrize("W,H,bw,bh", [(48, 96, 300, 330), (32, 40, 260, 200), (24, 24, 97, 61), (32, 64, 210, 260), (64, 128, 330, 300)])
def test_hog_miner_matches_slot_path(gpu_ctx, tmp_path, W, H, bw, bh):
    """HOG mining by level gradients + one CTA per window (k_hog_level_grad / k_hog_window_cascade, the default) must select
    exactly the windows the literal per-window path (scratch slot: setImage + predict, cc_mine_set_mode 2) selects -- same
    consumed count, reader state, accepted pixels and integral histograms -- over several images, pyramid levels and calls that
    resume inside a level."""
    import copy
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import _lib
    from cascadeclassifier_b200._lib import check
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(220, 4, W=W, H=H, bw=bw, bh=bh, seed=21)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    fp = ccb.CvFeatureParams.create(orc.HOG)
    cc.train(str(tmp_path), "pos.vec", "bg.txt", 150, 200, 1024, 1024, 3, CvCascadeParams(0, orc.HOG, (W, H)), fp,
             ccb.CvCascadeBoostParams(maxWeakCount=8), positives=pos, backgrounds=bgs)
    assert len(cc.stageClassifiers) >= 1
    cc._sync_cascade()
    lib, h = gpu_ctx.lib, gpu_ctx.h
    start = copy.copy(cc._reader)
    results = []
    for m in (2, 0):
        check(lib.cc_mine_set_mode(h, m), "cc_mine_set_mode")
        st = _lib.ReaderState.from_buffer_copy(bytes(start))
        rec = []
        for count in (90, 110):           # the second call resumes where the first one stopped (inside a level)
            consumed, got = C.c_int64(0), C.c_int32(0)
            check(lib.cc_mine_negatives(h, 150, count, 1e-7, C.byref(st), C.byref(consumed), C.byref(got)), "cc_mine_negatives")
            planes = [cc.featureEvaluator.hogPlanesCompact(150 + k).copy() for k in (0, got.value - 1)]
            rec.append((got.value, consumed.value, _state_dict(st), cc.featureEvaluator.images(150, got.value).copy(), planes))
        results.append(rec)
    # every window of every level of every image: same set of survivors (order-independent digest of their scan ordinals)
    scans = []
    for m in (2, 0):
        check(lib.cc_mine_set_mode(h, m), "cc_mine_set_mode")
        win, surv, dig = C.c_int64(0), C.c_int64(0), C.c_uint64(0)
        check(lib.cc_mine_scan_digest(h, 0, len(bgs), C.byref(win), C.byref(surv), C.byref(dig)), "cc_mine_scan_digest")
        scans.append((win.value, surv.value, dig.value))
    check(lib.cc_mine_set_mode(h, 0), "cc_mine_set_mode")
    assert scans[0] == scans[1] and 0 < scans[0][1] < scans[0][0], scans
    for a, b in zip(results[0], results[1]):
        assert a[0] == b[0] > 0 and a[1] == b[1] and a[2] == b[2], (a[:3], b[:3])
        assert np.array_equal(a[3], b[3])
        for pa, pb in zip(a[4], b[4]):
            assert np.array_equal(pa, pb)
    assert results[0][0][1] > results[0][0][0]          # the cascade rejected windows: the stages were really evaluated


@pytest.mark.parametrize("ftype,mode", [(orc.HAAR, orc.ALL), (orc.LBP, 0)])
def test_resume_from_stage_files(gpu_ctx, tmp_path, ftype, mode):
    """CvCascadeClassifier::load (src/cascadeclassifier.cpp:534-564): a second train() on a directory that holds
    params.xml + stage<i>.xml keeps those stages (leaf values / thresholds as the f32 the XML holds), announces them
    like the reference, trains only the missing stages and writes a cascade.xml OpenCV's detector loads."""
    import cv2
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(260, 6)
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    d = str(tmp_path)
    a = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    assert a.train(d, "pos.vec", "bg.txt", 200, 300, 1024, 1024, 2, CvCascadeParams(0, ftype, (24, 24)), fp,
                   ccb.CvCascadeBoostParams(maxFalseAlarm=0.7), positives=pos, backgrounds=bgs)
    assert len(a.stageClassifiers) >= 1
    # an interrupted run: only params.xml and stage0.xml are left in the directory
    for f in os.listdir(d):
        if f not in ("params.xml", "stage0.xml"):
            os.remove(os.path.join(d, f))
    out = io.StringIO()
    b = CvCascadeClassifier(gpu_ctx, out=out)
    # other parameters on the command line are ignored in favour of params.xml, as in the reference; numStages is not
    # part of params.xml (20 stages: the acceptance-ratio rule does not fire on the first rejected windows)
    assert b.train(d, "pos.vec", "bg.txt", 200, 300, 1024, 1024, 20, CvCascadeParams(0, ftype, (24, 24)), fp,
                   ccb.CvCascadeBoostParams(maxFalseAlarm=0.4), positives=pos, backgrounds=bgs)
    log = out.getvalue()
    assert "Training parameters are pre-loaded from the parameter file in data folder!" in log
    assert "Stage 0 is loaded" in log
    assert "===== TRAINING 0-stage =====" not in log and "===== TRAINING 1-stage =====" in log
    assert b.stageParams.maxFalseAlarm == pytest.approx(0.7)
    assert len(b.stageClassifiers) >= 2, log[-1500:]
    for sa, sb in zip(a.stageClassifiers[:1], b.stageClassifiers[:1]):
        assert sb.threshold == float(np.float32(sa.threshold))
        assert len(sa.trees) == len(sb.trees)
        for ta, tb in zip(sa.trees, sb.trees):
            assert ta[0]["var_idx"] == tb[0]["var_idx"]
            if ftype == orc.HAAR:
                assert float(np.float32(ta[0]["c"])) == tb[0]["c"]
            else:
                assert list(ta[0]["subset"]) == list(tb[0]["subset"])
            assert float(np.float32(ta[ta[0]["left"]]["value"])) == tb[tb[0]["left"]]["value"]
    # the loaded stages really are in the mining cascade: every mined negative passes them
    st = b.stageClassifiers[0]
    n = b.curNumSamples
    preds = st.predict(b.featureEvaluator, 0, False, n)
    assert (preds == 1.0).all()
    assert os.path.exists(os.path.join(d, "stage1.xml"))
    det = cv2.CascadeClassifier(os.path.join(d, "cascade.xml"))
    assert not det.empty()


def test_base_format_save_is_loadable(gpu_ctx, tmp_path):
    """-baseFormatSave (src/cascadeclassifier.cpp:457-531): the legacy opencv-haar-classifier XML is written for Haar
    cascades and OpenCV's detector loads it (the reference's own integration test checks exactly that,
    test_integration.cpp:287-343); other feature types are refused with the reference's message."""
    import cv2
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(260, 6)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    assert cc.train(str(tmp_path), "pos.vec", "bg.txt", 200, 300, 1024, 1024, 3, CvCascadeParams(0, orc.HAAR, (24, 24)),
                    ccb.CvHaarFeatureParams(orc.ALL), ccb.CvCascadeBoostParams(maxDepth=2), baseFormatSave=True, positives=pos, backgrounds=bgs)
    path = os.path.join(str(tmp_path), "cascade.xml")
    text = open(path).read()
    assert 'type_id="opencv-haar-classifier"' in text and "<stage_threshold>" in text and "<left_val>" in text
    det = cv2.CascadeClassifier(path)
    assert not det.empty()
    # the detector built from the legacy file fires on a positive sample pasted into a flat scene
    scene = np.full((96, 96), 128, np.uint8)
    scene[30:54, 40:64] = pos[0]
    hits = det.detectMultiScale(scene, scaleFactor=1.1, minNeighbors=0, minSize=(24, 24), maxSize=(24, 24))
    assert len(hits) >= 1
    lbp = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    with pytest.raises(ccb.CascadeCudaError, match="Haar-like features only"):
        lbp.train(str(tmp_path / "lbp"), "pos.vec", "bg.txt", 200, 300, 1024, 1024, 1, CvCascadeParams(0, orc.LBP, (24, 24)),
                  ccb.CvFeatureParams.create(orc.LBP), ccb.CvCascadeBoostParams(), baseFormatSave=True, positives=pos, backgrounds=bgs)


@pytest.mark.parametrize("ftype,mode,depth,stages", [(orc.LBP, 0, 1, 1), (orc.HAAR, orc.BASIC, 1, 2), (orc.HAAR, orc.ALL, 1, 1),
                                                     (orc.HAAR, orc.CORE, 2, 1)])
def test_reference_integration_scenarios(gpu_ctx, tmp_path, ftype, mode, depth, stages):
    """The shape of the reference's own end-to-end tests (test_integration.cpp:84-174,287-433): a 75x32 window, numPos=20,
    numNeg=1, 64 MB buffers, the synthetic negative image (r*7 + c*13) & 0xFF of 256x128 (test_integration.cpp:59-64),
    LBP / Haar BASIC / CORE / ALL, depth-2 trees.  The reference only asserts that the files exist and that
    cv::CascadeClassifier loads cascade.xml; here the run must also equal the oracle's (21 samples, up to 2.8 M features)."""
    import cv2
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    W, H = 75, 32
    pos = synth.make_positives(40, W, H, seed=5)
    r, c = np.mgrid[0:128, 0:256]
    bgs = [((r * 7 + c * 13) & 0xFF).astype(np.uint8)]
    out = io.StringIO()
    cc = CvCascadeClassifier(gpu_ctx, out=out)
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    ok = cc.train(str(tmp_path), "pos.vec", "bg.txt", 20, 1, 64, 64, stages, CvCascadeParams(0, ftype, (W, H)), fp,
                  ccb.CvCascadeBoostParams(maxDepth=depth), positives=pos, backgrounds=bgs)
    o = orc.Cascade(ftype, mode, W, H, 20, 1, stages, orc.boost_params(max_depth=depth), sort_mode=orc.SORT_STABLE, threads=8)
    o.set_pos(pos)
    o.add_bg(bgs[0])
    o.train()
    ref_stages = o.stages()
    assert ok == (len(ref_stages) > 0)
    assert len(cc.stageClassifiers) == len(ref_stages), out.getvalue()
    from test_gpu_boost import compare_stage

    class _R:
        pass
    for g, rs in zip(cc.stageClassifiers, ref_stages):
        rr = _R()
        rr.trees = rs["trees"]
        assert compare_stage(g, rr) == len(rs["trees"])
        assert g.threshold == pytest.approx(rs["threshold"], rel=1e-6, abs=1e-7)
    for g, rs in zip(cc.reports, o.reports()):
        for k in ("pos_count", "pos_consumed", "neg_count", "neg_consumed"):
            assert g[k] == rs[k], (k, g, rs)
    if ok:
        for f in ("params.xml", "stage0.xml", "cascade.xml"):
            assert os.path.exists(os.path.join(str(tmp_path), f))
        assert not cv2.CascadeClassifier(os.path.join(str(tmp_path), "cascade.xml")).empty()


_CASCADES = [
    # ftype, mode, W, H, numPos, numNeg, stages, backgrounds (w, h), parameters
    (orc.LBP, 0, 32, 20, 150, 250, 3, (210, 150), dict()),                                   # non-square window through the level planes
    (orc.HAAR, orc.CORE, 20, 28, 150, 250, 3, (190, 170), dict(max_depth=2)),                # depth-2 trees in the mining cascade
    (orc.HAAR, orc.ALL, 24, 24, 150, 250, 3, (200, 160), dict(min_hit_rate=0.95, max_false_alarm=0.4)),
    (orc.HAAR, orc.BASIC, 24, 24, 150, 250, 3, (200, 160), dict(boost_type=orc.RAB)),        # Real AdaBoost leaf values in predict()
    (orc.HOG, 0, 48, 96, 80, 120, 2, (260, 300), dict(weak_count=6)),                        # HOG: the window miner (k_hog_window_cascade)
]


@pytest.mark.parametrize("case", range(len(_CASCADES)))
def test_cascade_configuration_matrix(gpu_ctx, tmp_path, case):
    """Mining + stage training over several stages for flag combinations the BASELINE configs do not cover, against the
    oracle: every tree of every stage, consumed counts and the final NegReader state."""
    ftype, mode, W, H, npos, nneg, nst, (bw, bh), kw = _CASCADES[case]
    pos, bgs = _data(npos + 60, 6, W=W, H=H, bw=bw, bh=bh, seed=11 + case)
    ok, cc, o, log = _run_cascade(gpu_ctx, tmp_path, ftype, mode, W, H, npos, nneg, nst, pos, bgs, **kw)
    ref_stages, ref_reports = o.stages(), o.reports()
    from test_gpu_boost import compare_stage

    class _R:
        pass
    assert len(cc.stageClassifiers) == len(ref_stages), log
    for s, (g, r) in enumerate(zip(cc.stageClassifiers, ref_stages)):
        rr = _R()
        rr.trees = r["trees"]
        assert compare_stage(g, rr) == len(r["trees"]), "stage %d" % s       # strict: any difference fails (no unproven ties)
        assert g.threshold == pytest.approx(r["threshold"], rel=1e-6, abs=1e-7)
    for g, r in zip(cc.reports, ref_reports):
        for k in ("pos_count", "pos_consumed", "neg_count", "neg_consumed"):
            assert g[k] == r[k], (k, g, r)
    assert _state_dict(cc._reader) == o.neg_state()


def _mine_sequence(gpu_ctx, cc, bgs, plan, cache, first=200):
    """(re)register `bgs`, then mine with growing cascade prefixes from a fresh reader: [(n_stages, count), ...]"""
    import copy
    from cascadeclassifier_b200 import _lib
    from cascadeclassifier_b200._lib import check
    lib, h = gpu_ctx.lib, gpu_ctx.h
    check(lib.cc_mine_set_cache(h, cache, 0), "cc_mine_set_cache")
    check(lib.cc_bg_clear(h), "cc_bg_clear")
    for im in bgs:
        im = np.ascontiguousarray(im)
        check(lib.cc_bg_register(h, im.ctypes.data_as(C.c_void_p), im.shape[1], im.shape[0]), "cc_bg_register")
    stats0 = np.zeros(6, np.int64)
    check(lib.cc_mine_get_stats(h, stats0.ctypes.data_as(C.c_void_p)), "cc_mine_get_stats")
    st = _lib.ReaderState()
    all_stages = cc.stageClassifiers
    out = []
    for nst, count in plan:
        cc.stageClassifiers = all_stages[:nst]
        cc._sync_cascade()
        consumed, got = C.c_int64(0), C.c_int32(0)
        check(lib.cc_mine_negatives(h, first, count, 1e-9, C.byref(st), C.byref(consumed), C.byref(got)), "cc_mine_negatives")
        out.append((got.value, consumed.value, _state_dict(st), cc.featureEvaluator.images(first, got.value).copy(),
                    cc.featureEvaluator.integrals(first + got.value - 1)[0].copy()))
    cc.stageClassifiers = all_stages
    stats = np.zeros(6, np.int64)
    check(lib.cc_mine_get_stats(h, stats.ctypes.data_as(C.c_void_p)), "cc_mine_get_stats")
    return out, stats - stats0


@pytest.mark.parametrize("ftype,mode", [(orc.HAAR, orc.ALL), (orc.LBP, 0)])
def test_survivor_cache_is_transparent(gpu_ctx, tmp_path, ftype, mode):
    """The survivor cache (cc_mine_set_cache) must not change a single bit of what mining produces: got, consumed, the reader
    state after every call, the accepted pixels and their integrals -- over a stream that wraps the background list's round cycle
    several times (short list of small images), with a cascade that grows between the calls (stale entries are brought up to
    date by running only the new stages) and repeated calls with the same cascade (entries answer without any evaluation)."""
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200._lib import check
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(300, 5, bw=230, bh=170)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    fp = ccb.CvFeatureParams.create(ftype)
    if ftype == orc.HAAR:
        fp.mode = mode
    cc.train(str(tmp_path), "pos.vec", "bg.txt", 200, 300, 1024, 1024, 4, CvCascadeParams(0, ftype, (24, 24)), fp,
             ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
    ns = len(cc.stageClassifiers)
    assert ns >= 2
    # three tiny images: a pyramid is one to a few windows, the round offsets are clamped to a few dozen distinct values per
    # image -> the stream returns to the same (image, offsets) pyramid again and again, within a call and across calls
    small = [synth.make_background(44, 38, 301, clutter=2.0), synth.make_background(31, 52, 302, clutter=2.0),
             synth.make_background(30, 28, 303, clutter=1.0)]
    plan = [(1, 300)] * 2 + [(2, 300)] * 3 + [(min(3, ns), 300)] * 4 + [(ns, 300)] * 4 + [(ns, 77)]
    ref, s_off = _mine_sequence(gpu_ctx, cc, small, plan, cache=0)
    got, s_on = _mine_sequence(gpu_ctx, cc, small, plan, cache=1)
    check(gpu_ctx.lib.cc_mine_set_cache(gpu_ctx.h, 1, 0), "cc_mine_set_cache")
    for a, b in zip(ref, got):
        assert a[0] == b[0] and a[1] == b[1], (a[:3], b[:3])
        assert a[2] == b[2]
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    assert s_off[3] == 0 and s_off[2] == 0
    assert s_on[0] == s_off[0] > 3000                # same windows consumed
    assert s_on[3] > s_on[0] // 4, s_on              # a good part of the stream was answered by the cache ...
    assert s_on[2] > 0, s_on                         # ... and stale entries were brought up to date with the newer stages
    assert s_on[1] < s_off[1], (s_on, s_off)         # fewer windows evaluated from pyramid planes
    # a larger list: cache on vs off over full-size-like pyramids that are never revisited must also agree (pure insert path)
    ref2, _ = _mine_sequence(gpu_ctx, cc, bgs, [(ns, 200)], cache=0)
    got2, _ = _mine_sequence(gpu_ctx, cc, bgs, [(ns, 200)], cache=1)
    assert ref2[0][:3] == got2[0][:3] and np.array_equal(ref2[0][3], got2[0][3])


def test_fused_builder_stress_full_hd(gpu_ctx, tmp_path):
    """Race / determinism stress for the fused pyramid builder (k_lvl_build: cp.async staging ring, one barrier per level row):
    > 10^8 windows of 1920x1080 backgrounds through mode 0, three passes, every pass must select exactly the same windows
    (digest of the surviving scan ordinals) as the builder made of separate row / diagonal / column kernels (mode 1, no shared
    staging, no intra-CTA pipeline) on the whole set and as the literal per-window path (mode 2: setImage + predict per slot)
    and the fused tile miner (mode 3) on a subset.  compute-sanitizer is not available on the GPU pool
    (profiles/r02_compute_sanitizer_closed.log), so this is the check that would have caught the round-1 prologue race."""
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200._lib import check
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    pos, bgs = _data(300, 5, bw=230, bh=170)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    cc.train(str(tmp_path), "pos.vec", "bg.txt", 200, 300, 1024, 1024, 3, CvCascadeParams(0, orc.HAAR, (24, 24)),
             ccb.CvHaarFeatureParams(orc.ALL), ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
    # a stage of tilted stumps so that the tilted plane of the builder is exercised as well
    tilted = [fi for fi in range(0, cc.featureEvaluator.getNumFeatures(), 997) if cc.featureEvaluator.feature(fi)[2]][:2]
    trees = [[dict(left=1, right=2, var_idx=fi, c=0.02 * k, value=0.0), dict(left=-1, right=-1, var_idx=-1, value=1.0),
              dict(left=-1, right=-1, var_idx=-1, value=-1.0)] for k, fi in enumerate(tilted)]
    cc.stageClassifiers.append(ccb.CvCascadeBoost.from_trees(-1.5, trees))
    cc._sync_cascade()
    lib, h = gpu_ctx.lib, gpu_ctx.h
    check(lib.cc_bg_clear(h), "cc_bg_clear")
    distinct = [np.ascontiguousarray(synth.make_background(1920, 1080, 8800 + i, clutter=1.0)) for i in range(6)]
    NIMG = 2560                                   # 3.5e7 windows per pass
    for i in range(NIMG):
        b = distinct[i % len(distinct)]
        check(lib.cc_bg_register(h, b.ctypes.data_as(C.c_void_p), b.shape[1], b.shape[0]), "cc_bg_register")

    def scan(mode, nimg):
        check(lib.cc_mine_set_mode(h, mode), "cc_mine_set_mode")
        win, surv, dig = C.c_int64(0), C.c_int64(0), C.c_uint64(0)
        check(lib.cc_mine_scan_digest(h, 0, nimg, C.byref(win), C.byref(surv), C.byref(dig)), "cc_mine_scan_digest")
        return win.value, surv.value, dig.value
    try:
        fused = [scan(0, NIMG) for _ in range(3)]
        assert fused[0][0] > NIMG * 13000 and 3 * fused[0][0] > 10 ** 8       # ~13.6e3 windows per 1920x1080 pyramid
        assert fused[0][1] > 0
        assert fused[0] == fused[1] == fused[2], fused
        assert scan(1, NIMG) == fused[0]
        sub = scan(0, 96)
        assert scan(2, 96) == sub
        assert scan(3, 96) == sub
        # every registration of the same image must give the same survivors: the per-image survivor count is periodic
        per = [scan(0, k)[1] for k in (6, 12)]
        assert per[1] == 2 * per[0]
    finally:
        check(lib.cc_mine_set_mode(h, 0), "cc_mine_set_mode")
        check(lib.cc_bg_clear(h), "cc_bg_clear")
