"""GPU parity: K1 integral images, operator(), K2 precalc, K3 sort — through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import orc
from conftest import make_samples

pytestmark = pytest.mark.gpu


def _gpu_eval(ftype, mode, W, H, n, ctx):
    import cascadeclassifier_b200 as ccb
    ev = ccb.CvFeatureEvaluator.create(ftype, ctx)
    params = ccb.CvFeatureParams.create(ftype)
    if ftype == ccb.CvFeatureParams.HAAR:
        params.mode = mode
    ev.init(params, n, (W, H))
    return ev


@pytest.mark.parametrize("mode", [orc.BASIC, orc.ALL])
def test_haar_integrals_bit_exact(gpu_ctx, mode):
    imgs, labels = make_samples(40, 61)     # 101 samples: partial last tile
    n = len(imgs)
    ev = _gpu_eval(orc.HAAR, mode, 24, 24, n, gpu_ctx)
    ev.setImages(imgs, labels)
    o = orc.Evaluator(orc.HAAR, mode, 24, 24, n)
    o.set_images(imgs, labels)
    for si in [0, 1, 31, 32, 63, 64, 100]:
        s, t, nf = ev.integrals(si)
        os_, ot, onf = o.integrals(si)
        assert (s == os_).all()
        if mode == orc.ALL:
            assert (t == ot).all()
        assert np.float32(nf) == np.float32(onf)
    assert (ev.getCls()[:n] == labels).all()


def test_set_image_single_slot_isolated(gpu_ctx):
    """setImage(img, label, idx) touches only slot idx (test_features.cpp:364-392)."""
    imgs, labels = make_samples(10, 30)
    ev = _gpu_eval(orc.HAAR, orc.ALL, 24, 24, 40, gpu_ctx)
    ev.setImages(imgs, labels)
    before = [ev.integrals(si) for si in (4, 5, 6)]
    new = (255 - imgs[5])
    ev.setImage(new, 1, 5)
    s4, t4, n4 = ev.integrals(4)
    assert (s4 == before[0][0]).all() and (t4 == before[0][1]).all() and n4 == before[0][2]
    s6, t6, n6 = ev.integrals(6)
    assert (s6 == before[2][0]).all() and (t6 == before[2][1]).all() and n6 == before[2][2]
    s5, t5, n5 = ev.integrals(5)
    rs, _, rt = orc.integral(new, tilted=True)
    assert (s5 == rs.ravel()).all() and (t5 == rt.ravel()).all()
    assert ev.getCls(5) == 1.0


def test_wide_window_integrals(gpu_ctx):
    """75x32 (the reference's barcode window): multi-chunk rows, unstaged output path."""
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (37, 32, 75), dtype=np.uint8)
    labels = np.zeros(37, np.uint8)
    ev = _gpu_eval(orc.HAAR, orc.ALL, 75, 32, 37, gpu_ctx)
    ev.setImages(imgs, labels)
    for si in [0, 17, 36]:
        s, t, nf = ev.integrals(si)
        rs, rq, rt = orc.integral(imgs[si], tilted=True, sq=True)
        assert (s == rs.ravel()).all() and (t == rt.ravel()).all()
        assert np.float32(nf) == np.float32(orc.norm_factor(imgs[si]))


@pytest.mark.parametrize("ftype,mode,W,H", [(orc.HAAR, orc.BASIC, 24, 24), (orc.HAAR, orc.ALL, 24, 24),
                                            (orc.LBP, 0, 24, 24), (orc.HOG, 0, 48, 96), (orc.HAAR, orc.CORE, 20, 16)])
def test_operator_call_matches_oracle(gpu_ctx, ftype, mode, W, H):
    imgs, labels = make_samples(20, 45, W, H)
    n = len(imgs)
    ev = _gpu_eval(ftype, mode, W, H, n, gpu_ctx)
    ev.setImages(imgs, labels)
    o = orc.Evaluator(ftype, mode, W, H, n)
    o.set_images(imgs, labels)
    rng = np.random.default_rng(1)
    nv = o.var_count
    vis = np.unique(np.concatenate([rng.integers(0, nv, 400), [0, nv - 1]])).astype(np.int32)
    g = ev.values(vis, 0, n)
    r = o.values(vis, 0, n)
    if ftype == orc.HOG:
        # f32 planes built in the same sequential order: expected bit-exact vs the oracle
        assert np.array_equal(g, r)
    else:
        assert np.array_equal(g.view(np.uint32), r.view(np.uint32))   # bit-exact incl. Haar division
    assert ev(int(vis[3]), 7) == r[3, 7]


def test_flat_image_invariants(gpu_ctx):
    """Haar == 0, LBP == 255, HOG == 0 on a constant image (test_features.cpp:252-340,394-417)."""
    import cascadeclassifier_b200 as ccb
    flat = np.full((1, 24, 24), 128, np.uint8)
    for ftype, expect in [(orc.HAAR, 0.0), (orc.LBP, 255.0)]:
        ev = _gpu_eval(ftype, orc.ALL, 24, 24, 1, gpu_ctx)
        ev.setImages(flat, np.ones(1, np.uint8))
        vals = ev.values(np.arange(0, ev.getNumFeatures(), 97, dtype=np.int32), 0, 1)
        assert (vals == expect).all()
    flat = np.full((1, 96, 48), 77, np.uint8)
    ev = _gpu_eval(orc.HOG, 0, 48, 96, 1, gpu_ctx)
    ev.setImages(flat, np.ones(1, np.uint8))
    vals = ev.values(np.arange(0, ev.getNumFeatures() * 36, 53, dtype=np.int32), 0, 1)
    assert (vals == 0.0).all()
    assert isinstance(ev, ccb.CvFeatureEvaluator)


@pytest.mark.parametrize("W,H", [(48, 96), (64, 64), (50, 38)])
def test_hog_planes_match_oracle(gpu_ctx, W, H):
    """the 9 integral histograms + the norm plane, bit for bit: in the reference's full layout (parity getter) and as they are
    resident -- only rows / columns that are multiples of the block step 4, the only entries a HOG feature can touch"""
    imgs, labels = make_samples(20, 30, W, H)
    ev = _gpu_eval(orc.HOG, 0, W, H, 50, gpu_ctx)
    ev.setImages(imgs, labels)
    for si in (0, 6, 31, 32, 49):
        g = ev.hogPlanes(si)
        r = orc.hog_planes(imgs[si]).reshape(10, -1)
        assert np.array_equal(g, r)
        c = ev.hogPlanesCompact(si)
        full = r.reshape(10, H + 1, W + 1)
        assert np.array_equal(c, full[:, ::4, ::4][:, :H // 4 + 1, :W // 4 + 1])


@pytest.mark.parametrize("ftype,mode,W,H,npos,nneg", [(orc.HAAR, orc.BASIC, 24, 24, 150, 317), (orc.HAAR, orc.ALL, 24, 24, 90, 200),
                                                      (orc.LBP, 0, 24, 24, 100, 233), (orc.HOG, 0, 48, 96, 40, 60),
                                                      # windows whose 32-sample integral tile exceeds shared memory (taps read from HBM / L2)
                                                      (orc.HAAR, orc.ALL, 52, 44, 20, 45), (orc.LBP, 0, 75, 32, 30, 40)])
def test_precalc_and_sort_match_oracle(gpu_ctx, ftype, mode, W, H, npos, nneg):
    import ctypes as C
    from cascadeclassifier_b200._lib import check
    imgs, labels = make_samples(npos, nneg, W, H)
    n = len(imgs)
    ev = _gpu_eval(ftype, mode, W, H, n, gpu_ctx)
    ev.setImages(imgs, labels)
    lib, h = gpu_ctx.lib, gpu_ctx.h
    check(lib.cc_boost_precalc(h, n), "precalc")
    o = orc.Evaluator(ftype, mode, W, H, n)
    o.set_images(imgs, labels)
    F = o.var_count
    rng = np.random.default_rng(5)
    starts = sorted(set([0, max(0, F - 64)] + [int(x) for x in rng.integers(0, max(1, F - 64), 6)]))
    for r0 in starts:
        nr = min(64, F - r0)
        rows = np.arange(r0, r0 + nr, dtype=np.int32)
        td = orc.TrainData(o, n, orc.SORT_STABLE, 8, subset=rows)
        g = np.zeros((nr, n), np.float32)
        check(lib.cc_boost_get_fval(h, r0, nr, g.ctypes.data_as(C.c_void_p)), "get_fval")
        assert np.array_equal(g.view(np.uint32), td.fval().view(np.uint32)), "fval rows %d.." % r0
        if ftype != orc.LBP:
            gi = np.zeros((nr, n), np.uint16)
            check(lib.cc_boost_get_sorted_idx(h, r0, nr, gi.ctypes.data_as(C.c_void_p)), "get_sidx")
            assert np.array_equal(gi, td.sidx()), "sorted idx rows %d.." % r0
            gv = np.zeros((nr, n), np.float32)
            check(lib.cc_boost_get_sorted_val(h, r0, nr, gv.ctypes.data_as(C.c_void_p)), "get_sval")
            assert np.array_equal(gv, np.take_along_axis(td.fval(), td.sidx().astype(np.int64), axis=1))


@pytest.mark.parametrize("ftype,mode,W,H,npos,nneg,flat", [(orc.HAAR, orc.ALL, 24, 24, 700, 1500, 0), (orc.HAAR, orc.BASIC, 24, 24, 1200, 2500, 900),
                                                           (orc.HAAR, orc.CORE, 24, 24, 2500, 5200, 3), (orc.HOG, 0, 48, 96, 300, 500, 0)])
def test_sort_paths_agree(gpu_ctx, ftype, mode, W, H, npos, nneg, flat):
    """The bucket + rank sort (with the rows it defers to the radix kernel) writes exactly the rows the stable radix
    passes write.  `flat` constant windows (norm factor 0 -> value 0 in every row) make long runs of equal values: every
    row is deferred; a few of them make a tie group inside one bucket; HOG rows hold many exact zeros."""
    import ctypes as C
    from cascadeclassifier_b200._lib import check
    imgs, labels = make_samples(npos, nneg, W, H)
    imgs = imgs.copy()
    rng = np.random.default_rng(11)
    for i in rng.choice(len(imgs), flat, replace=False):
        imgs[i] = rng.integers(0, 256)
    n = len(imgs)
    ev = _gpu_eval(ftype, mode, W, H, n, gpu_ctx)
    ev.setImages(imgs, labels)
    lib, h = gpu_ctx.lib, gpu_ctx.h
    F = ev.getNumFeatures() * (36 if ftype == orc.HOG else 1)
    rows = sorted(set([0, F - 256] + [int(x) for x in rng.integers(0, F - 256, 8)]))
    out = {}
    try:
        for sm in (1, 2):
            gpu_ctx.set_sort_mode(sm)
            check(lib.cc_boost_precalc(h, n), "precalc")
            if sm == 2:
                deferred = gpu_ctx.sort_deferred_rows()
                assert 0 <= deferred <= F
                if flat >= 900:
                    assert deferred == F          # 900 equal values: sum k^2 >= 810000 > 96 N
                if ftype == orc.HAAR and flat <= 3:
                    assert deferred < F // 4      # the bucket path is the one under test
            got = []
            for r0 in rows:
                gi = np.zeros((256, n), np.uint16)
                gv = np.zeros((256, n), np.float32)
                check(lib.cc_boost_get_sorted_idx(h, r0, 256, gi.ctypes.data_as(C.c_void_p)), "get_sidx")
                check(lib.cc_boost_get_sorted_val(h, r0, 256, gv.ctypes.data_as(C.c_void_p)), "get_sval")
                got.append((gi, gv))
            out[sm] = got
    finally:
        gpu_ctx.set_sort_mode(0)
    for (ai, av), (bi, bv) in zip(out[1], out[2]):
        assert np.array_equal(ai, bi)
        assert np.array_equal(av.view(np.uint32), bv.view(np.uint32))
    # and both are the canonical order
    gi, gv = out[2][0]
    fv = np.zeros((256, n), np.float32)
    check(lib.cc_boost_get_fval(h, rows[0], 256, fv.ctypes.data_as(C.c_void_p)), "get_fval")
    for k in range(0, 256, 17):
        assert np.array_equal(gi[k].astype(np.int64), np.argsort(fv[k], kind="stable"))


def test_fast_division_is_exact(gpu_ctx):
    """The reciprocal-based Haar division of K2 equals IEEE division on 2e9 (integer, norm-factor) pairs."""
    import ctypes as C
    from cascadeclassifier_b200._lib import check
    for steps in (2, 1):     # 1 step is what the separable kernels use (2 for warps holding an all-ones-significand divisor)
        m = C.c_uint64(1)
        check(gpu_ctx.lib.cc_debug_div_check(gpu_ctx.h, steps, 12345 + steps, 2_000_000_000, C.byref(m)), "div_check")
        assert m.value == 0


@pytest.mark.parametrize("mode,W,H", [(orc.ALL, 24, 24), (orc.ALL, 20, 28), (orc.CORE, 30, 18)])
def test_separable_precalc_equals_operator_call_everywhere(gpu_ctx, mode, W, H):
    """Every row of the fval matrix written by the separable kernels (upright row/column patterns, tilted a/b patterns)
    equals operator()(featureIdx, sampleIdx) evaluated tap by tap (k_eval_values, itself pinned to the oracle above):
    all features, non-square windows included."""
    import ctypes as C
    from cascadeclassifier_b200._lib import check
    imgs, labels = make_samples(30, 45, W, H)
    n = len(imgs)
    ev = _gpu_eval(orc.HAAR, mode, W, H, n, gpu_ctx)
    ev.setImages(imgs, labels)
    lib, h = gpu_ctx.lib, gpu_ctx.h
    check(lib.cc_boost_precalc(h, n), "precalc")
    F = ev.getNumFeatures()
    step = 8192
    for r0 in range(0, F, step):
        nr = min(step, F - r0)
        got = np.zeros((nr, n), np.float32)
        check(lib.cc_boost_get_fval(h, r0, nr, got.ctypes.data_as(C.c_void_p)), "get_fval")
        ref = ev.values(np.arange(r0, r0 + nr, dtype=np.int32), 0, n)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "rows %d..%d differ" % (r0, r0 + nr)
