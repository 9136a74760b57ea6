"""GPU, BASELINE.json full sizes (configs[2]: Haar ALL 24x24, N = 15000 samples, F = 261600 features): size-independent
properties instead of an oracle run (the oracle needs hours at this size) -- sortedness and permutation property of the
sorted-index rows, fval == operator() on sampled rows, pruned search == exhaustive search, determinism of a repeated
stage training, integral-image checksum against numpy."""
import ctypes as C

import numpy as np
import pytest

import orc
from conftest import make_samples

pytestmark = pytest.mark.gpu

NPOS, NNEG = 5000, 10000


@pytest.fixture(scope="module")
def full(gpu_ctx):
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200._lib import check
    imgs, labels = make_samples(NPOS, NNEG)
    n = len(imgs)
    ev = ccb.CvFeatureEvaluator.create(orc.HAAR, gpu_ctx)
    ev.init(ccb.CvHaarFeatureParams(orc.ALL), n, (24, 24))
    ev.setImages(imgs, labels)
    check(gpu_ctx.lib.cc_boost_precalc(gpu_ctx.h, n), "precalc")
    return gpu_ctx, ev, imgs, labels, n


def test_fullsize_integrals_checksum(full):
    ctx, ev, imgs, labels, n = full
    rng = np.random.default_rng(0)
    for si in [0, 31, 32, n - 1] + [int(x) for x in rng.integers(0, n, 12)]:
        s, t, nf = ev.integrals(si)
        ref = np.zeros((25, 25), np.int64)
        ref[1:, 1:] = imgs[si].astype(np.int64).cumsum(0).cumsum(1)
        assert np.array_equal(s.reshape(25, 25), ref)
        # tilted integral: total mass of the last row's apexes is bounded by the image sum, corner cells are exact
        assert t.reshape(25, 25)[0].sum() == 0 and t.reshape(25, 25)[1, 1] == imgs[si][0, 0]
        inner = imgs[si][1:-1, 1:-1].astype(np.float64)
        assert np.float32(nf) == np.float32(np.sqrt(inner.size * (inner ** 2).sum() - inner.sum() ** 2))


def test_fullsize_sorted_rows_are_sorted_permutations(full):
    ctx, ev, imgs, labels, n = full
    from cascadeclassifier_b200._lib import check
    F = ev.getNumFeatures()
    assert F == 261600 and n == 15000
    rng = np.random.default_rng(1)
    rows = sorted(set([0, F - 16] + [int(x) for x in rng.integers(0, F - 16, 10)]))
    for r0 in rows:
        fv = np.zeros((16, n), np.float32)
        si = np.zeros((16, n), np.uint16)
        sv = np.zeros((16, n), np.float32)
        check(ctx.lib.cc_boost_get_fval(ctx.h, r0, 16, fv.ctypes.data_as(C.c_void_p)), "fval")
        check(ctx.lib.cc_boost_get_sorted_idx(ctx.h, r0, 16, si.ctypes.data_as(C.c_void_p)), "sidx")
        check(ctx.lib.cc_boost_get_sorted_val(ctx.h, r0, 16, sv.ctypes.data_as(C.c_void_p)), "sval")
        for k in range(16):
            idx = si[k].astype(np.int64)
            assert np.array_equal(np.sort(idx), np.arange(n))                   # a permutation of the samples
            vals = fv[k][idx]
            assert np.array_equal(vals.view(np.uint32), sv[k].view(np.uint32))  # sorted values are the gathered values
            assert (np.diff(vals) >= 0).all()                                   # ascending
            ties = np.diff(vals) == 0
            assert (np.diff(idx)[ties] > 0).all()                               # ties in ascending sample index (canonical order)
            # identical to numpy's stable argsort (-0.0 does not occur: nf > 0 or value 0)
            assert np.array_equal(idx, np.argsort(fv[k], kind="stable"))


def test_fullsize_fval_equals_operator_call(full):
    ctx, ev, imgs, labels, n = full
    from cascadeclassifier_b200._lib import check
    F = ev.getNumFeatures()
    rng = np.random.default_rng(2)
    for r0 in [0, F - 64] + [int(x) for x in rng.integers(0, F - 64, 6)]:
        got = np.zeros((64, n), np.float32)
        check(ctx.lib.cc_boost_get_fval(ctx.h, r0, 64, got.ctypes.data_as(C.c_void_p)), "fval")
        ref = ev.values(np.arange(r0, r0 + 64, dtype=np.int32), 0, n)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_fullsize_pruned_equals_exhaustive_and_is_deterministic(full):
    ctx, ev, imgs, labels, n = full
    import cascadeclassifier_b200 as ccb
    params = ccb.CvCascadeBoostParams(maxFalseAlarm=0.05, maxWeakCount=6)
    runs = []
    for mode in (0, 1, 0):
        ctx.set_search_mode(mode)
        st = ccb.CvCascadeBoost()
        assert st.train(ev, n, 1024, 1024, params)
        runs.append(st)
    ctx.set_search_mode(0)
    a, ex, b = runs
    assert len(a.trees) >= 3
    assert a.trees == ex.trees == b.trees and a.threshold == ex.threshold == b.threshold
    assert a.log == ex.log == b.log
    # the weak trees of a GAB stump stage: every split has positive quality and the stage reaches the requested rates
    assert all(t[0]["quality"] > 0 for t in a.trees)
    assert a.log[-1][1] >= 0.995 - 1e-6
