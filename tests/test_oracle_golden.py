"""CPU: pin the oracle against (a) the committed cv2 golden fixtures, (b) the reference's own known answers
(traincascade/test/test_features.cpp:252-340,394-417,462-560), (c) the published feature counts
(traincascade/res/README.md:41,91)."""
import os
import struct

import numpy as np
import pytest

import orc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_golden.npz"))


@pytest.mark.parametrize("k", [0, 1, 2])
def test_integral_matches_cv2_integral3(k):
    img = GOLD["int%d_img" % k]
    s, q, t = orc.integral(img, tilted=True, sq=True)
    assert np.array_equal(s, GOLD["int%d_sum" % k])
    assert np.array_equal(q, GOLD["int%d_sq" % k])
    assert np.array_equal(t, GOLD["int%d_tilt" % k])


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_resize_linear_exact_matches_cv2(k):
    src, dst = GOLD["rs%d_src" % k], GOLD["rs%d_dst" % k]
    assert np.array_equal(orc.resize_linear_exact(src, dst.shape[1], dst.shape[0]), dst)


def test_cart_to_polar_restatement():
    dx, dy = GOLD["c2p_dx"], GOLD["c2p_dy"]
    L = orc.lib()
    ang = np.array([L.orc_fast_atan2(float(y), float(x)) for x, y in zip(dx, dy)], np.float32)
    # cv2's SIMD build fuses a few multiply-adds: <= 1e-6 rad, and never a different HOG bin on these inputs
    assert np.abs(ang - GOLD["c2p_ang"]).max() <= 1e-6
    b1 = np.floor(ang * np.float32(9 / np.pi) - np.float32(0.5))
    b2 = np.floor(GOLD["c2p_ang"] * np.float32(9 / np.pi) - np.float32(0.5))
    assert (b1 == b2).all()
    assert np.array_equal(np.sqrt(dx * dx + dy * dy).astype(np.float32), GOLD["c2p_mag"])


def test_published_feature_counts():
    L = orc.lib()
    assert L.orc_catalog_count(orc.LBP, 0, 75, 32) == 152625          # res/README.md:41
    assert L.orc_catalog_count(orc.HAAR, orc.BASIC, 75, 32) == 2790554  # res/README.md:91
    assert [L.orc_catalog_count(orc.HAAR, m, 24, 24) for m in (0, 1, 2)] == [162336, 210400, 261600]
    assert L.orc_catalog_count(orc.LBP, 0, 24, 24) == 8464
    assert L.orc_catalog_count(orc.HOG, 0, 48, 96) == 591


def _calc(stride, tilted, rects, wts, integral):
    r = np.zeros((3, 4), np.int32)
    w = np.zeros(3, np.float32)
    for i, (rc, wt) in enumerate(zip(rects, wts)):
        r[i], w[i] = rc, wt
    integral = np.ascontiguousarray(integral.ravel(), np.int32)
    return orc.lib().orc_haar_calc(stride, int(tilted), r.ctypes.data, w.ctypes.data, integral.ctypes.data)


def test_reference_known_answers_feature_calc():
    img = np.zeros((8, 8), np.uint8); img[:, 4:] = 100                 # test_features.cpp:462-483
    s, _, _ = orc.integral(img)
    assert _calc(9, False, [(0, 0, 4, 8), (4, 0, 4, 8)], [1.0, -1.0], s) == -3200.0
    img = np.full((8, 8), 42, np.uint8)                                # :485-502
    s, _, _ = orc.integral(img)
    assert _calc(9, False, [(0, 0, 4, 8), (4, 0, 4, 8)], [1.0, -1.0], s) == 0.0
    img = np.zeros((3, 9), np.uint8); img[:, 3:6] = 200                # :504-529
    s, _, _ = orc.integral(img)
    assert _calc(10, False, [(0, 0, 9, 3), (3, 0, 3, 3)], [1.0, -3.0], s) == -3600.0
    img = np.ones((16, 16), np.uint8)                                  # :531-560
    _, _, t = orc.integral(img, tilted=True)
    assert _calc(17, True, [(8, 2, 4, 4)], [1.0], t) == 32.0


def test_flat_image_invariants():
    flat = np.full((1, 24, 24), 128, np.uint8)
    for ftype, expect in [(orc.HAAR, 0.0), (orc.LBP, 255.0)]:
        e = orc.Evaluator(ftype, orc.ALL, 24, 24, 1)
        e.set_images(flat, np.ones(1, np.uint8))
        vals = e.values(np.arange(0, e.num_features, 101, dtype=np.int32), 0, 1)
        assert (vals == expect).all()
    flat = np.full((1, 96, 48), 77, np.uint8)
    e = orc.Evaluator(orc.HOG, 0, 48, 96, 1)
    e.set_images(flat, np.ones(1, np.uint8))
    assert (e.values(np.arange(0, e.var_count, 97, dtype=np.int32), 0, 1) == 0.0).all()


def test_hog_planes_vs_cv2_pipeline():
    """integralHistogram rebuilt from cv2 primitives (cartToPolar + sequential integrals); tolerance as SURVEY A.4."""
    import cv2
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (96, 48), dtype=np.uint8)
    H, W = img.shape
    p = np.pad(img.astype(np.int32), 1, mode="edge")
    dx = (p[1:-1, 2:] - p[1:-1, :-2]).astype(np.float32)
    dy = (p[2:, 1:-1] - p[:-2, 1:-1]).astype(np.float32)
    mag, ang = cv2.cartToPolar(dx, dy, angleInDegrees=False)
    a = ang * np.float32(9 / np.pi) - np.float32(0.5)
    b = np.floor(a).astype(np.int32)
    b[b < 0] += 9
    b[b >= 9] -= 9
    planes = orc.hog_planes(img)
    norm = np.zeros((H + 1, W + 1), np.float64)
    norm[1:, 1:] = np.cumsum(np.cumsum(mag.astype(np.float64), 0), 1)
    assert np.allclose(planes[9], norm, rtol=2e-6, atol=1e-2)
    for k in range(9):
        m = np.where(b == k, mag, 0).astype(np.float64)
        ref = np.zeros((H + 1, W + 1))
        ref[1:, 1:] = np.cumsum(np.cumsum(m, 0), 1)
        assert np.allclose(planes[k], ref, rtol=2e-6, atol=1e-2)


def test_vec_reader_on_reference_fixture():
    from cascadeclassifier_b200 import synth
    count, vsz, size = [int(v) for v in GOLD["vec_header"]]
    assert (count, vsz, size) == (100, 2400, 480112)                   # barcode.vec: 12 + 100*(1+2*2400)
    raw = GOLD["vec_first3_raw"].tobytes()
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".vec", delete=False) as f:
        f.write(struct.pack("<iihh", 3, vsz, 0, 0))
        f.write(raw)
        name = f.name
    try:
        s = synth.read_vec(name, 75, 32)
        assert s.shape == (3, 32, 75)
        vals = np.frombuffer(raw, np.uint8).reshape(3, 1 + 2 * vsz)[:, 1:].copy().view("<i2")
        assert np.array_equal(s.reshape(3, -1), vals.astype(np.uint8))
        synth.write_vec(name, s)
        assert open(name, "rb").read()[12:] == raw                      # writer reproduces the layout byte for byte
    finally:
        os.unlink(name)


def test_oracle_stage_training_sanity():
    from conftest import make_samples
    imgs, labels = make_samples(120, 240)
    e = orc.Evaluator(orc.HAAR, orc.BASIC, 24, 24, len(imgs))
    e.set_images(imgs, labels)
    rows = np.arange(0, e.num_features, 16, dtype=np.int32)
    td = orc.TrainData(e, len(imgs), orc.SORT_STABLE, 8, subset=rows)
    fv, si = td.fval(), td.sidx()
    assert (np.diff(np.take_along_axis(fv, si.astype(np.int64), 1), axis=1) >= 0).all()      # sorted
    a = orc.boost_train(td, orc.boost_params(max_false_alarm=0.05, weak_count=10), record=True)
    assert a.ok and len(a.trees) >= 1 and a.log[-1][1] >= 0.995
    w, m = a.snapshot(0)
    assert abs(w.sum() - 1.0) < 1e-12 and m.sum() > 0
    # tie order of the sort does not change the chosen splits (SURVEY hard part 2)
    td2 = orc.TrainData(e, len(imgs), orc.SORT_STD, 8, subset=rows)
    b = orc.boost_train(td2, orc.boost_params(max_false_alarm=0.05, weak_count=10))
    assert [t[0]["var_idx"] for t in a.trees] == [t[0]["var_idx"] for t in b.trees]
    assert [t[0]["c"] for t in a.trees] == [t[0]["c"] for t in b.trees]
