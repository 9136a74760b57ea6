"""CPU: the createsamples restatement (oracle/createsamples_ref.py) pinned to OpenCV, and the host half of cc_createsamples
(cc_createsamples_plan: cv::RNG stream, Rodrigues, perspective systems, edge walks, crops) against it."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import createsamples_ref as cs  # noqa: E402


def _glyph(size=96):
    from cascadeclassifier_b200 import synth
    return np.clip(np.rint(synth._pattern(size)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("seed", [12345, 1, 987654321])
def test_cv_rng_restatement_matches_opencv(seed):
    """theRNG().next(): cv2.randShuffle draws `next() % n` from the same generator (core/src/rand.cpp randShuffle_)"""
    import cv2
    cv2.setRNGSeed(seed)
    a = np.arange(1000, dtype=np.int32).reshape(1, -1)
    cv2.randShuffle(a)
    r = cs.CvRNG(seed)
    b = np.arange(1000, dtype=np.int32)
    for i in range(1000):
        j = r.next() % 1000
        b[j], b[i] = b[i], b[j]
    assert np.array_equal(a.ravel(), b)


def test_gaussian_blur_3x3_u8_formula():
    """what k_cs_blur computes: ([1 2 1] x [1 2 1] * p + 8) >> 4 with BORDER_REFLECT_101 == cv2.GaussianBlur(u8, (3, 3), 0)"""
    import cv2
    rng = np.random.default_rng(0)
    for img in (rng.integers(0, 256, (67, 83), dtype=np.uint8), (rng.integers(0, 2, (40, 31)) * 255).astype(np.uint8)):
        p = np.pad(img.astype(np.int64), 1, mode="reflect")
        s = np.zeros(img.shape, np.int64)
        for dy in range(3):
            for dx in range(3):
                s += (2 - (dy != 1)) * (2 - (dx != 1)) * p[dy:dy + img.shape[0], dx:dx + img.shape[1]]
        assert np.array_equal((s + 8) >> 4, cv2.GaussianBlur(img, (3, 3), 0, 0))


def test_warp_of_identity_quad_reproduces_the_image():
    """the tool's own cvWarpPerspective (utility.cpp:226-417): the quadrangle of the undistorted image maps every pixel onto itself"""
    src = _glyph(48)
    h, w = src.shape
    quad = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float64) + 10.0
    dst = np.zeros((h + 20, w + 20), np.uint8)
    with np.errstate(all="ignore"):          # axis-aligned edges: the reference's edge equations divide by zero here too (k = inf is never used)
        cs.warp_perspective(src, dst, quad)
    inner = dst[11:10 + h, 10:10 + w]          # the walk starts one row below the top vertex (y_min = q.y - 1 -> iy_min = round(y_min) + 1)
    assert np.abs(inner.astype(int) - src[1:].astype(int)).max() <= 1
    assert dst[:10].max() == 0 and dst[:, :10].max() == 0


@pytest.mark.parametrize("w,h,win,invert,seed", [(96, 96, (24, 24), 0, 12345), (75, 32, (75, 32), cs.CV_RANDOM_INVERT, 7), (40, 64, (20, 32), 1, 99)])
def test_createsamples_plan_matches_oracle(w, h, win, invert, seed):
    """cc_createsamples_plan (C++: own cv::RNG, Rodrigues, 8x8 LU, edge walk) vs the oracle (cv2.Rodrigues / gemm / solve):
    quadrangles and coefficients to 1e-11 relative, crops / intensity deviation / inversion / number of warped rows and pixels exact."""
    from cascadeclassifier_b200 import _lib
    lib = _lib.load()
    n = 40
    quads = np.zeros((n, 4, 2), np.float64)
    coeffs = np.zeros((n, 9), np.float64)
    misc = np.zeros((n, 8), np.int32)
    _lib.check(lib.cc_createsamples_plan(w, h, n, invert, 40, 1.1, 1.1, 0.5, win[0], win[1], seed, quads.ctypes.data_as(C.c_void_p),
                                         coeffs.ctypes.data_as(C.c_void_p), misc.ctypes.data_as(C.c_void_p)), "cc_createsamples_plan")
    rng = cs.CvRNG(seed)
    dx, dy = w // 2, h // 2
    inverse = invert
    for i in range(n):
        if invert == cs.CV_RANDOM_INVERT:
            inverse = rng.uniform_i(0, 2)
        q, _, _ = cs.random_quad(rng, w, h, 1.1, 1.1, 0.5)
        q = q + np.array([float(dx), float(dy)])
        np.testing.assert_allclose(quads[i], q, rtol=1e-11, atol=1e-9)
        c = cs.perspective_transform(w, h, q)
        np.testing.assert_allclose(coeffs[i], c.ravel(), rtol=1e-9, atol=1e-11)
        rows = cs.warp_rows(q, w + 2 * dx, h + 2 * dy)
        rng.uniform_d(0., 0.); rng.uniform_d(0., 0.); rng.uniform_d(0., 0.)          # xshift, yshift, randscale draws
        fdev = rng.uniform_i(-40, 40)
        assert misc[i, 4] == fdev and misc[i, 5] == inverse
        assert misc[i, 6] == len(rows)
        assert misc[i, 7] == sum(max(0, x1 - x0 + 1) for _, x0, x1 in rows)
        f = np.float32
        scale = max(f(w) / f(win[0]), f(h) / f(win[1]))
        rx = int(f(-0.5) * (scale * f(win[0]) - f(w)) + f(dx))
        ry = int(f(-0.5) * (scale * f(win[1]) - f(h)) + f(dy))
        rw, rh = int(scale * f(win[0])), int(scale * f(win[1]))
        x0, y0 = max(rx, 0), max(ry, 0)
        assert tuple(misc[i, :4]) == (x0, y0, min(rx + rw, w + 2 * dx) - x0, min(ry + rh, h + 2 * dy) - y0)


def test_oracle_samples_are_deterministic_and_in_vec_layout(tmp_path):
    src = _glyph(64)
    a = cs.create_training_samples(src, 6, 24, 24, rngseed=5)
    b = cs.create_training_samples(src, 6, 24, 24, rngseed=5)
    c = cs.create_training_samples(src, 6, 24, 24, rngseed=6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.std() > 10          # not blank
    from cascadeclassifier_b200 import synth
    p = str(tmp_path / "s.vec")
    cs.write_vec(p, a)
    assert os.path.getsize(p) == 12 + 6 * (1 + 2 * 24 * 24)           # utility.cpp:128-153
    assert np.array_equal(synth.read_vec(p, 24, 24), a)
