"""GPU parity: cc_createsamples (training samples from one image by random perspective distortions,
tools/createsamples/utility.cpp:419-672,952-1029) against the restatement in oracle/createsamples_ref.py, which takes
Rodrigues / gemm / solve / erode / dilate / GaussianBlur / INTER_LINEAR_EXACT resize from OpenCV itself (python cv2)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import createsamples_ref as cs  # noqa: E402


def _glyph(size):
    from cascadeclassifier_b200 import synth
    return np.clip(np.rint(synth._pattern(size)), 0, 255).astype(np.uint8)


def _photo(w, h, seed):
    from cascadeclassifier_b200 import synth
    return synth.make_background(w, h, seed, clutter=2.0)


@pytest.mark.parametrize("case", [
    dict(src=("glyph", 96), n=64, win=(24, 24), kw=dict()),                                                     # createsamples defaults
    dict(src=("photo", 75, 32), n=40, win=(75, 32), kw=dict(bgcolor=128, bgthreshold=10, maxxangle=0.6, maxyangle=0.3, maxzangle=1.0, rngseed=7)),
    dict(src=("glyph", 48), n=48, win=(20, 28), kw=dict(invert=cs.CV_RANDOM_INVERT, maxintensitydev=90, rngseed=99)),
    dict(src=("photo", 61, 97), n=32, win=(48, 96), kw=dict(invert=1, bgcolor=255, bgthreshold=60, rngseed=3)),
])
def test_createsamples_matches_oracle(gpu_ctx, case, tmp_path):
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200.createsamples import cvCreateTrainingSamples
    src = _glyph(case["src"][1]) if case["src"][0] == "glyph" else _photo(case["src"][1], case["src"][2], 5)
    w, h = case["win"]
    kw = case["kw"]
    path = str(tmp_path / "out.vec")
    got = cvCreateTrainingSamples(gpu_ctx, path, src, count=case["n"], winwidth=w, winheight=h, **kw)
    ref = cs.create_training_samples(src, case["n"], w, h, **kw)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), "%d of %d pixels differ (max %d)" % ((got != ref).sum(), got.size, np.abs(got.astype(int) - ref.astype(int)).max())
    assert got.std() > 5
    # the .vec on disk: reference layout (utility.cpp:128-153), readable by the trainer's PosReader mirror
    assert os.path.getsize(path) == 12 + case["n"] * (1 + 2 * w * h)
    assert np.array_equal(synth.read_vec(path, w, h), got)


def test_createsamples_feed_the_trainer(gpu_ctx, tmp_path):
    """end to end: 700 generated samples -> .vec -> two cascade stages (the reference's createsamples -> traincascade pipeline)"""
    import io
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    from cascadeclassifier_b200.createsamples import cvCreateTrainingSamples
    vec = str(tmp_path / "pos.vec")
    pos = cvCreateTrainingSamples(gpu_ctx, vec, _glyph(96), count=700, bgcolor=0, bgthreshold=0, maxxangle=0.4, maxyangle=0.4, maxzangle=0.3)
    assert pos.shape == (700, 24, 24)
    bgs = synth.make_backgrounds(6, 320, 240, seed0=60, clutter=1.0)
    cc = CvCascadeClassifier(gpu_ctx, out=io.StringIO())
    assert cc.train(str(tmp_path / "cascade"), vec, "bg.txt", 500, 800, 1024, 1024, 2, CvCascadeParams(0, 0, (24, 24)), ccb.CvHaarFeatureParams(0),
                    ccb.CvCascadeBoostParams(), backgrounds=bgs)
    assert len(cc.stageClassifiers) >= 1
