"""The oracle's stage trainer against cv2.ml.Boost (see tests/cv2_boost_ref.py for what is and is not comparable).
This is the independent pin of the training half of the oracle: chosen variable, threshold (f32-exact), split quality and
leaf values of the first weak trees, with and without weight trimming."""
import numpy as np
import pytest

import orc
from conftest import make_samples
from cv2_boost_ref import assert_stumps_match, cv2_boost_stumps


@pytest.fixture(scope="module")
def haar12():
    imgs, labels = make_samples(120, 200, 12, 12)
    n = len(imgs)
    ev = orc.Evaluator(orc.HAAR, orc.ALL, 12, 12, n)
    ev.set_images(imgs, labels)
    td = orc.TrainData(ev, n, orc.SORT_STABLE, 4)
    return td, labels, td.fval().T.copy()


@pytest.mark.parametrize("trim", [0.95, 0.0])
@pytest.mark.parametrize("bt", ["GAB", "LB", "RAB"])
def test_oracle_ordered_stumps_match_cv2_boost(haar12, bt, trim):
    td, labels, X = haar12
    kind = {"GAB": orc.GAB, "LB": orc.LB, "RAB": orc.RAB}[bt]
    # maxFalseAlarm 0 / minHitRate 1: the stage stops only when the training set is separated (or at weak_count)
    ref = orc.boost_train(td, orc.boost_params(boost_type=kind, weak_count=8, max_false_alarm=0.0, min_hit_rate=1.0, weight_trim_rate=trim))
    cvt = cv2_boost_stumps(X, labels, bt, 8, trim)
    k = assert_stumps_match(ref.trees, cvt, 3, check_quality=(bt != "RAB"), what="%s trim %g" % (bt, trim))
    assert k >= 3


def test_oracle_std_sort_mode_matches_cv2_boost(haar12):
    # the faithful std::sort tie order (SORT_STD) gives the same stumps: splits never separate equal values
    td_stable, labels, X = haar12
    td = orc.TrainData(td_stable.ev, td_stable.N, orc.SORT_STD, 4)
    ref = orc.boost_train(td, orc.boost_params(weak_count=8, max_false_alarm=0.0, min_hit_rate=1.0))
    assert_stumps_match(ref.trees, cv2_boost_stumps(X, labels, "GAB", 8, 0.95), 3, True, "GAB std::sort")


def test_oracle_categorical_stumps_match_cv2_boost():
    imgs, labels = make_samples(150, 250, 12, 12)
    n = len(imgs)
    ev = orc.Evaluator(orc.LBP, 0, 12, 12, n)
    ev.set_images(imgs, labels)
    td = orc.TrainData(ev, n, orc.SORT_STABLE, 4)
    ref = orc.boost_train(td, orc.boost_params(weak_count=8, max_false_alarm=0.0, min_hit_rate=1.0))
    cvt = cv2_boost_stumps(td.fval().T.copy(), labels, "GAB", 8, 0.95, categorical=True)
    assert_stumps_match(ref.trees, cvt, 3, True, "GAB LBP")


def test_discrete_boost_differs_only_by_documented_criterion(haar12):
    # cv::ml 4.x always splits classification nodes by Gini; the legacy DAB default is MISCLASS (src/o_cvboost.cpp:84-92).
    # The leaf weights of the first tree still follow the same formula: +-log((1-err)/err) (src/boost.cpp:284-317).
    td, labels, X = haar12
    ref = orc.boost_train(td, orc.boost_params(boost_type=orc.DAB, weak_count=2, max_false_alarm=0.0, min_hit_rate=1.0))
    r = ref.trees[0]
    lv, rv = r[r[0]["left"]]["value"], r[r[0]["right"]]["value"]
    assert lv == -rv and abs(lv) > 0
    # misclassification rate of the oracle's first stump reproduces its own leaf weight
    f = X[:, r[0]["var_idx"]]
    pred = np.where(f <= np.float32(r[0]["c"]), np.sign(lv), np.sign(rv))
    err = float(np.mean(pred != (2 * labels.astype(np.int32) - 1)))
    assert abs(abs(lv) - np.log((1 - err) / err)) < 1e-9
