"""Independent pin for stage-training results: OpenCV's own cv2.ml.Boost (4.x `DTreesImplForBoost`), the maintained descendant of
the legacy CvBoost / CvBoostTree classes the reference vendors (src/o_cvboost.cpp, src/o_cvboostree.cpp:361-516, src/boost.cpp:168-407).
It is written by other people than the oracle restatement, ships in the image as a binary wheel, and shares the legacy algorithm for

  * GENTLE / LOGIT boosting on ordered variables  (find_split_ord_reg, calc_node_value regression, update_weights GAB / LB, trim_weights)
  * REAL boosting on ordered variables            (find_split_ord_class with the Gini criterion, update_weights RAB)
  * GENTLE boosting on categorical variables      (find_split_cat_reg: categories ordered by mean response)

Not comparable (documented differences of cv::ml 4.x, not of the restatement): DISCRETE boosting — the legacy code defaults its split
criterion to MISCLASS (src/o_cvboost.cpp:84-92) while 4.x always uses Gini; categorical classification splits — 4.x orders the
categories differently; the `quality` scale of classification splits.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def cv2_boost_stumps(X, y, boost_type, weak_count, trim_rate, categorical=False):
    """Trains cv2.ml.Boost stumps on the fval matrix X [N][F] (float32) / labels y; returns [(var, threshold | None, quality, left, right)]."""
    import cv2
    kinds = {"GAB": cv2.ml.BOOST_GENTLE, "RAB": cv2.ml.BOOST_REAL, "LB": cv2.ml.BOOST_LOGIT, "DAB": cv2.ml.BOOST_DISCRETE}
    m = cv2.ml.Boost_create()
    m.setBoostType(kinds[boost_type])
    m.setWeakCount(int(weak_count))
    m.setWeightTrimRate(float(trim_rate))
    m.setMaxDepth(1)
    m.setUseSurrogates(False)
    m.setMinSampleCount(10)          # CvDTreeParams default, src/o_cvdtreeparams.cpp:7
    m.setCVFolds(0)
    m.setRegressionAccuracy(0.01)    # src/o_cvdtreeparams.cpp:12
    X = np.ascontiguousarray(X, np.float32)
    y = np.ascontiguousarray(y, np.int32)
    if categorical:
        m.setMaxCategories(256)
        vt = np.full(X.shape[1] + 1, cv2.ml.VAR_CATEGORICAL, np.uint8)
        m.train(cv2.ml.TrainData_create(X, cv2.ml.ROW_SAMPLE, y, varType=vt))
    else:
        m.train(X, cv2.ml.ROW_SAMPLE, y)
    fs = cv2.FileStorage("boost.yml", cv2.FILE_STORAGE_WRITE | cv2.FILE_STORAGE_MEMORY)
    m.write(fs)
    fr = cv2.FileStorage(fs.releaseAndGetString(), cv2.FILE_STORAGE_READ | cv2.FILE_STORAGE_MEMORY)
    trees = fr.getNode("trees")
    out = []
    for ti in range(trees.size()):
        nodes = trees.at(ti).getNode("nodes")
        if nodes.size() != 3:
            break                                        # the root did not split
        sp = nodes.at(0).getNode("splits").at(0)
        le = sp.getNode("le")
        thr = None if le.empty() else float(np.float32(le.real()))
        out.append((int(sp.getNode("var").real()), thr, float(np.float32(sp.getNode("quality").real())),
                    nodes.at(1).getNode("value").real(), nodes.at(2).getNode("value").real()))
    return out


def assert_stumps_match(trees, cv_trees, min_trees, check_quality, what=""):
    """trees = oracle / GPU stage trees (lists of node dicts); cv_trees = cv2_boost_stumps output."""
    k = min(len(trees), len(cv_trees))
    assert k >= min_trees, (what, len(trees), len(cv_trees))
    for ti in range(k):
        r, (var, thr, q, lv, rv) = trees[ti], cv_trees[ti]
        root = r[0]
        assert root["var_idx"] == var, (what, ti, root, cv_trees[ti])
        if thr is not None:
            assert np.float32(root["c"]) == np.float32(thr), (what, ti, root["c"], thr)
        if check_quality:
            assert abs(root["quality"] - q) <= 2e-6 * abs(q), (what, ti, root["quality"], q)
        for mine, ref in ((r[root["left"]]["value"], lv), (r[root["right"]]["value"], rv)):
            assert abs(mine - ref) <= 1e-6 * max(1.0, abs(ref)), (what, ti, mine, ref)
    return k
