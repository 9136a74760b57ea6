"""Host-side mirror of CvCascadeBoostParams / CvCascadeBoost over the C ABI
(traincascade/lib/include/boost.h:30-106, traincascade/lib/src/boost.cpp:41-166,409-573).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CascadeCudaError, check
from .evaluator import _p


class CvCascadeBoostParams:
    """boost_type: 0 DAB, 1 RAB, 2 LB, 3 GAB (cv::ml::Boost)."""
    DISCRETE, REAL, LOGIT, GENTLE = 0, 1, 2, 3
    _NAMES = {0: "DAB", 1: "RAB", 2: "LB", 3: "GAB"}

    def __init__(self, boostType=3, minHitRate=0.995, maxFalseAlarm=0.5, weightTrimRate=0.95, maxDepth=1,
                 maxWeakCount=100):
        # NB the reference's 6-argument constructor forces GENTLE (src/boost.cpp:47-56); the type is
        # only changed by scanAttr("-bt") / read().  The mirror keeps the caller's value, which is what
        # the CLI path (-bt) produces.
        self.boost_type = boostType
        self.minHitRate = float(np.float32(minHitRate))
        self.maxFalseAlarm = float(np.float32(maxFalseAlarm))
        self.weight_trim_rate = weightTrimRate
        self.max_depth = maxDepth
        self.weak_count = maxWeakCount
        self.min_sample_count = 10          # CvDTreeParams default (o_cvdtreeparams.cpp:7)
        self.regression_accuracy = 0.01     # CvDTreeParams default (o_cvdtreeparams.cpp:12)

    def boostTypeStr(self):
        return self._NAMES[self.boost_type]

    def to_c(self):
        return _lib.BoostParams(self.boost_type, self.weak_count, self.max_depth, self.min_sample_count,
                                self.weight_trim_rate, self.regression_accuracy, self.minHitRate, self.maxFalseAlarm)


def tree_to_list(tree):
    return [dict(left=n.left, right=n.right, var_idx=n.var_idx, split_point=n.split_point, c=n.c, quality=n.quality,
                 subset=list(n.subset), value=n.value, depth=n.depth, sample_count=n.sample_count)
            for n in tree.nodes[:tree.num_nodes]]


class CvCascadeBoost:
    """One boosted stage.  train() runs K2 (precalc), K3 (sort) and the K4 loop on the GPU."""

    def __init__(self):
        self.threshold = 0.0
        self.trees = []          # list of trees; a tree is a list of node dicts, node 0 = root
        self.log = []            # (N, HR, FA) rows of the reference's stdout table
        self.evaluator = None
        self._cstage = None
        self._keep = None

    # bool train(const CvFeatureEvaluator*, int numSamples, int precalcValBufSize, int precalcIdxBufSize, params)
    def train(self, featureEvaluator, numSamples, precalcValBufSize, precalcIdxBufSize, params, verbose=False):
        if precalcValBufSize < 0 or precalcIdxBufSize < 0:
            raise CascadeCudaError("_numPrecalcVal and _numPrecalcIdx must be positive or 0")
        self.evaluator = featureEvaluator
        ctx = featureEvaluator.ctx
        cap = max(int(params.weak_count), 1)
        trees = (_lib.Tree * cap)()
        log = (_lib.LogRow * cap)()
        st = _lib.Stage(0.0, 0, 0, 0, trees, log)
        cp = params.to_c()
        check(ctx.lib.cc_boost_train(ctx.h, int(numSamples), C.byref(cp), C.byref(st)), "cc_boost_train")
        self._keep = (trees, log)
        self._cstage = st
        self.threshold = float(st.threshold)
        self.trees = [tree_to_list(trees[t]) for t in range(st.num_weak)]
        self.log = [(log[i].n, float(log[i].hit_rate), float(log[i].false_alarm)) for i in range(st.num_log)]
        if verbose:
            print("+----+---------+---------+\n|  N |    HR   |    FA   |\n+----+---------+---------+")
            for n, hr, fa in self.log:
                print("|%4d|%9.6g|%9.6g|\n+----+---------+---------+" % (n, hr, fa))
        return bool(st.trained)

    @staticmethod
    def from_trees(threshold, trees):
        s = CvCascadeBoost()
        s.threshold = float(threshold)
        s.trees = trees
        s._build_c()
        return s

    # bool read(const FileNode&, const CvFeatureEvaluator*, const CvCascadeBoostParams&), src/boost.cpp:543-565 +
    # CvCascadeBoostTree::read, src/o_cvcascadeboosttree.cpp:95-165 (resume: leaf values and thresholds come back as f32)
    @staticmethod
    def read(node, maxCatCount):
        """node = cv2.FileNode of one stage (stage<i>.xml top level, or an element of cascade.xml's `stages`)."""
        subsetN = (maxCatCount + 31) // 32
        step = 3 + (subsetN if maxCatCount > 0 else 1)
        threshold = float(np.float32(node.getNode("stageThreshold").real()))
        trees = []
        weak = node.getNode("weakClassifiers")
        for wi in range(weak.size()):
            wn = weak.at(wi)
            inn, lvn = wn.getNode("internalNodes"), wn.getNode("leafValues")
            vals = [inn.at(i).real() for i in range(inn.size())]
            leaves = [float(np.float32(lvn.at(i).real())) for i in range(lvn.size())]
            n_int = len(vals) // step
            if n_int == 0 or len(vals) != n_int * step:
                raise CascadeCudaError("malformed internalNodes in stage file")
            nodes = [None] * n_int
            for i in range(n_int):
                rec = vals[i * step:(i + 1) * step]
                nd = dict(left=-1, right=-1, var_idx=int(rec[2]), split_point=-1, c=0.0, quality=0.0, subset=[0] * 8, value=0.0,
                          depth=0, sample_count=0)
                if maxCatCount > 0:
                    sub = [int(v) for v in rec[3:3 + subsetN]]
                    nd["subset"] = (sub + [0] * 8)[:8]
                else:
                    nd["c"] = float(np.float32(rec[3]))
                nodes[i] = nd
            for i in range(n_int):
                for side, ref in (("left", int(vals[i * step])), ("right", int(vals[i * step + 1]))):
                    if ref <= 0:            # leaf: index -ref into leafValues
                        nodes.append(dict(left=-1, right=-1, var_idx=-1, split_point=-1, c=0.0, quality=0.0, subset=[0] * 8,
                                          value=leaves[-ref], depth=0, sample_count=0))
                        nodes[i][side] = len(nodes) - 1
                    else:
                        nodes[i][side] = ref
            trees.append(nodes)
        return CvCascadeBoost.from_trees(threshold, trees)

    def _build_c(self):
        cap = max(len(self.trees), 1)
        trees = (_lib.Tree * cap)()
        log = (_lib.LogRow * cap)()
        for t, tr in enumerate(self.trees):
            trees[t].num_nodes = len(tr)
            for i, nd in enumerate(tr):
                n = trees[t].nodes[i]
                n.left, n.right, n.var_idx = nd["left"], nd["right"], nd["var_idx"]
                n.split_point = nd.get("split_point", -1)
                n.c, n.quality = nd.get("c", 0.0), nd.get("quality", 0.0)
                for j in range(8):
                    n.subset[j] = nd.get("subset", [0] * 8)[j]
                n.value = nd["value"]
                n.depth, n.sample_count = nd.get("depth", 0), nd.get("sample_count", 0)
        self._keep = (trees, log)
        self._cstage = _lib.Stage(self.threshold, len(self.trees), 0, 1, trees, log)

    def c_stage(self):
        if self._cstage is None:
            self._build_c()
        return self._cstage

    def getThreshold(self):
        return self.threshold

    # float predict(int sampleIdx, bool returnSum=false)
    def predict(self, featureEvaluator, sampleIdx, returnSum=False, count=1):
        ctx = featureEvaluator.ctx
        out = np.zeros(count, np.float32)
        st = self.c_stage()
        check(ctx.lib.cc_stage_predict(ctx.h, C.byref(st), int(sampleIdx), int(count), int(returnSum), _p(out)),
              "cc_stage_predict")
        return out if count > 1 else float(out[0])

    def markUsedFeaturesInMap(self, featureMap):
        for tr in self.trees:
            for nd in tr:
                if nd["left"] >= 0:
                    featureMap[nd["var_idx"]] = 1

    # void write(FileStorage&, const Mat& featureMap), src/boost.cpp:520-532 + o_cvcascadeboosttree.cpp:41-93
    def write(self, fs, featureMap, maxCatCount):
        fs.write("maxWeakCount", len(self.trees))
        fs.write("stageThreshold", float(np.float32(self.threshold)))
        fs.start_seq("weakClassifiers")
        subsetN = (maxCatCount + 31) // 32
        for tr in self.trees:
            internal, leaves = [], []
            queue = [0]
            internalIdx, leafIdx = 1, 0
            while queue:
                nd = tr[queue.pop(0)]
                for child in (nd["left"], nd["right"]):
                    ch = tr[child]
                    if ch["left"] < 0:
                        leaves.append(float(np.float32(ch["value"])))
                        internal.append(leafIdx)
                        leafIdx -= 1
                    else:
                        queue.append(child)
                        internal.append(internalIdx)
                        internalIdx += 1
                fidx = nd["var_idx"] if featureMap is None else int(featureMap[nd["var_idx"]])
                internal.append(fidx)
                if not maxCatCount:
                    internal.append(float(np.float32(nd["c"])))
                else:
                    internal.extend(int(v) for v in nd["subset"][:subsetN])
            fs.start_map(None)
            fs.write_flow_seq("internalNodes", internal)
            fs.write_flow_seq("leafValues", leaves)
            fs.end_map()
        fs.end_seq()
