"""Host-side mirror of the reference's feature-evaluator interface over the C ABI.

Same names, argument meaning and error behaviour as
  CvFeatureParams / CvHaarFeatureParams / CvLBPFeatureParams / CvHOGFeatureParams and
  CvFeatureEvaluator::{create, init, setImage, operator(), writeFeatures, getNumFeatures,
  getMaxCatCount, getFeatureSize, getCls}
(traincascade/lib/include/traincascade_features.h:129-188, haarfeatures.h:31-101,
 lbpfeatures.h:20-66, HOGfeatures.h:28-83), with cv::Mat replaced by numpy arrays.  All arithmetic
runs in libcascade_cuda.so; nothing here computes features on the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CascadeCudaError, check


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One cc_ctx == one GPU (one process rank)."""

    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.cc_create(int(device), C.byref(h)), "cc_create")
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, cuda_stream_ptr):
        check(self.lib.cc_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "cc_set_stream")

    def synchronize(self):
        check(self.lib.cc_synchronize(self.h), "cc_synchronize")

    def timing_reset(self):
        check(self.lib.cc_timing_reset(self.h), "cc_timing_reset")

    def timing(self, name):
        ms, n = C.c_double(0), C.c_int64(0)
        check(self.lib.cc_timing_get(self.h, name.encode(), C.byref(ms), C.byref(n)), "cc_timing_get")
        return ms.value, n.value

    def launch_count(self):
        return int(self.lib.cc_launch_count(self.h))

    def set_search_mode(self, mode):
        """0 = pruned split search (default), 1 = exhaustive (every variable searched exactly)."""
        check(self.lib.cc_boost_set_search_mode(self.h, int(mode)), "cc_boost_set_search_mode")

    def set_sort_mode(self, mode):
        """0 = auto (Haar: bucket + rank sort, HOG: radix passes), 1 = radix passes, 2 = bucket + rank."""
        check(self.lib.cc_boost_set_sort_mode(self.h, int(mode)), "cc_boost_set_sort_mode")

    def sort_deferred_rows(self):
        """Rows of the last precalc that the bucket sort handed to the radix kernel."""
        a = C.c_int64(0)
        check(self.lib.cc_boost_get_sort_stats(self.h, C.byref(a)), "cc_boost_get_sort_stats")
        return a.value

    def search_stats(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(self.lib.cc_boost_get_search_stats(self.h, C.byref(a), C.byref(b), C.byref(c)), "cc_boost_get_search_stats")
        return dict(searches=a.value, seed_rows=b.value, flagged_rows=c.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CvFeatureParams:
    HAAR, LBP, HOG = 0, 1, 2

    def __init__(self):
        self.maxCatCount = 0
        self.featSize = 1
        self.name = "featureParams"

    @staticmethod
    def create(featureType):
        if featureType == CvFeatureParams.HAAR:
            return CvHaarFeatureParams()
        if featureType == CvFeatureParams.LBP:
            return CvLBPFeatureParams()
        if featureType == CvFeatureParams.HOG:
            return CvHOGFeatureParams()
        return None


class CvHaarFeatureParams(CvFeatureParams):
    BASIC, CORE, ALL = 0, 1, 2
    _NAMES = {0: "BASIC", 1: "CORE", 2: "ALL"}

    def __init__(self, mode=0):
        super().__init__()
        self.mode = mode
        self.name = "haarFeatureParams"
        self.featureType = CvFeatureParams.HAAR

    def modeStr(self):
        return self._NAMES[self.mode]


class CvLBPFeatureParams(CvFeatureParams):
    def __init__(self):
        super().__init__()
        self.maxCatCount = 256
        self.name = "lbpFeatureParams"
        self.featureType = CvFeatureParams.LBP


class CvHOGFeatureParams(CvFeatureParams):
    def __init__(self):
        super().__init__()
        self.maxCatCount = 0
        self.featSize = 36
        self.name = "HOGFeatureParams"
        self.featureType = CvFeatureParams.HOG


class CvFeatureEvaluator:
    """CvFeatureEvaluator backed by the GPU.  winSize is (width, height) like cv::Size."""

    def __init__(self, featureType, ctx=None):
        self.featureType = featureType
        self.ctx = ctx
        self.featureParams = None
        self.winSize = None
        self.numFeatures = 0
        self._inited = False

    @staticmethod
    def create(featureType, ctx=None):
        if featureType not in (CvFeatureParams.HAAR, CvFeatureParams.LBP, CvFeatureParams.HOG):
            return None
        return CvFeatureEvaluator(featureType, ctx)

    # -- CvFeatureEvaluator::init (src/features.cpp:72-81) --
    def init(self, featureParams, maxSampleCount, winSize):
        if maxSampleCount <= 0:
            raise CascadeCudaError("CV_Assert(_maxSampleCount > 0)")
        if self.ctx is None:
            self.ctx = Context(0)
        self.featureParams = featureParams
        self.winSize = (int(winSize[0]), int(winSize[1]))
        self.maxSampleCount = int(maxSampleCount)
        mode = getattr(featureParams, "mode", 0)
        check(self.ctx.lib.cc_eval_init(self.ctx.h, self.featureType, mode, self.winSize[0], self.winSize[1],
                                        self.maxSampleCount), "cc_eval_init")
        self.numFeatures = self.ctx.lib.cc_eval_num_features(self.ctx.h)
        self._inited = True

    def getNumFeatures(self):
        return self.numFeatures

    def getMaxCatCount(self):
        return self.featureParams.maxCatCount

    def getFeatureSize(self):
        return self.featureParams.featSize

    # -- setImage (src/features.cpp:83-89 + per-family overrides) --
    def setImage(self, img, clsLabel, idx):
        img = np.ascontiguousarray(img, np.uint8)
        W, H = self.winSize
        if img.shape != (H, W):
            raise CascadeCudaError("CV_Assert(img.cols == winSize.width && img.rows == winSize.height)")
        if not (0 <= idx < self.maxSampleCount):
            raise CascadeCudaError("CV_Assert(idx < cls.rows)")
        lab = np.array([clsLabel], np.uint8)
        check(self.ctx.lib.cc_eval_set_images(self.ctx.h, _p(img), _p(lab), int(idx), 1), "cc_eval_set_images")

    def setImages(self, imgs, labels, first=0):
        """Batched setImage: imgs [n,H,W] u8, labels [n] -> slots first..first+n-1."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        labels = np.ascontiguousarray(labels, np.uint8)
        W, H = self.winSize
        if imgs.ndim != 3 or imgs.shape[1:] != (H, W) or labels.shape[0] != imgs.shape[0]:
            raise CascadeCudaError("CV_Assert(img.cols == winSize.width && img.rows == winSize.height)")
        check(self.ctx.lib.cc_eval_set_images(self.ctx.h, _p(imgs), _p(labels), int(first), imgs.shape[0]),
              "cc_eval_set_images")

    # -- operator()(featureIdx, sampleIdx) --
    def __call__(self, featureIdx, sampleIdx):
        out = C.c_float(0)
        check(self.ctx.lib.cc_eval_value(self.ctx.h, int(featureIdx), int(sampleIdx), C.byref(out)), "cc_eval_value")
        return out.value

    def values(self, featureIdx, s0, ns):
        vi = np.ascontiguousarray(featureIdx, np.int32)
        out = np.zeros((len(vi), ns), np.float32)
        check(self.ctx.lib.cc_eval_values(self.ctx.h, _p(vi), len(vi), int(s0), int(ns), _p(out)), "cc_eval_values")
        return out

    def getCls(self, si=None):
        if si is None:
            out = np.zeros(self.maxSampleCount, np.float32)
            check(self.ctx.lib.cc_eval_get_labels(self.ctx.h, 0, self.maxSampleCount, _p(out)), "cc_eval_get_labels")
            return out
        out = np.zeros(1, np.float32)
        check(self.ctx.lib.cc_eval_get_labels(self.ctx.h, int(si), 1, _p(out)), "cc_eval_get_labels")
        return float(out[0])

    # -- parity getters --
    def integrals(self, si):
        P = (self.winSize[0] + 1) * (self.winSize[1] + 1)
        s = np.zeros(P, np.int32)
        t = np.zeros(P, np.int32)
        nf = np.zeros(1, np.float32)
        check(self.ctx.lib.cc_eval_get_integrals(self.ctx.h, int(si), _p(s), _p(t), _p(nf)), "cc_eval_get_integrals")
        return s, t, float(nf[0])

    def hogPlanes(self, si):
        P = (self.winSize[0] + 1) * (self.winSize[1] + 1)
        out = np.zeros((10, P), np.float32)
        check(self.ctx.lib.cc_eval_get_hog_planes(self.ctx.h, int(si), _p(out)), "cc_eval_get_hog_planes")
        return out

    def hogPlanesCompact(self, si):
        """the resident HOG planes: rows / columns 0, 4, 8, ... of hogPlanes() -> [10][H // 4 + 1][W // 4 + 1]"""
        W, H = self.winSize
        out = np.zeros((10, H // 4 + 1, W // 4 + 1), np.float32)
        check(self.ctx.lib.cc_eval_get_hog_planes_compact(self.ctx.h, int(si), _p(out)), "cc_eval_get_hog_planes_compact")
        return out

    def images(self, first, n):
        W, H = self.winSize
        out = np.zeros((n, H, W), np.uint8)
        check(self.ctx.lib.cc_eval_get_images(self.ctx.h, int(first), int(n), _p(out)), "cc_eval_get_images")
        return out

    def feature(self, fi):
        rects = np.zeros(12, np.int32)
        wts = np.zeros(3, np.float32)
        tilted = C.c_int32(0)
        check(self.ctx.lib.cc_eval_get_feature(self.ctx.h, int(fi), _p(rects), _p(wts), C.byref(tilted)),
              "cc_eval_get_feature")
        return rects.reshape(3, 4), wts, int(tilted.value)

    # -- writeFeatures (src/haarfeatures.cpp:116-119,311-320; src/lbpfeatures.cpp:30-33,65-68;
    #    src/HOGfeatures.cpp:49-65,156-160): fs is a cascadeclassifier_b200.xml_io.XmlWriter --
    def writeFeatures(self, fs, featureMap):
        fs.start_seq("features")
        for fi in range(len(featureMap)):
            if featureMap[fi] < 0:
                continue
            fs.start_map(None)
            if self.featureType == CvFeatureParams.HAAR:
                rects, wts, tilted = self.feature(fi)
                fs.start_seq("rects")
                for ri in range(3):
                    if rects[ri][2] == 0:
                        break
                    fs.write_flow_seq(None, [int(rects[ri][0]), int(rects[ri][1]), int(rects[ri][2]),
                                             int(rects[ri][3]), float(wts[ri])])
                fs.end_seq()
                fs.write("tilted", int(tilted))
            elif self.featureType == CvFeatureParams.LBP:
                rects, _, _ = self.feature(fi)
                fs.write_flow_seq("rect", [int(v) for v in rects[0]])
            else:
                fidx, comp = fi // 36, fi % 36
                rects, _, _ = self.feature(fidx)
                fs.write_flow_seq("rect", [int(v) for v in rects[0]] + [int(comp)])
            fs.end_map()
        fs.end_seq()
