"""Host-side mirror of CvCascadeParams / CvCascadeClassifier (training half) over the C ABI
(traincascade/lib/include/cascadeclassifier.h:84-204, traincascade/lib/src/cascadeclassifier.cpp:24-578).

The stage loop, the stdout text and the files written (params.xml, stage<i>.xml, cascade.xml) follow the
reference; the two hot calls — updateTrainingSet (mining) and CvCascadeBoost::train — run on the GPU.
"""
import ctypes as C
import math
import os
import sys
import time

import numpy as np

from . import _lib, synth
from ._lib import CascadeCudaError, check
from .boost import CvCascadeBoost, CvCascadeBoostParams
from .evaluator import Context, CvFeatureEvaluator, CvFeatureParams, CvHaarFeatureParams, _p
from .xml_io import XmlWriter


class CvCascadeParams:
    BOOST = 0
    defaultStageType = BOOST
    defaultFeatureType = CvFeatureParams.HAAR
    _FT = {0: "HAAR", 1: "LBP", 2: "HOG"}

    def __init__(self, stageType=0, featureType=0, winSize=(24, 24)):
        self.stageType = stageType
        self.featureType = featureType
        self.winSize = (int(winSize[0]), int(winSize[1]))     # (width, height)
        self.name = "cascadeParams"

    def write(self, fs):      # src/cascadeclassifier.cpp:33-46
        fs.write("stageType", "BOOST")
        fs.write("featureType", self._FT[self.featureType])
        fs.write("height", self.winSize[1])
        fs.write("width", self.winSize[0])

    def printAttrs(self, out):
        out.write("stageType: BOOST\n")
        out.write("featureType: %s\n" % self._FT[self.featureType])
        out.write("sampleWidth: %d\n" % self.winSize[0])
        out.write("sampleHeight: %d\n" % self.winSize[1])


def _fmt(v):
    """std::cout << float/double with default precision 6."""
    return "%g" % v


def load_backgrounds(bg_filename):
    """CvCascadeImageReader::NegReader::create (src/imagestorage.cpp:35-55) + cv::imread(path, 0)."""
    names = []
    with open(bg_filename, "r") as f:
        for line in f:
            s = line.rstrip(" \n\r\t")
            if not s:
                break
            if s[0] == "#":
                continue
            names.append(s)
    images = []
    for n in names:
        if n.lower().endswith(".pgm"):
            images.append(synth.read_pgm(n))
        else:
            import cv2
            im = cv2.imread(n, 0)
            images.append(im if im is not None else np.zeros((0, 0), np.uint8))
    return images


class CvCascadeClassifier:
    def __init__(self, ctx=None, out=None):
        self.ctx = ctx
        self.out = out or sys.stdout
        self.stageClassifiers = []
        self.featureEvaluator = None
        self.cascadeParams = None
        self.featureParams = None
        self.stageParams = None
        self.numPos = self.numNeg = self.numStages = self.curNumSamples = 0
        self.reports = []
        self.timings = []
        self._reader = _lib.ReaderState()
        self._vec_pos = C.c_int32(0)
        self._vec = None

    # -- data sources (imgReader.create) --
    def _create_readers(self, positives, backgrounds):
        W, H = self.cascadeParams.winSize
        self._vec = np.ascontiguousarray(positives, np.uint8)
        if self._vec.ndim != 3 or self._vec.shape[1:] != (H, W):
            raise CascadeCudaError("vec samples must be [count,%d,%d]" % (H, W))
        lib, h = self.ctx.lib, self.ctx.h
        check(lib.cc_bg_clear(h), "cc_bg_clear")
        for im in backgrounds:
            im = np.ascontiguousarray(im, np.uint8)
            if im.size == 0:
                # unreadable file: the reference keeps it in imgFilenames and nextImg() skips it without touching `round`
                # (src/imagestorage.cpp:62-70), so it must keep its slot in the list
                check(lib.cc_bg_register(h, None, 0, 0), "cc_bg_register")
                continue
            check(lib.cc_bg_register(h, _p(im), im.shape[1], im.shape[0]), "cc_bg_register")
        self._reader = _lib.ReaderState()
        self._vec_pos = C.c_int32(0)

    def _sync_cascade(self):
        n = len(self.stageClassifiers)
        arr = (_lib.Stage * max(n, 1))()
        for i, st in enumerate(self.stageClassifiers):
            cs = st.c_stage()
            arr[i] = cs
        self._stage_arr = arr
        check(self.ctx.lib.cc_cascade_set_stages(self.ctx.h, arr, n), "cc_cascade_set_stages")

    # -- fillPassedSamples (src/cascadeclassifier.cpp:329-357) --
    def fillPassedSamples(self, first, count, isPositive, minimumAcceptanceRatio):
        lib, h = self.ctx.lib, self.ctx.h
        consumed, got = C.c_int64(0), C.c_int32(0)
        if isPositive:
            check(lib.cc_mine_positives(h, _p(self._vec), self._vec.shape[0], int(count), C.byref(self._vec_pos),
                                        C.byref(consumed), C.byref(got)), "cc_mine_positives")
        else:
            check(lib.cc_mine_negatives(h, int(first), int(count), float(minimumAcceptanceRatio),
                                        C.byref(self._reader), C.byref(consumed), C.byref(got)), "cc_mine_negatives")
        return got.value, consumed.value

    # -- updateTrainingSet (src/cascadeclassifier.cpp:308-327) --
    def updateTrainingSet(self, minimumAcceptanceRatio):
        self._vec_pos = C.c_int32(0)          # imgReader.restart(): rewinds the .vec only
        self._sync_cascade()
        posCount, posConsumed = self.fillPassedSamples(0, self.numPos, True, 0)
        rep = dict(pos_count=posCount, pos_consumed=posConsumed, neg_count=0, neg_consumed=0, acceptance_ratio=0.0)
        if not posCount:
            return False, 0.0, rep
        self.out.write("POS count : consumed   %d : %d\n" % (posCount, int(posConsumed)))
        proNumNeg = int(np.rint((float(self.numNeg) * float(posCount)) / self.numPos))      # cvRound
        negCount, negConsumed = self.fillPassedSamples(posCount, proNumNeg, False, minimumAcceptanceRatio)
        rep.update(neg_count=negCount, neg_consumed=negConsumed)
        if not negCount:
            if not (negConsumed > 0 and (float(negCount) + 1) / float(negConsumed) <= minimumAcceptanceRatio):
                return False, 0.0, rep
        self.curNumSamples = posCount + negCount
        acceptanceRatio = 0.0 if negConsumed == 0 else float(negCount) / float(negConsumed)
        rep["acceptance_ratio"] = acceptanceRatio
        self.out.write("NEG count : acceptanceRatio    %d : %s\n" % (negCount, _fmt(acceptanceRatio)))
        return True, acceptanceRatio, rep

    # -- train (src/cascadeclassifier.cpp:137-295) --
    def train(self, cascadeDirName, posFilename, negFilename, numPos, numNeg, precalcValBufSize, precalcIdxBufSize,
              numStages, cascadeParams, featureParams, stageParams, baseFormatSave=False,
              acceptanceRatioBreakValue=-1.0, positives=None, backgrounds=None):
        t_start = time.time()
        out = self.out
        if not cascadeDirName or (positives is None and not posFilename) or (backgrounds is None and not negFilename):
            raise CascadeCudaError("_cascadeDirName or _bgfileName or _vecFileName is NULL")
        dirName = cascadeDirName if cascadeDirName[-1] in "/\\" else cascadeDirName + "/"
        self.numPos, self.numNeg, self.numStages = numPos, numNeg, numStages
        self.cascadeParams, self.featureParams, self.stageParams = cascadeParams, featureParams, stageParams
        W, H = cascadeParams.winSize
        try:        # imgReader.create comes first in the reference too (src/cascadeclassifier.cpp:163-168)
            if positives is None:
                positives = synth.read_vec(posFilename, W, H)
            if backgrounds is None:
                backgrounds = load_backgrounds(negFilename)
        except (IOError, OSError, ValueError):
            out.write("Image reader can not be created from -vec %s and -bg %s.\n" % (posFilename, negFilename))
            return False
        if self.ctx is None:
            self.ctx = Context(0)
        self.stageClassifiers = []
        if not self.load(dirName):
            self.cascadeParams, self.featureParams, self.stageParams = cascadeParams, featureParams, stageParams
            self.featureEvaluator = CvFeatureEvaluator.create(cascadeParams.featureType, self.ctx)
            self.featureEvaluator.init(featureParams, numPos + numNeg, cascadeParams.winSize)
        else:
            # src/cascadeclassifier.cpp:179-186
            out.write("---------------------------------------------------------------------------------\n"
                      "Training parameters are pre-loaded from the parameter file in data folder!\n"
                      "Please empty this folder if you want to use a NEW set of training parameters.\n"
                      "---------------------------------------------------------------------------------\n")
            cascadeParams, featureParams, stageParams = self.cascadeParams, self.featureParams, self.stageParams
            W, H = cascadeParams.winSize
        self._create_readers(positives, backgrounds)
        out.write("PARAMETERS:\ncascadeDirName: %s\nvecFileName: %s\nbgFileName: %s\n" % (cascadeDirName, posFilename, negFilename))
        out.write("numPos: %d\nnumNeg: %d\nnumStages: %d\n" % (numPos, numNeg, numStages))
        out.write("precalcValBufSize[Mb] : %d\nprecalcIdxBufSize[Mb] : %d\n" % (precalcValBufSize, precalcIdxBufSize))
        out.write("acceptanceRatioBreakValue : %s\n" % _fmt(acceptanceRatioBreakValue))
        cascadeParams.printAttrs(out)
        out.write("boostType: %s\nminHitRate: %s\nmaxFalseAlarmRate: %s\nweightTrimRate: %s\nmaxDepth: %d\nmaxWeakCount: %d\n" % (
            stageParams.boostTypeStr(), _fmt(stageParams.minHitRate), _fmt(stageParams.maxFalseAlarm),
            _fmt(stageParams.weight_trim_rate), stageParams.max_depth, stageParams.weak_count))
        if isinstance(featureParams, CvHaarFeatureParams):
            out.write("mode: %s\n" % featureParams.modeStr())
        out.write("Number of unique features given windowSize [%d,%d] : %d\n" % (W, H, self.featureEvaluator.getNumFeatures()))

        startNumStages = len(self.stageClassifiers)
        if startNumStages > 1:
            out.write("\nStages 0-%d are loaded\n" % (startNumStages - 1))
        elif startNumStages == 1:
            out.write("\nStage 0 is loaded\n")
        requiredLeafFARate = math.pow(float(stageParams.maxFalseAlarm), float(numStages)) / float(stageParams.max_depth)
        for i in range(startNumStages, numStages):
            out.write("\n===== TRAINING %d-stage =====\n<BEGIN\n" % i)
            t0 = time.time()
            ok, tempLeafFARate, rep = self.updateTrainingSet(requiredLeafFARate)
            t1 = time.time()
            self.reports.append(rep)
            if not ok:
                out.write("Train dataset for temp stage can not be filled. Branch training terminated.\n")
                break
            if tempLeafFARate <= requiredLeafFARate:
                out.write("Required leaf false alarm rate achieved. Branch training terminated.\n")
                break
            if tempLeafFARate <= acceptanceRatioBreakValue and acceptanceRatioBreakValue >= 0:
                out.write("The required acceptanceRatio for the model has been reached to avoid overfitting of "
                          "trainingdata. Branch training terminated.\n")
                break
            tempStage = CvCascadeBoost()
            out.write("+----+---------+---------+\n|  N |    HR   |    FA   |\n+----+---------+---------+\n")
            isStageTrained = tempStage.train(self.featureEvaluator, self.curNumSamples, precalcValBufSize,
                                             precalcIdxBufSize, stageParams)
            for n, hr, fa in tempStage.log:
                out.write("|%4d|%9s|%9s|\n+----+---------+---------+\n" % (n, _fmt(hr), _fmt(fa)))
            out.write("END>\n")
            t2 = time.time()
            self.timings.append(dict(stage=i, mine_s=t1 - t0, train_s=t2 - t1, weak=len(tempStage.trees),
                                     neg_consumed=rep["neg_consumed"]))
            if not isStageTrained:
                break
            self.stageClassifiers.append(tempStage)
            os.makedirs(dirName, exist_ok=True)
            if i == 0:
                fs = XmlWriter(dirName + "params.xml")
                fs.start_map("params")
                self.writeParams(fs)
                fs.end_map()
                fs.release()
            fs = XmlWriter(dirName + "stage%d.xml" % i)
            fs.start_map("stage%d" % i)
            tempStage.write(fs, None, self.featureEvaluator.getMaxCatCount())
            fs.end_map()
            fs.release()
            seconds = int(time.time() - t_start)
            out.write("Training until now has taken %d days %d hours %d minutes %d seconds.\n" % (
                seconds // 86400, (seconds // 3600) % 24, (seconds // 60) % 60, seconds % 60))
        if not self.stageClassifiers:
            out.write("Cascade classifier can't be trained. Check the used training parameters.\n")
            return False
        self.save(dirName + "cascade.xml", baseFormatSave)
        return True

    # bool load(const string& cascadeDirName), src/cascadeclassifier.cpp:534-564 (+ readParams :388-404): resume from
    # params.xml and the stage<i>.xml files already in the directory; the NegReader state is not persisted (as in the reference)
    def load(self, cascadeDirName):
        import cv2
        from .evaluator import CvLBPFeatureParams, CvHOGFeatureParams
        ppath = cascadeDirName + "params.xml"
        if not os.path.exists(ppath):
            return False
        fs = cv2.FileStorage(ppath, cv2.FILE_STORAGE_READ)
        if not fs.isOpened():
            return False
        node = fs.getFirstTopLevelNode()
        try:
            ft = {"HAAR": 0, "LBP": 1, "HOG": 2}[node.getNode("featureType").string()]
            if node.getNode("stageType").string() != "BOOST":
                return False
            self.cascadeParams = CvCascadeParams(0, ft, (int(node.getNode("width").real()), int(node.getNode("height").real())))
            sp = node.getNode("stageParams")
            bt = {"DAB": 0, "RAB": 1, "LB": 2, "GAB": 3}[sp.getNode("boostType").string()]
            self.stageParams = CvCascadeBoostParams(bt, float(np.float32(sp.getNode("minHitRate").real())),
                                                    float(np.float32(sp.getNode("maxFalseAlarm").real())), sp.getNode("weightTrimRate").real(),
                                                    int(sp.getNode("maxDepth").real()), int(sp.getNode("maxWeakCount").real()))
            fp = node.getNode("featureParams")
            if ft == 0:
                self.featureParams = CvHaarFeatureParams({"BASIC": 0, "CORE": 1, "ALL": 2}[fp.getNode("mode").string()])
            elif ft == 1:
                self.featureParams = CvLBPFeatureParams()
            else:
                self.featureParams = CvHOGFeatureParams()
        except KeyError:
            fs.release()
            return False
        fs.release()
        if self.ctx is None:
            self.ctx = Context(0)
        self.featureEvaluator = CvFeatureEvaluator.create(self.cascadeParams.featureType, self.ctx)
        self.featureEvaluator.init(self.featureParams, self.numPos + self.numNeg, self.cascadeParams.winSize)
        for si in range(self.numStages):
            spath = cascadeDirName + "stage%d.xml" % si
            if not os.path.exists(spath):
                break
            fs = cv2.FileStorage(spath, cv2.FILE_STORAGE_READ)
            if not fs.isOpened():
                break
            try:
                st = CvCascadeBoost.read(fs.getFirstTopLevelNode(), self.featureEvaluator.getMaxCatCount())
            except (CascadeCudaError, cv2.error):
                fs.release()
                break
            fs.release()
            self.stageClassifiers.append(st)
        return True

    # -baseFormatSave: the legacy "opencv-haar-classifier" layout, Haar only (src/cascadeclassifier.cpp:457-531)
    def _save_base_format(self, filename):
        if self.cascadeParams.featureType != CvFeatureParams.HAAR:
            raise CascadeCudaError("old file format is used for Haar-like features only")
        fs = XmlWriter(filename)
        fs.start_map("cascade", type_name="opencv-haar-classifier", flow=True)
        fs.write_flow_seq("size", [int(self.cascadeParams.winSize[0]), int(self.cascadeParams.winSize[1])])
        fs.start_seq("stages")
        for si, st in enumerate(self.stageClassifiers):
            fs.start_map(None)
            fs.start_seq("trees")
            for tr in st.trees:
                fs.start_seq(None)
                queue = [0]
                total_inner = 0
                while queue:
                    nd = tr[queue.pop(0)]
                    fs.start_map(None)
                    fs.start_map("feature")
                    rects, wts, tilted = self.featureEvaluator.feature(nd["var_idx"])
                    fs.start_seq("rects")
                    for ri in range(3):
                        if rects[ri][2] == 0:
                            break
                        fs.write_flow_seq(None, [int(rects[ri][0]), int(rects[ri][1]), int(rects[ri][2]), int(rects[ri][3]), float(wts[ri])])
                    fs.end_seq()
                    fs.write("tilted", int(tilted))
                    fs.end_map()
                    fs.write("threshold", float(np.float32(nd["c"])))
                    for side, node_key, val_key in (("left", "left_node", "left_val"), ("right", "right_node", "right_val")):
                        ch = tr[nd[side]]
                        if ch["left"] >= 0:
                            queue.append(nd[side])
                            total_inner += 1
                            fs.write(node_key, total_inner)
                        else:
                            fs.write(val_key, float(ch["value"]))
                    fs.end_map()
                fs.end_seq()
            fs.end_seq()
            fs.write("stage_threshold", float(np.float32(st.threshold)))
            fs.write("parent", si - 1)
            fs.write("next", -1)
            fs.end_map()
        fs.end_seq()
        fs.end_map()
        fs.release()

    def predict(self, sampleIdx):
        for st in self.stageClassifiers:
            if st.predict(self.featureEvaluator, sampleIdx) == 0.0:
                return 0
        return 1

    # -- XML (src/cascadeclassifier.cpp:359-386, 439-456, 566-578) --
    def writeParams(self, fs):
        self.cascadeParams.write(fs)
        sp = self.stageParams
        fs.start_map("stageParams")
        fs.write("boostType", sp.boostTypeStr())
        fs.write("minHitRate", float(np.float32(sp.minHitRate)))
        fs.write("maxFalseAlarm", float(np.float32(sp.maxFalseAlarm)))
        fs.write("weightTrimRate", float(sp.weight_trim_rate))
        fs.write("maxDepth", int(sp.max_depth))
        fs.write("maxWeakCount", int(sp.weak_count))
        fs.end_map()
        fp = self.featureParams
        fs.start_map("featureParams")
        fs.write("maxCatCount", int(fp.maxCatCount))
        fs.write("featSize", int(fp.featSize))
        if isinstance(fp, CvHaarFeatureParams):
            fs.write("mode", fp.modeStr())
        fs.end_map()

    def getUsedFeaturesIdxMap(self):
        varCount = self.featureEvaluator.getNumFeatures() * self.featureEvaluator.getFeatureSize()
        fmap = np.full(varCount, -1, np.int32)
        for st in self.stageClassifiers:
            st.markUsedFeaturesInMap(fmap)
        idx = 0
        for fi in range(varCount):
            if fmap[fi] >= 0:
                fmap[fi] = idx
                idx += 1
        return fmap

    def save(self, filename, baseFormat=False):
        if baseFormat:
            return self._save_base_format(filename)
        fs = XmlWriter(filename)
        fmap = self.getUsedFeaturesIdxMap()
        fs.start_map("cascade")
        self.writeParams(fs)
        fs.write("stageNum", len(self.stageClassifiers))
        fs.start_seq("stages")
        for i, st in enumerate(self.stageClassifiers):
            fs.comment("stage %d" % i)
            fs.start_map(None)
            st.write(fs, fmap, self.featureEvaluator.getMaxCatCount())
            fs.end_map()
        fs.end_seq()
        self.featureEvaluator.writeFeatures(fs, fmap)
        fs.end_map()
        fs.release()
