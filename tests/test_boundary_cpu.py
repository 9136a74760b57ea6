"""CPU: the C-ABI library loads, exports every symbol declared in include/cascade_cuda.h, fails loudly without a
GPU, and its host-only logic (NegReader state machine, catalogs via the header contract) matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cascade_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cascadeclassifier_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), "libcascade_cuda.so does not export %s" % name
    assert set(_lib.exported_symbols()) <= set(declared)
    assert b"sm_100a" in lib.cc_version()


def test_no_cpu_fallback():
    """Without a CUDA device cc_create must fail (the product path never falls back to the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cascadeclassifier_b200 as ccb
    with pytest.raises(ccb.CascadeCudaError) as e:
        ccb.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cascadeclassifier_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in text and "import orc" not in text and "oracle/" not in text, f


@pytest.mark.parametrize("sizes", [
    [(100, 80), (64, 48), (20, 20), (57, 33), (24, 24), (130, 25), (31, 200)],          # too-small images are skipped (round still advances)
    [(100, 80), (0, 0), (64, 48), (20, 20), (0, 0), (57, 33)],                          # unreadable files in the middle of bg.txt
    [(0, 0), (90, 70), (0, 0)],                                                         # first and last unreadable: `round` never advances
    [(70, 50), (0, 0)],
])
def test_reader_state_machine_matches_oracle(sizes):
    """NegReader::nextImg / get (src/imagestorage.cpp:57-126) without pixels vs the oracle's reader with pixels.  A (0, 0) entry is
    a file cv::imread could not decode: it keeps its slot in the list, nextImg does `last++; last %= count; continue` on it and
    leaves `round` alone (:62-70)."""
    from cascadeclassifier_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    imgs = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (w, h) in sizes]
    rd = orc.NegReader(24, 24, imgs)
    w = np.array([s[0] for s in sizes], np.int32)
    h = np.array([s[1] for s in sizes], np.int32)
    st = _lib.ReaderState()
    win = np.zeros(5, np.int32)
    for step in [1, 1, 2, 5, 40, 1, 300, 777, 4000, 9000]:
        last = None
        for _ in range(step):
            last = rd.get()
            assert last is not None
        _lib.check(lib.cc_reader_advance(w.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), len(sizes), 24, 24,
                                         C.byref(st), step, win.ctypes.data_as(C.c_void_p)))
        o = rd.state()
        mine = dict(last=st.last, round=st.round, point=(st.point_x, st.point_y), offset=(st.offset_x, st.offset_y),
                    img_size=(st.level_w, st.level_h), scale=st.scale)
        assert mine == o
        # the window the library says was consumed last is the one the oracle returned
        x, y, lw, lh, src = [int(v) for v in win]
        level = orc.resize_linear_exact(imgs[src], lw, lh)
        assert np.array_equal(level[y:y + 24, x:x + 24], last)


def test_reader_1080p_level_table():
    """1920x1080, 24x24, round 0 -> 11 levels, 13584 windows."""
    from cascadeclassifier_b200 import _lib
    lib = _lib.load()
    w = np.array([1920, 1920], np.int32)       # two images: the first one is scanned at round 0 (last/count == 0)
    h = np.array([1080, 1080], np.int32)
    st = _lib.ReaderState()
    _lib.check(lib.cc_reader_advance(w.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), 2, 24, 24, C.byref(st), 13584, None))
    # image 0 fully consumed: the reader has loaded image 1, which closes the list cycle (round 1, offset (1,0))
    assert (st.last, st.round, st.point_x, st.point_y, st.level_w, st.level_h) == (0, 1, 1, 0, 43, 24)
    st2 = _lib.ReaderState()
    _lib.check(lib.cc_reader_advance(w.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), 2, 24, 24, C.byref(st2), 13583, None))
    assert (st2.last, st2.round, st2.level_w, st2.level_h) == (1, 0, 1365, 768)


def test_tree_xml_encoding(tmp_path):
    """CvCascadeBoostTree::write (o_cvcascadeboosttree.cpp:41-93): BFS internal nodes, leaves 0,-1,-2.."""
    import cv2
    from cascadeclassifier_b200.boost import CvCascadeBoost
    from cascadeclassifier_b200.xml_io import XmlWriter
    leaf = lambda v: dict(left=-1, right=-1, var_idx=-1, value=v)
    depth2 = [dict(left=1, right=2, var_idx=7, c=0.5, value=0.0), dict(left=3, right=4, var_idx=9, c=-0.25, value=0.0), leaf(0.3),
              leaf(-0.7), leaf(0.9)]
    stump = [dict(left=1, right=2, var_idx=3, c=0.125, value=0.0), leaf(-1.0), leaf(1.0)]
    st = CvCascadeBoost.from_trees(-0.5, [stump, depth2]) if False else CvCascadeBoost()
    st.threshold, st.trees = -0.5, [stump, depth2]
    fmap = np.full(16, -1, np.int32)
    st.markUsedFeaturesInMap(fmap)
    assert sorted(np.nonzero(fmap >= 0)[0].tolist()) == [3, 7, 9]
    k = 0
    for i in range(16):
        if fmap[i] >= 0:
            fmap[i] = k
            k += 1
    path = str(tmp_path / "stage.xml")
    fs = XmlWriter(path)
    fs.start_map("stage")
    st.write(fs, fmap, 0)
    fs.end_map()
    fs.release()
    rd = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    node = rd.getNode("stage")
    assert int(node.getNode("maxWeakCount").real()) == 2
    wc = node.getNode("weakClassifiers")
    t0 = wc.at(0)
    inn = [t0.getNode("internalNodes").at(i).real() for i in range(4)]
    assert inn == [0.0, -1.0, 0.0, 0.125]                       # leaf 0, leaf -1, mapped feature 0, threshold
    lv = [t0.getNode("leafValues").at(i).real() for i in range(2)]
    assert lv == [-1.0, 1.0]
    t1 = wc.at(1)
    inn = [t1.getNode("internalNodes").at(i).real() for i in range(8)]
    assert inn == [1.0, 0.0, 1.0, 0.5, -1.0, -2.0, 2.0, -0.25]  # root: child node 1, leaf 0; node 1: leaves -1,-2
    lv = [t1.getNode("leafValues").at(i).real() for i in range(3)]
    assert lv == [np.float32(0.3), np.float32(-0.7), np.float32(0.9)]


def test_stage_xml_write_read_round_trip(tmp_path):
    """CvCascadeBoost::write -> stage<i>.xml -> CvCascadeBoost::read (the resume path of CvCascadeClassifier::load,
    src/cascadeclassifier.cpp:534-564): tree structure, variable indices, thresholds / category subsets survive; leaf
    values and the stage threshold come back rounded to f32, exactly what the reference's reader yields."""
    import cv2
    from cascadeclassifier_b200.boost import CvCascadeBoost
    from cascadeclassifier_b200.xml_io import XmlWriter
    rng = np.random.default_rng(3)

    def leaf(v):
        return dict(left=-1, right=-1, var_idx=-1, split_point=-1, c=0.0, quality=0.0, subset=[0] * 8, value=float(v), depth=0, sample_count=0)

    def inner(l, r, var, cat):
        nd = dict(left=l, right=r, var_idx=int(var), split_point=5, c=0.0, quality=1.0, subset=[0] * 8, value=0.0, depth=0, sample_count=0)
        if cat:
            nd["subset"] = [int(x) for x in rng.integers(-2**31, 2**31 - 1, 8)]
        else:
            nd["c"] = float(rng.normal())
        return nd

    for cat, max_cat in ((False, 0), (True, 256)):
        trees = []
        for t in range(5):
            if t % 2 == 0:       # stump
                trees.append([inner(1, 2, rng.integers(0, 5000), cat), leaf(rng.normal()), leaf(rng.normal())])
            else:                # depth 2, right child internal: root, leaf, inner, leaf, leaf
                trees.append([inner(1, 2, rng.integers(0, 5000), cat), leaf(rng.normal()), inner(3, 4, rng.integers(0, 5000), cat),
                              leaf(rng.normal()), leaf(rng.normal())])
        st = CvCascadeBoost.from_trees(-0.73251, trees)
        path = str(tmp_path / ("stage_%d.xml" % max_cat))
        fs = XmlWriter(path)
        fs.start_map("stage0")
        st.write(fs, None, max_cat)
        fs.end_map()
        fs.release()
        rd = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        back = CvCascadeBoost.read(rd.getFirstTopLevelNode(), max_cat)
        rd.release()
        assert back.threshold == float(np.float32(-0.73251))
        assert len(back.trees) == len(trees)

        def walk(tr, i):
            nd = tr[i]
            if nd["left"] < 0:
                return ("leaf", float(np.float32(nd["value"])))
            key = tuple(nd["subset"]) if cat else float(np.float32(nd["c"]))
            return (nd["var_idx"], key, walk(tr, nd["left"]), walk(tr, nd["right"]))

        for a, b in zip(trees, back.trees):
            assert walk(a, 0) == walk(b, 0)
        # and the C-side stage built from the read trees is what cc_cascade_set_stages consumes
        assert back.c_stage().num_weak == len(trees)


def test_train_error_paths_of_the_reference(tmp_path):
    """test_integration.cpp:221-281: a missing .vec makes train() return false (message as the reference prints it, no
    cascade.xml), an empty cascade directory name is rejected up front.  Both fail before any GPU work."""
    import io
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    bg = tmp_path / "bg.txt"
    bg.write_text("")
    data = tmp_path / "data"
    data.mkdir()
    out = io.StringIO()
    cc = CvCascadeClassifier(None, out=out)
    ok = cc.train(str(data), str(tmp_path / "no_such.vec"), str(bg), 10, 1, 64, 64, 1, CvCascadeParams(), ccb.CvHaarFeatureParams(),
                  ccb.CvCascadeBoostParams())
    assert ok is False
    assert "Image reader can not be created from -vec" in out.getvalue()
    assert not (data / "cascade.xml").exists()
    with pytest.raises(ccb.CascadeCudaError):
        CvCascadeClassifier(None, out=io.StringIO()).train("", "x.vec", str(bg), 10, 1, 64, 64, 1, CvCascadeParams(1 * 0, 1, (75, 32)),
                                                           ccb.CvFeatureParams.create(1), ccb.CvCascadeBoostParams())


def test_createsamples_from_info_collection(tmp_path):
    """`opencv_createsamples -info` (cvCreateTrainingSamplesFromInfo, tools/createsamples/utility.cpp:1125-1230): rectangles of the
    collection file, image paths relative to it, INTER_AREA when shrinking in both directions else INTER_LINEAR_EXACT, `num` in the
    header, a parse error ends the file."""
    import struct
    import cv2
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200.createsamples import cvCreateTrainingSamplesFromInfo
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    b = rng.integers(0, 256, (90, 70), dtype=np.uint8)
    d = tmp_path / "set"
    d.mkdir()
    cv2.imwrite(str(d / "a.png"), a)
    cv2.imwrite(str(d / "b.png"), b)
    (d / "info.dat").write_text("a.png 2 10 12 48 60 0 0 20 30\nb.png 1 5 6 24 24\na.png 1 1 2 30 oops\n")
    vec = tmp_path / "out" / "deep" / "s.vec"                     # icvMkDir: the directory hierarchy is created
    n = cvCreateTrainingSamplesFromInfo(str(d / "info.dat"), str(vec), 10, 0, 24, 24)
    assert n == 3                                                 # the 4th object has a parse error: the file ends there
    with open(vec, "rb") as f:
        count, vec_size, z0, z1 = struct.unpack("<iihh", f.read(12))
    assert (count, vec_size, z0, z1) == (10, 576, 0, 0)           # the header carries the requested number
    got = synth.read_vec(str(vec), 24, 24)
    assert got.shape == (3, 24, 24)
    assert np.array_equal(got[0], cv2.resize(a[12:72, 10:58], (24, 24), interpolation=cv2.INTER_AREA))
    assert np.array_equal(got[1], cv2.resize(a[0:30, 0:20], (24, 24), interpolation=cv2.INTER_LINEAR_EXACT))
    assert np.array_equal(got[2], b[6:30, 5:29])                  # same size: INTER_AREA of a 1:1 resize is the crop
    # `num` smaller than the collection: stops after num objects
    assert cvCreateTrainingSamplesFromInfo(str(d / "info.dat"), str(tmp_path / "two.vec"), 2, 0, 24, 24) == 2
    assert synth.read_vec(str(tmp_path / "two.vec"), 24, 24).shape == (2, 24, 24)
    # unreadable collection file: nothing written, 0 returned
    assert cvCreateTrainingSamplesFromInfo(str(d / "missing.dat"), str(tmp_path / "none.vec"), 2, 0, 24, 24) == 0
