"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(cascadeclassifier_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "liborc.so")

HAAR, LBP, HOG = 0, 1, 2
BASIC, CORE, ALL = 0, 1, 2
DAB, RAB, LB, GAB = 0, 1, 2, 3
SORT_STD, SORT_STABLE = 0, 1


class Node(C.Structure):
    _fields_ = [("left", C.c_int32), ("right", C.c_int32), ("var_idx", C.c_int32), ("split_point", C.c_int32),
                ("c", C.c_float), ("quality", C.c_float), ("subset", C.c_int32 * 8), ("value", C.c_double),
                ("depth", C.c_int32), ("sample_count", C.c_int32)]


class BoostParams(C.Structure):
    _fields_ = [("boost_type", C.c_int32), ("weak_count", C.c_int32), ("max_depth", C.c_int32),
                ("min_sample_count", C.c_int32), ("weight_trim_rate", C.c_double),
                ("regression_accuracy", C.c_float), ("min_hit_rate", C.c_float), ("max_false_alarm", C.c_float)]


class StageReport(C.Structure):
    _fields_ = [("pos_count", C.c_int32), ("neg_count", C.c_int32), ("trained", C.c_int32), ("pad", C.c_int32),
                ("pos_consumed", C.c_int64), ("neg_consumed", C.c_int64), ("acceptance_ratio", C.c_double)]


def boost_params(boost_type=GAB, weak_count=100, max_depth=1, min_sample_count=10, weight_trim_rate=0.95,
                 regression_accuracy=0.01, min_hit_rate=0.995, max_false_alarm=0.5):
    return BoostParams(boost_type, weak_count, max_depth, min_sample_count, weight_trim_rate,
                       regression_accuracy, min_hit_rate, max_false_alarm)


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src_mtime = max(os.path.getmtime(os.path.join(_ORACLE_DIR, f)) for f in os.listdir(_ORACLE_DIR)
                    if f.endswith((".cpp", ".hpp")))
    if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_mtime:
        build()
    L = C.CDLL(_LIB_PATH)
    vp, i32, f32, f64, i64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64
    P = C.POINTER
    sig = {
        "orc_integral": (None, [vp, i32, i32, vp, vp, vp]),
        "orc_norm_factor": (f32, [vp, i32, i32]),
        "orc_resize_linear_exact": (None, [vp, i32, i32, vp, i32, i32]),
        "orc_hog_planes": (None, [vp, i32, i32, vp]),
        "orc_fast_atan2": (f32, [f32, f32]),
        "orc_catalog_count": (i32, [i32, i32, i32, i32]),
        "orc_haar_catalog": (i32, [i32, i32, i32, vp, vp, vp]),
        "orc_lbp_catalog": (i32, [i32, i32, vp]),
        "orc_hog_catalog": (i32, [i32, i32, vp]),
        "orc_haar_calc": (f32, [i32, i32, vp, vp, vp]),
        "orc_eval_create": (vp, [i32, i32, i32, i32, i32]),
        "orc_eval_destroy": (None, [vp]),
        "orc_eval_num_features": (i32, [vp]),
        "orc_eval_var_count": (i32, [vp]),
        "orc_eval_feat_size": (i32, [vp]),
        "orc_eval_max_cat": (i32, [vp]),
        "orc_eval_set_images": (None, [vp, vp, vp, i32, i32, i32]),
        "orc_eval_value": (f32, [vp, i32, i32]),
        "orc_eval_values": (None, [vp, vp, i32, i32, i32, vp]),
        "orc_eval_get_integrals": (None, [vp, i32, vp, vp, vp]),
        "orc_eval_get_hog": (None, [vp, i32, vp]),
        "orc_td_create": (vp, [vp, i32, i32, i32, vp, i32]),
        "orc_td_destroy": (None, [vp]),
        "orc_td_rows": (i32, [vp]),
        "orc_td_fval": (vp, [vp]),
        "orc_td_sidx": (vp, [vp]),
        "orc_td_sidx32": (vp, [vp]),
        "orc_boost_train": (vp, [vp, P(BoostParams), i32, i32]),
        "orc_stage_destroy": (None, [vp]),
        "orc_stage_ok": (i32, [vp]),
        "orc_stage_seconds": (f64, [vp]),
        "orc_stage_num_weak": (i32, [vp]),
        "orc_stage_threshold": (f32, [vp]),
        "orc_stage_tree_num_nodes": (i32, [vp, i32]),
        "orc_stage_tree_nodes": (None, [vp, i32, P(Node)]),
        "orc_stage_log": (i32, [vp, vp, vp, vp, i32]),
        "orc_stage_snapshot": (i32, [vp, i32, vp, vp]),
        "orc_neg_create": (vp, [i32, i32]),
        "orc_neg_destroy": (None, [vp]),
        "orc_neg_add_image": (None, [vp, vp, i32, i32]),
        "orc_neg_get": (i32, [vp, vp]),
        "orc_neg_state": (None, [vp, vp, vp]),
        "orc_cascade_create": (vp, [i32, i32, i32, i32, i32, i32, i32, P(BoostParams), i32, i32, f64]),
        "orc_cascade_destroy": (None, [vp]),
        "orc_cascade_set_pos": (None, [vp, vp, i32]),
        "orc_cascade_add_bg": (None, [vp, vp, i32, i32]),
        "orc_cascade_train": (None, [vp]),
        "orc_cascade_termination": (i32, [vp]),
        "orc_cascade_num_stages": (i32, [vp]),
        "orc_cascade_num_reports": (i32, [vp]),
        "orc_cascade_report": (None, [vp, i32, P(StageReport)]),
        "orc_cascade_stage_threshold": (f32, [vp, i32]),
        "orc_cascade_stage_num_weak": (i32, [vp, i32]),
        "orc_cascade_tree_num_nodes": (i32, [vp, i32, i32]),
        "orc_cascade_tree_nodes": (None, [vp, i32, i32, P(Node)]),
        "orc_cascade_stage_log": (i32, [vp, i32, vp, vp, vp, i32]),
        "orc_cascade_push_stage": (None, [vp, f32, i32, vp, P(Node)]),
        "orc_cascade_update_training_set": (i32, [vp, f64, P(StageReport)]),
        "orc_cascade_fill": (i64, [vp, i32, i32, i32, f64, P(C.c_int)]),
        "orc_cascade_evaluator": (vp, [vp]),
        "orc_cascade_cur_num_samples": (i32, [vp]),
        "orc_cascade_neg_state": (None, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------------------------------------
def integral(img, tilted=False, sq=False):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    s = np.zeros((H + 1, W + 1), np.int32)
    q = np.zeros((H + 1, W + 1), np.float64) if sq else None
    t = np.zeros((H + 1, W + 1), np.int32) if tilted else None
    lib().orc_integral(_p(img), W, H, _p(s), _p(q), _p(t))
    return s, q, t


def norm_factor(img):
    img = np.ascontiguousarray(img, np.uint8)
    return float(lib().orc_norm_factor(_p(img), img.shape[1], img.shape[0]))


def resize_linear_exact(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_exact(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def hog_planes(img):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    out = np.zeros((10, H + 1, W + 1), np.float32)
    lib().orc_hog_planes(_p(img), W, H, _p(out))
    return out


def haar_catalog(W, H, mode):
    n = lib().orc_catalog_count(HAAR, mode, W, H)
    rects = np.zeros((n, 3, 4), np.int32)
    wts = np.zeros((n, 3), np.float32)
    tilted = np.zeros(n, np.int32)
    lib().orc_haar_catalog(W, H, mode, _p(rects), _p(wts), _p(tilted))
    return rects, wts, tilted


def lbp_catalog(W, H):
    n = lib().orc_catalog_count(LBP, 0, W, H)
    rects = np.zeros((n, 4), np.int32)
    lib().orc_lbp_catalog(W, H, _p(rects))
    return rects


def hog_catalog(W, H):
    n = lib().orc_catalog_count(HOG, 0, W, H)
    rects = np.zeros((n, 4), np.int32)
    lib().orc_hog_catalog(W, H, _p(rects))
    return rects


def tree_nodes_to_list(nodes):
    return [dict(left=n.left, right=n.right, var_idx=n.var_idx, split_point=n.split_point, c=n.c,
                 quality=n.quality, subset=list(n.subset), value=n.value, depth=n.depth,
                 sample_count=n.sample_count) for n in nodes]


class Evaluator:
    def __init__(self, ftype, mode, W, H, max_samples, handle=None):
        self.type, self.mode, self.W, self.H, self.max_samples = ftype, mode, W, H, max_samples
        self._own = handle is None
        self.h = lib().orc_eval_create(ftype, mode, W, H, max_samples) if handle is None else handle
        self.P = (W + 1) * (H + 1)

    def __del__(self):
        if getattr(self, "h", None) and self._own:
            lib().orc_eval_destroy(self.h)
            self.h = None

    @property
    def num_features(self):
        return lib().orc_eval_num_features(self.h)

    @property
    def var_count(self):
        return lib().orc_eval_var_count(self.h)

    def set_images(self, imgs, labels, first=0, threads=8):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        labels = np.ascontiguousarray(labels, np.uint8)
        assert imgs.shape[1:] == (self.H, self.W)
        lib().orc_eval_set_images(self.h, _p(imgs), _p(labels), first, imgs.shape[0], threads)

    def value(self, vi, si):
        return float(lib().orc_eval_value(self.h, vi, si))

    def values(self, vis, s0, ns):
        vis = np.ascontiguousarray(vis, np.int32)
        out = np.zeros((len(vis), ns), np.float32)
        lib().orc_eval_values(self.h, _p(vis), len(vis), s0, ns, _p(out))
        return out

    def integrals(self, si):
        s = np.zeros(self.P, np.int32)
        t = np.zeros(self.P, np.int32)
        nf = np.zeros(1, np.float32)
        lib().orc_eval_get_integrals(self.h, si, _p(s), _p(t), _p(nf))
        return s, t, float(nf[0])

    def hog(self, si):
        out = np.zeros((10, self.P), np.float32)
        lib().orc_eval_get_hog(self.h, si, _p(out))
        return out


class TrainData:
    def __init__(self, ev, N, sort_mode=SORT_STABLE, threads=8, subset=None):
        self.ev, self.N = ev, N
        sub = np.ascontiguousarray(subset, np.int32) if subset is not None else None
        self.h = lib().orc_td_create(ev.h, N, sort_mode, threads, _p(sub), 0 if sub is None else len(sub))
        self.F = lib().orc_td_rows(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_td_destroy(self.h)
            self.h = None

    def fval(self):
        ptr = lib().orc_td_fval(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(self.F, self.N))

    def sidx(self):
        ptr = lib().orc_td_sidx(self.h)
        if not ptr:
            ptr32 = lib().orc_td_sidx32(self.h)          # is_buf_16u == false (N >= 65536)
            if not ptr32:
                return None
            return np.ctypeslib.as_array(C.cast(ptr32, C.POINTER(C.c_uint32)), shape=(self.F, self.N))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint16)), shape=(self.F, self.N))


class StageResult:
    def __init__(self, h, N):
        self.h, self.N = h, N
        L = lib()
        self.ok = bool(L.orc_stage_ok(h))
        self.seconds = L.orc_stage_seconds(h)
        self.threshold = float(L.orc_stage_threshold(h))
        self.trees = []
        for t in range(L.orc_stage_num_weak(h)):
            n = L.orc_stage_tree_num_nodes(h, t)
            arr = (Node * n)()
            L.orc_stage_tree_nodes(h, t, arr)
            self.trees.append(tree_nodes_to_list(arr))
        cap = 1024
        n_ = np.zeros(cap, np.int32)
        hr = np.zeros(cap, np.float32)
        fa = np.zeros(cap, np.float32)
        k = L.orc_stage_log(h, _p(n_), _p(hr), _p(fa), cap)
        self.log = [(int(n_[i]), float(hr[i]), float(fa[i])) for i in range(k)]

    def snapshot(self, t):
        w = np.zeros(self.N, np.float64)
        m = np.zeros(self.N, np.uint8)
        ok = lib().orc_stage_snapshot(self.h, t, _p(w), _p(m))
        return (w, m) if ok else None

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_stage_destroy(self.h)
            self.h = None


def boost_train(td, params, cat_sort=SORT_STABLE, record=False):
    h = lib().orc_boost_train(td.h, C.byref(params), cat_sort, int(record))
    return StageResult(h, td.N)


class NegReader:
    def __init__(self, W, H, images):
        self.W, self.H = W, H
        self.h = lib().orc_neg_create(W, H)
        for im in images:
            im = np.ascontiguousarray(im, np.uint8)
            lib().orc_neg_add_image(self.h, _p(im), im.shape[1], im.shape[0])

    def get(self):
        out = np.zeros((self.H, self.W), np.uint8)
        ok = lib().orc_neg_get(self.h, _p(out))
        return out if ok else None

    def state(self):
        st = np.zeros(8, np.int64)
        sc = np.zeros(1, np.float32)
        lib().orc_neg_state(self.h, _p(st), _p(sc))
        return dict(last=int(st[0]), round=int(st[1]), point=(int(st[2]), int(st[3])), offset=(int(st[4]), int(st[5])),
                    img_size=(int(st[6]), int(st[7])), scale=float(sc[0]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_neg_destroy(self.h)
            self.h = None


class Cascade:
    def __init__(self, ftype, mode, W, H, num_pos, num_neg, num_stages, params, sort_mode=SORT_STABLE, threads=8,
                 break_value=-1.0):
        self.W, self.H = W, H
        self.ftype, self.mode = ftype, mode
        self.max_samples = num_pos + num_neg
        self.h = lib().orc_cascade_create(ftype, mode, W, H, num_pos, num_neg, num_stages, C.byref(params),
                                          sort_mode, threads, break_value)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cascade_destroy(self.h)
            self.h = None

    def set_pos(self, samples):
        samples = np.ascontiguousarray(samples, np.uint8)
        lib().orc_cascade_set_pos(self.h, _p(samples), samples.shape[0])

    def add_bg(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        lib().orc_cascade_add_bg(self.h, _p(img), img.shape[1], img.shape[0])

    def train(self):
        lib().orc_cascade_train(self.h)

    @property
    def termination(self):
        return lib().orc_cascade_termination(self.h)

    def reports(self):
        out = []
        for i in range(lib().orc_cascade_num_reports(self.h)):
            r = StageReport()
            lib().orc_cascade_report(self.h, i, C.byref(r))
            out.append(dict(pos_count=r.pos_count, neg_count=r.neg_count, trained=r.trained,
                            pos_consumed=r.pos_consumed, neg_consumed=r.neg_consumed,
                            acceptance_ratio=r.acceptance_ratio))
        return out

    def stages(self):
        L = lib()
        out = []
        for s in range(L.orc_cascade_num_stages(self.h)):
            trees = []
            for t in range(L.orc_cascade_stage_num_weak(self.h, s)):
                n = L.orc_cascade_tree_num_nodes(self.h, s, t)
                arr = (Node * n)()
                L.orc_cascade_tree_nodes(self.h, s, t, arr)
                trees.append(tree_nodes_to_list(arr))
            out.append(dict(threshold=float(L.orc_cascade_stage_threshold(self.h, s)), trees=trees))
        return out

    def stage_log(self, s):
        cap = 1024
        n_ = np.zeros(cap, np.int32)
        hr = np.zeros(cap, np.float32)
        fa = np.zeros(cap, np.float32)
        k = lib().orc_cascade_stage_log(self.h, s, _p(n_), _p(hr), _p(fa), cap)
        return [(int(n_[i]), float(hr[i]), float(fa[i])) for i in range(k)]

    def push_stage(self, threshold, trees):
        counts = np.array([len(t) for t in trees], np.int32)
        total = int(counts.sum())
        arr = (Node * total)()
        k = 0
        for t in trees:
            for nd in t:
                a = arr[k]
                a.left, a.right, a.var_idx = nd["left"], nd["right"], nd["var_idx"]
                a.split_point = nd.get("split_point", -1)
                a.c, a.quality = nd.get("c", 0.0), nd.get("quality", 0.0)
                for j in range(8):
                    a.subset[j] = nd.get("subset", [0] * 8)[j]
                a.value = nd["value"]
                a.depth, a.sample_count = nd.get("depth", 0), nd.get("sample_count", 0)
                k += 1
        lib().orc_cascade_push_stage(self.h, threshold, len(trees), _p(counts), arr)

    def update_training_set(self, min_ratio):
        r = StageReport()
        ok = lib().orc_cascade_update_training_set(self.h, min_ratio, C.byref(r))
        return bool(ok), dict(pos_count=r.pos_count, neg_count=r.neg_count, pos_consumed=r.pos_consumed,
                              neg_consumed=r.neg_consumed, acceptance_ratio=r.acceptance_ratio)

    def fill(self, first, count, is_positive, min_ratio):
        got = C.c_int(0)
        consumed = lib().orc_cascade_fill(self.h, first, count, int(is_positive), min_ratio, C.byref(got))
        return got.value, int(consumed)

    def evaluator(self):
        return Evaluator(self.ftype, self.mode, self.W, self.H, self.max_samples,
                         handle=lib().orc_cascade_evaluator(self.h))

    def neg_state(self):
        st = np.zeros(8, np.int64)
        sc = np.zeros(1, np.float32)
        lib().orc_cascade_neg_state(self.h, _p(st), _p(sc))
        return dict(last=int(st[0]), round=int(st[1]), point=(int(st[2]), int(st[3])), offset=(int(st[4]), int(st[5])),
                    img_size=(int(st[6]), int(st[7])), scale=float(sc[0]))
