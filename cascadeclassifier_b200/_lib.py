"""ctypes loader for libcascade_cuda.so (the C ABI declared in include/cascade_cuda.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
context is created, an exception is raised.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libcascade_cuda.so")

CC_MAX_TREE_NODES = 63


class CascadeCudaError(RuntimeError):
    pass


class BoostParams(C.Structure):
    _fields_ = [("boost_type", C.c_int32), ("weak_count", C.c_int32), ("max_depth", C.c_int32),
                ("min_sample_count", C.c_int32), ("weight_trim_rate", C.c_double),
                ("regression_accuracy", C.c_float), ("min_hit_rate", C.c_float), ("max_false_alarm", C.c_float)]


class Node(C.Structure):
    _fields_ = [("left", C.c_int32), ("right", C.c_int32), ("var_idx", C.c_int32), ("split_point", C.c_int32),
                ("c", C.c_float), ("quality", C.c_float), ("subset", C.c_int32 * 8), ("value", C.c_double),
                ("depth", C.c_int32), ("sample_count", C.c_int32)]


class Tree(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("pad", C.c_int32), ("nodes", Node * CC_MAX_TREE_NODES)]


class LogRow(C.Structure):
    _fields_ = [("n", C.c_int32), ("hit_rate", C.c_float), ("false_alarm", C.c_float)]


class Stage(C.Structure):
    _fields_ = [("threshold", C.c_float), ("num_weak", C.c_int32), ("num_log", C.c_int32), ("trained", C.c_int32),
                ("trees", C.POINTER(Tree)), ("log", C.POINTER(LogRow))]


class ReaderState(C.Structure):
    _fields_ = [("last", C.c_int64), ("round", C.c_int64), ("point_x", C.c_int32), ("point_y", C.c_int32),
                ("offset_x", C.c_int32), ("offset_y", C.c_int32), ("level_w", C.c_int32), ("level_h", C.c_int32),
                ("scale", C.c_float), ("pad", C.c_int32)]


_SIGNATURES = None
_lib = None


def _signatures():
    vp, i32, f64, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    P = C.POINTER
    return {
        "cc_create": (i32, [i32, P(vp)]),
        "cc_destroy": (None, [vp]),
        "cc_last_error": (C.c_char_p, []),
        "cc_version": (C.c_char_p, []),
        "cc_set_stream": (i32, [vp, vp]),
        "cc_synchronize": (i32, [vp]),
        "cc_eval_init": (i32, [vp, i32, i32, i32, i32, i32]),
        "cc_eval_num_features": (i32, [vp]),
        "cc_eval_feature_size": (i32, [vp]),
        "cc_eval_max_cat_count": (i32, [vp]),
        "cc_eval_get_feature": (i32, [vp, i32, vp, vp, vp]),
        "cc_eval_set_images": (i32, [vp, vp, vp, i32, i32]),
        "cc_eval_set_images_device": (i32, [vp, vp, vp, i32, i32]),
        "cc_eval_value": (i32, [vp, i32, i32, vp]),
        "cc_eval_values": (i32, [vp, vp, i32, i32, i32, vp]),
        "cc_eval_get_integrals": (i32, [vp, i32, vp, vp, vp]),
        "cc_eval_get_hog_planes": (i32, [vp, i32, vp]),
        "cc_eval_get_hog_planes_compact": (i32, [vp, i32, vp]),
        "cc_eval_get_labels": (i32, [vp, i32, i32, vp]),
        "cc_eval_get_images": (i32, [vp, i32, i32, vp]),
        "cc_boost_set_shard": (i32, [vp, i32, i32]),
        "cc_comm_unique_id": (i32, [vp]),
        "cc_comm_init": (i32, [vp, vp, i32, i32]),
        "cc_boost_precalc": (i32, [vp, i32]),
        "cc_boost_get_fval": (i32, [vp, i32, i32, vp]),
        "cc_boost_get_sorted_idx": (i32, [vp, i32, i32, vp]),
        "cc_boost_get_sorted_val": (i32, [vp, i32, i32, vp]),
        "cc_boost_get_sorted_idx32": (i32, [vp, i32, i32, vp]),
        "cc_boost_train": (i32, [vp, i32, P(BoostParams), P(Stage)]),
        "cc_boost_get_weights": (i32, [vp, vp, vp]),
        "cc_boost_get_row_quality": (i32, [vp, i32, i32, vp]),
        "cc_boost_set_search_mode": (i32, [vp, i32]),
        "cc_boost_set_sort_mode": (i32, [vp, i32]),
        "cc_boost_set_tree_path": (i32, [vp, i32]),
        "cc_boost_get_sort_stats": (i32, [vp, P(i64)]),
        "cc_boost_get_search_stats": (i32, [vp, P(i64), P(i64), P(i64)]),
        "cc_stage_predict": (i32, [vp, P(Stage), i32, i32, i32, vp]),
        "cc_bg_clear": (i32, [vp]),
        "cc_bg_register": (i32, [vp, vp, i32, i32]),
        "cc_bg_register_device": (i32, [vp, vp, i32, i32, i32]),
        "cc_bg_count": (i32, [vp]),
        "cc_cascade_set_stages": (i32, [vp, P(Stage), i32]),
        "cc_mine_negatives": (i32, [vp, i32, i32, f64, P(ReaderState), P(i64), P(C.c_int32)]),
        "cc_mine_positives": (i32, [vp, vp, i32, i32, P(C.c_int32), P(i64), P(C.c_int32)]),
        "cc_mine_scan": (i32, [vp, i32, i32, P(i64), P(i64)]),
        "cc_mine_scan_digest": (i32, [vp, i32, i32, P(i64), P(i64), P(C.c_uint64)]),
        "cc_mine_set_mode": (i32, [vp, i32]),
        "cc_mine_set_cache": (i32, [vp, i32, i64]),
        "cc_mine_get_stats": (i32, [vp, vp]),
        "cc_mine_serial_cut": (i32, [vp, i64, i64, i32, f64, P(i64), P(C.c_int32), P(i64), P(C.c_int32), vp, P(C.c_int32)]),
        "cc_reader_advance": (i32, [vp, vp, i32, i32, i32, P(ReaderState), i64, vp]),
        "cc_createsamples": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, f64, f64, f64, i32, i32, i32, vp]),
        "cc_createsamples_plan": (i32, [i32, i32, i32, i32, i32, f64, f64, f64, i32, i32, i32, vp, vp, vp]),
        "cc_timing_reset": (i32, [vp]),
        "cc_timing_get": (i32, [vp, C.c_char_p, P(f64), P(i64)]),
        "cc_launch_count": (i64, [vp]),
        "cc_debug_div_check": (i32, [vp, i32, C.c_uint64, i64, P(C.c_uint64)]),
    }


def exported_symbols():
    """Names every build of the library must export (mirrors include/cascade_cuda.h)."""
    return sorted(_signatures().keys())


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CascadeCudaError(
            "libcascade_cuda.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C cascadeclassifier_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _signatures().items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().cc_last_error()
        raise CascadeCudaError("%s failed (status %d): %s" % (what or "libcascade_cuda call", status,
                                                              msg.decode() if msg else "?"))
