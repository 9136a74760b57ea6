"""Multi-GPU plumbing for feature-sharded stage training: one process per GPU.

The data path stays in libcascade_cuda.so: every rank precalcs / sorts / searches its own variable slice, and
per tree node one NCCL all-gather of a 64-byte candidate per rank + a device-side argmax picks the winner
(highest f32 quality, lowest variable index on ties — CvDTree::find_best_split, o_cvdtree.cpp:313-357).
This module only (1) computes the slice, (2) ships the ncclUniqueId through torch.distributed and
(3) states the merge rule in Python so that it can be tested with gloo on CPU.
"""
import ctypes as C

from . import _lib


def shard_range(num_vars, rank, world):
    """Contiguous slice [lo, hi) of the boosting variables owned by `rank` (matches cc_boost_set_shard)."""
    lo = num_vars * rank // world
    hi = num_vars * (rank + 1) // world
    return lo, hi


def select_winner(candidates):
    """candidates: iterable of dicts {valid, quality (f32), var_idx, ...}; returns the global best or None."""
    best = None
    for c in candidates:
        if not c.get("valid", False):
            continue
        if best is None or c["quality"] > best["quality"] or (c["quality"] == best["quality"] and c["var_idx"] < best["var_idx"]):
            best = c
    return best


def init_comm(ctx, dist, device=None):
    """Create the library's NCCL communicator on every rank (unique id broadcast via torch.distributed)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    raw = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.check(ctx.lib.cc_comm_unique_id(raw), "cc_comm_unique_id")
    t = torch.tensor(list(raw), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, 0)
    raw = (C.c_ubyte * 128)(*t.cpu().tolist())
    _lib.check(ctx.lib.cc_comm_init(ctx.h, raw, rank, world), "cc_comm_init")
    return rank, world


# ---- image-sharded mining --------------------------------------------------------------------------
def image_owner(image_ordinal, world):
    """Rank that builds and evaluates the pyramid of the image_ordinal-th background of a mining batch (round-robin,
    matches eval_levels in csrc/cc_mine.cu)."""
    return image_ordinal % world


def survivors_to_bitmap(ordinals, nwin):
    """uint32 bitmap over the batch ordinals; the ranks' bitmaps are disjoint, so their sum is their union."""
    import numpy as np
    bm = np.zeros((nwin + 31) // 32, np.uint32)
    o = np.asarray(ordinals, np.int64)
    np.bitwise_or.at(bm, o >> 5, (np.uint32(1) << (o & 31).astype(np.uint32)))
    return bm


def bitmap_to_survivors(bitmap):
    """Ascending batch ordinals of the set bits (the ordered survivor list every rank feeds to the serial cut)."""
    import numpy as np
    bits = np.unpackbits(np.asarray(bitmap, np.uint32).view(np.uint8), bitorder="little")
    return np.nonzero(bits)[0].astype(np.uint32)


def serial_cut(survivors, nwin, count, min_ratio, consumed=0, got=0):
    """cc_mine_serial_cut: the reference's serial stopping rule over an evaluated batch (host only)."""
    import numpy as np
    lib = _lib.load()
    surv = np.ascontiguousarray(survivors, np.uint32)
    c, g, nxt, st, na = C.c_int64(consumed), C.c_int32(got), C.c_int64(0), C.c_int32(0), C.c_int32(0)
    acc = np.zeros(max(1, count), np.uint32)
    _lib.check(lib.cc_mine_serial_cut(surv.ctypes.data_as(C.c_void_p), len(surv), int(nwin), int(count), float(min_ratio), C.byref(c),
                                      C.byref(g), C.byref(nxt), C.byref(st), acc.ctypes.data_as(C.c_void_p), C.byref(na)), "cc_mine_serial_cut")
    return dict(consumed=c.value, got=g.value, next_window=nxt.value, stopped=bool(st.value), accepted=acc[:na.value].copy())
