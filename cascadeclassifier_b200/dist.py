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
