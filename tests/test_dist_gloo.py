"""CPU, world_size 2 over gloo: the feature-sharded best-split merge (cascadeclassifier_b200.dist) gives the same
winner as an unsharded search.  The per-shard searches are done by the oracle (the checker), because no GPU is
present here; on the GPU box the same rule runs in k_argmax_candidates after the NCCL all-gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import orc
    from conftest import make_samples
    from cascadeclassifier_b200.dist import select_winner, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs, labels = make_samples(60, 120)
    e = orc.Evaluator(orc.HAAR, orc.BASIC, 24, 24, len(imgs))
    e.set_images(imgs, labels)
    F = 4096
    all_rows = np.arange(0, F * 8, 8, dtype=np.int32)          # a fixed subset of variables, sharded contiguously
    lo, hi = shard_range(F, rank, world)
    td = orc.TrainData(e, len(imgs), orc.SORT_STABLE, 2, subset=all_rows[lo:hi])
    res = orc.boost_train(td, orc.boost_params(weak_count=1, max_false_alarm=0.01))
    root = res.trees[0][0]
    cand = torch.tensor([1.0, root["quality"], float(all_rows[lo + root["var_idx"]]), root["c"]], dtype=torch.float64)
    gathered = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, cand)
    cands = [dict(valid=bool(g[0]), quality=float(np.float32(g[1])), var_idx=int(g[2]), c=float(g[3])) for g in gathered]
    win = select_winner(cands)
    if rank == 0:
        q.put((win, cands))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_argmax_matches_unsharded():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from conftest import make_samples
    from cascadeclassifier_b200.dist import shard_range
    assert shard_range(261600, 0, 8) == (0, 32700) and shard_range(261600, 7, 8) == (228900, 261600)
    assert [shard_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    win, cands = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    imgs, labels = make_samples(60, 120)
    e = orc.Evaluator(orc.HAAR, orc.BASIC, 24, 24, len(imgs))
    e.set_images(imgs, labels)
    all_rows = np.arange(0, 4096 * 8, 8, dtype=np.int32)
    td = orc.TrainData(e, len(imgs), orc.SORT_STABLE, 4, subset=all_rows)
    res = orc.boost_train(td, orc.boost_params(weak_count=1, max_false_alarm=0.01))
    root = res.trees[0][0]
    assert win["var_idx"] == int(all_rows[root["var_idx"]])
    assert win["quality"] == pytest.approx(root["quality"], rel=1e-6)
    assert win["c"] == root["c"]
    assert len(cands) == 2 and all(c["valid"] for c in cands)


def test_select_winner_tie_rule():
    from cascadeclassifier_b200.dist import select_winner
    c = [dict(valid=True, quality=0.5, var_idx=9), dict(valid=True, quality=0.5, var_idx=4), dict(valid=False, quality=9.0, var_idx=1),
         dict(valid=True, quality=0.25, var_idx=0)]
    assert select_winner(c)["var_idx"] == 4
    assert select_winner([dict(valid=False, quality=1.0, var_idx=0)]) is None


def _naive_fill(pass_mask, count, min_ratio, consumed=0, got=0):
    """fillPassedSamples window by window (src/cascadeclassifier.cpp:333-355) over a batch with known outcomes."""
    accepted, i, n = [], 0, len(pass_mask)
    stopped = False
    while got < count:
        # inner for(;;): one window per iteration
        if consumed != 0 and (got + 1.0) / consumed <= min_ratio:
            stopped = True
            break
        if i >= n:
            break
        consumed += 1
        ok = pass_mask[i]
        i += 1
        if ok:
            accepted.append(i - 1)
            got += 1
    if got >= count:
        stopped = True
    return dict(consumed=consumed, got=got, next_window=i, stopped=stopped, accepted=accepted)


def test_serial_cut_matches_window_by_window_loop():
    from cascadeclassifier_b200.dist import serial_cut
    rng = np.random.default_rng(11)
    for trial in range(300):
        nwin = int(rng.integers(1, 4000))
        rate = float(rng.choice([0.5, 0.1, 0.01, 0.001, 0.9]))
        mask = rng.random(nwin) < rate
        count = int(rng.integers(1, 60))
        min_ratio = float(rng.choice([0.0, 1e-6, 1e-3, 0.02, 0.2, 0.7]))
        c0 = int(rng.integers(0, 50)) if trial % 3 else 0
        g0 = int(rng.integers(0, min(count, c0 + 1))) if c0 else 0
        ref = _naive_fill(mask, count, min_ratio, c0, g0)
        got = serial_cut(np.nonzero(mask)[0], nwin, count, min_ratio, c0, g0)
        assert got["consumed"] == ref["consumed"] and got["got"] == ref["got"], (trial, got, ref)
        assert got["next_window"] == ref["next_window"] and list(got["accepted"]) == ref["accepted"]
        assert got["stopped"] == ref["stopped"]


def _mine_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from cascadeclassifier_b200.dist import image_owner, survivors_to_bitmap, bitmap_to_survivors, serial_cut
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                       # same stream on every rank: the batch description is replicated
    per_image = [int(x) for x in rng.integers(50, 400, 9)]           # windows per background of the batch
    nwin = sum(per_image)
    outcome = rng.random(nwin) < 0.03                     # what the cascade would decide for every window
    base = np.concatenate([[0], np.cumsum(per_image)])
    mine = np.zeros(nwin, bool)
    for i in range(len(per_image)):
        if image_owner(i, world) == rank:
            mine[base[i]:base[i + 1]] = True
    local = np.nonzero(outcome & mine)[0]                 # this rank only evaluates the images it owns
    bm = torch.from_numpy(survivors_to_bitmap(local, nwin).astype(np.int64))
    dist.all_reduce(bm, op=dist.ReduceOp.SUM)             # NCCL: ncclAllReduce(sum, uint32) on the device bitmap
    merged = bitmap_to_survivors(bm.numpy().astype(np.uint32))
    res = serial_cut(merged, nwin, 12, 1e-4)
    q.put((rank, merged.tolist(), res["consumed"], res["got"], res["accepted"].tolist(), np.nonzero(outcome)[0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_image_sharded_mining_merge_gloo():
    """world_size 2 over gloo: per-rank survivors of disjoint image sets -> summed bitmaps -> the same ordered survivor
    list and the same serial cut on every rank as an unsharded evaluation."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort()
    assert outs[0][1:] == outs[1][1:]                     # identical on both ranks
    _, merged, consumed, got, accepted, truth = outs[0]
    assert merged == truth
    mask = np.zeros(truth[-1] + 1, bool)
    mask[truth] = True
    ref = _naive_fill(mask, 12, 1e-4)
    assert got == 12 and accepted == truth[:12] and consumed == truth[11] + 1 == ref["consumed"]
