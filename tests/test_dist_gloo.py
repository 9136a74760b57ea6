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
