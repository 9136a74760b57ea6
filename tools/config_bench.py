"""BASELINE configs[0], [1], [3] at full size on synthetic data through the host mirror of CvCascadeClassifier::train
(a few stages each): timing + sanity (the parity of these paths is covered by tests/ at small sizes).
Usage: config_bench.py cfg1|cfg2|cfg4 [numStages]"""
import io, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

cfg = sys.argv[1]
spec = {"cfg1": dict(ft=0, fp=ccb.CvHaarFeatureParams(0), W=24, H=24, npos=500, nneg=1000, nst=2),
        "cfg2": dict(ft=1, fp=ccb.CvFeatureParams.create(1), W=24, H=24, npos=2000, nneg=4000, nst=10),
        "cfg4": dict(ft=2, fp=ccb.CvFeatureParams.create(2), W=48, H=96, npos=4000, nneg=8000, nst=15)}[cfg]
nst = int(sys.argv[2]) if len(sys.argv) > 2 else spec["nst"]
W, H = spec["W"], spec["H"]
t0 = time.time()
pos = synth.make_positives(int(spec["npos"] * 1.2) + 64, W, H, seed=12345)
bgs = [synth.make_background(960, 720, 4000 + i, clutter=1.5) for i in range(24)]

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
ctx = ccb.Context(local)
if world > 1:          # under torchrun: feature-sharded stage training + image-sharded mining over NCCL
    import torch
    import torch.distributed as dist
    from cascadeclassifier_b200.dist import init_comm
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    init_comm(ctx, dist, torch.device("cuda", local))
out = io.StringIO()
cc = CvCascadeClassifier(ctx, out=out)
t0 = time.time()
ok = cc.train(tempfile.mkdtemp(), "pos.vec", "bg.txt", spec["npos"], spec["nneg"], 1024, 1024, nst, CvCascadeParams(0, spec["ft"], (W, H)), spec["fp"],
              ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
ctx.synchronize()
if rank != 0:
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
print("%s: %d stages in %.2f s on %d GPU(s)" % (cfg, len(cc.stageClassifiers), time.time() - t0, world))
for t, r in zip(cc.timings, cc.reports):
    print("  stage %2d: weak %3d  mine %7.3f s  train %6.3f s  neg consumed %10d  acceptanceRatio %.3g" %
          (t["stage"], t["weak"], t["mine_s"], t["train_s"], r["neg_consumed"], r["acceptance_ratio"]))
for name in ["integral", "precalc", "sort", "split_search", "boost_small", "mine", "mine_host_assemble", "mine_host_eval", "mine_host_cut", "mine_host_alloc"]:
    ms, k = ctx.timing(name)
    print("   %-18s %9.3f ms  launches %d" % (name, ms, k))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
