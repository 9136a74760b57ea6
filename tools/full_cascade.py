"""End-to-end run of BASELINE configs[2]: traincascade Haar ALL 24x24, numStages=20, numPos=5000, numNeg=10000 on
synthetic positives / generated 1920x1080 backgrounds, through the host mirror of CvCascadeClassifier::train
(mining + stage training on the GPU, params.xml / stage<i>.xml / cascade.xml written by cv::FileStorage).
Usage: full_cascade.py [numStages] [numPos] [numNeg] [nBackgrounds]   (under torchrun: feature- and image-sharded)"""
import io, os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

nst = int(sys.argv[1]) if len(sys.argv) > 1 else 20
npos = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
nneg = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
nbg = int(sys.argv[4]) if len(sys.argv) > 4 else 48
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
t0 = time.time()
pos = synth.make_positives(int(npos * 1.15) + 64, 24, 24, seed=12345)
bgs = [synth.make_background(1920, 1080, 4000 + i, clutter=1.5) for i in range(nbg)]
t_data = time.time() - t0
ctx = ccb.Context(local)
if world > 1:
    import torch, torch.distributed as dist
    from cascadeclassifier_b200.dist import init_comm
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    init_comm(ctx, dist, torch.device("cuda", local))
out = io.StringIO()
cc = CvCascadeClassifier(ctx, out=out)
d = tempfile.mkdtemp(prefix="cascade_r%d_" % rank)
t0 = time.time()
ok = cc.train(d, "pos.vec", "bg.txt", npos, nneg, 1024, 1024, nst, CvCascadeParams(0, 0, (24, 24)), ccb.CvHaarFeatureParams(2),
              ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
ctx.synchronize()
t_train = time.time() - t0
if rank == 0:
    mine = sum(t["mine_s"] for t in cc.timings); train = sum(t["train_s"] for t in cc.timings)
    print("data generation %.1f s; cascade: %d stages in %.2f s (mining %.2f s, stage training %.2f s), world=%d" %
          (t_data, len(cc.stageClassifiers), t_train, mine, train, world))
    for t, r in zip(cc.timings, cc.reports):
        print("  stage %2d: weak %3d  mine %7.3f s  train %6.3f s  neg consumed %12d  acceptanceRatio %.3g" %
              (t["stage"], t["weak"], t["mine_s"], t["train_s"], r["neg_consumed"], r["acceptance_ratio"]))
    for nm in ("mine", "mine_host_assemble", "mine_host_eval", "mine_host_cut"):
        ms, k = ctx.timing(nm)
        print("  timer %-20s %10.1f ms over %d" % (nm, ms, k))
    print("files:", sorted(os.listdir(d))[:6], "...")
    print(json.dumps(dict(stages=len(cc.stageClassifiers), total_s=t_train, mining_s=mine, train_s=train, weak=[t["weak"] for t in cc.timings],
                          neg_consumed=[r["neg_consumed"] for r in cc.reports], world=world)))
    tail = out.getvalue().splitlines()[-12:]
    print("\n".join(tail))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
