"""End-to-end run of BASELINE configs[2]: traincascade Haar ALL 24x24, numStages=20, numPos=5000, numNeg=10000 on
synthetic positives / generated 1920x1080 backgrounds, through the host mirror of CvCascadeClassifier::train
(mining + stage training on the GPU, params.xml / stage<i>.xml / cascade.xml written by cv::FileStorage).
Usage: full_cascade.py [numStages] [numPos] [numNeg] [nBackgrounds] [--save-stages out.json] [--no-cache]
(under torchrun: feature- and image-sharded)"""
import io, os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams

save_stages = None
if "--save-stages" in sys.argv:
    i = sys.argv.index("--save-stages"); save_stages = sys.argv[i + 1]; del sys.argv[i:i + 2]
no_cache = "--no-cache" in sys.argv
if no_cache:
    sys.argv.remove("--no-cache")
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
if no_cache:
    ccb._lib.check(ctx.lib.cc_mine_set_cache(ctx.h, 0, 0), "cc_mine_set_cache")
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
    st6 = np.zeros(6, np.int64)
    ccb._lib.check(ctx.lib.cc_mine_get_stats(ctx.h, st6.ctypes.data_as(__import__("ctypes").c_void_p)), "cc_mine_get_stats")
    print("  mining: %d stream windows consumed, %d windows evaluated from pyramid planes, %d cached survivors re-evaluated, %d windows answered by the cache "
          "(%d entries, %.1f MB)" % (st6[0], st6[1], st6[2], st6[3], st6[4], st6[5] / 1e6))
    for nm in ("mine", "mine_host_assemble", "mine_host_eval", "mine_host_eval_levels", "mine_host_strips", "mine_host_order", "mine_host_alloc", "mine_host_eval_cand", "mine_host_wait", "mine_host_merge", "mine_host_cut"):
        ms, k = ctx.timing(nm)
        print("  timer %-20s %10.1f ms over %d" % (nm, ms, k))
    print("files:", sorted(os.listdir(d))[:6], "...")
    print(json.dumps(dict(stages=len(cc.stageClassifiers), total_s=t_train, mining_s=mine, train_s=train, weak=[t["weak"] for t in cc.timings],
                          neg_consumed=[r["neg_consumed"] for r in cc.reports], world=world, cache=not no_cache,
                          windows_consumed=int(st6[0]), windows_evaluated=int(st6[1]), cached_survivors_reevaluated=int(st6[2]),
                          windows_answered_by_cache=int(st6[3]), stream_windows_per_s=float(st6[0]) / max(mine, 1e-9),
                          evaluated_windows_per_s=float(st6[1]) / max(mine, 1e-9))))
    if save_stages:
        with open(save_stages, "w") as f:
            json.dump(dict(note="Haar ALL 24x24, numPos=%d numNeg=%d, %d stages, %d generated 1920x1080 backgrounds (tools/full_cascade.py)" % (npos, nneg, len(cc.stageClassifiers), nbg),
                           stages=[dict(threshold=s.threshold, trees=s.trees) for s in cc.stageClassifiers]), f)
    tail = out.getvalue().splitlines()[-12:]
    print("\n".join(tail))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
