"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank trains the same small Haar-ALL cascade twice: (a) with the NCCL communicator attached -> feature-sharded
stage training (all-gather of 64-byte candidates per node) + image-sharded mining (survivor bitmap / accepted pixels
summed over NVLink), (b) on a private single-GPU context.  Trees, thresholds, consumed counts, reader state and the
mined windows must be identical."""
import io, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import synth
from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
from cascadeclassifier_b200.dist import init_comm

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
FT = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # 0 Haar ALL (default), 1 LBP, 2 HOG 48x96
W, H = (48, 96) if FT == 2 else (24, 24)
pos = synth.make_positives(700 if FT != 2 else 400, W, H, seed=7)
bgs = synth.make_backgrounds(9, 400, 300, seed0=70, clutter=2.0)


def run(ctx):
    cc = CvCascadeClassifier(ctx, out=io.StringIO())
    d = tempfile.mkdtemp()
    fp = ccb.CvHaarFeatureParams(2) if FT == 0 else ccb.CvFeatureParams.create(FT)
    npos, nneg, nst = (500, 900, 5) if FT != 2 else (300, 600, 8)
    cc.train(d, "pos.vec", "bg.txt", npos, nneg, 1024, 1024, nst, CvCascadeParams(0, FT, (W, H)), fp,
             ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
    imgs = cc.featureEvaluator.images(0, cc.curNumSamples)
    st = cc._reader
    return cc, imgs, (st.last, st.round, st.point_x, st.point_y, st.offset_x, st.offset_y, st.level_w, st.level_h, st.scale)


ctx_d = ccb.Context(local)
init_comm(ctx_d, dist, dev)
cc_d, imgs_d, st_d = run(ctx_d)
ctx_s = ccb.Context(local)
cc_s, imgs_s, st_s = run(ctx_s)
ok = len(cc_d.stageClassifiers) == len(cc_s.stageClassifiers) and len(cc_s.stageClassifiers) >= 1
for a, b in zip(cc_d.stageClassifiers, cc_s.stageClassifiers):
    ok = ok and a.threshold == b.threshold and a.trees == b.trees
ok = ok and cc_d.reports == cc_s.reports and st_d == st_s and np.array_equal(imgs_d, imgs_s)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check ftype=%d world=%d: %s  stages=%d weak=%s neg_consumed=%s" % (FT, world, "OK" if flag.item() == 1 else "MISMATCH", len(cc_s.stageClassifiers),
          [len(s.trees) for s in cc_s.stageClassifiers], [r["neg_consumed"] for r in cc_s.reports]))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
