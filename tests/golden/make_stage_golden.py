"""Generates tests/golden/baseline_size_oracle.json: results of the CPU ORACLE (oracle/, the restatement of the reference) at the
BASELINE.json sizes, so that the GPU parity tests at those sizes (tests/test_gpu_fullsize.py) do not have to run the oracle for
minutes on the GPU box.  Run in the build container (8 cores: ~8 minutes):

    python tests/golden/make_stage_golden.py

  cfg1          BASELINE configs[0]: traincascade Haar BASIC 24x24, numStages 2, numPos 500, numNeg 1000 -- whole run: mining + both
                stages (N = 1500, F = 162336)
  cfg2          BASELINE configs[1]: traincascade LBP 24x24, numStages 10, numPos 2000, numNeg 4000 -- whole run, until the required
                leaf false-alarm rate stops it (N = 6000, F = 8464)
  cfg3_stage0   BASELINE configs[2]: Haar ALL 24x24, numPos 5000, numNeg 10000 -- the first stage on bench.py's sample set
                (N = 15000, F = 261600: every feature, std::sort tie order like the reference)

Inputs are the seeded generators of cascadeclassifier_b200/synth.py (numpy only), so the GPU box regenerates them bit-identically."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import orc  # noqa: E402
from cascadeclassifier_b200 import synth  # noqa: E402

SPECS = {
    "cfg1": dict(ftype=orc.HAAR, mode=orc.BASIC, W=24, H=24, npos=500, nneg=1000, nstages=2, nvec=640, nbg=8, bw=640, bh=480, seed=2101),
    "cfg2": dict(ftype=orc.LBP, mode=0, W=24, H=24, npos=2000, nneg=4000, nstages=10, nvec=2300, nbg=8, bw=640, bh=480, seed=2202),
}


def cascade_inputs(spec):
    pos = synth.make_positives(spec["nvec"], spec["W"], spec["H"], seed=spec["seed"])
    bgs = synth.make_backgrounds(spec["nbg"], spec["bw"], spec["bh"], seed0=spec["seed"] * 10, clutter=2.0)
    return pos, bgs


def run_cascade(spec, sort_mode):
    pos, bgs = cascade_inputs(spec)
    o = orc.Cascade(spec["ftype"], spec["mode"], spec["W"], spec["H"], spec["npos"], spec["nneg"], spec["nstages"], orc.boost_params(),
                    sort_mode=sort_mode, threads=os.cpu_count() or 8)
    o.set_pos(pos)
    for b in bgs:
        o.add_bg(b)
    t0 = time.time()
    o.train()
    return dict(stages=o.stages(), reports=o.reports(), neg_state=o.neg_state(), logs=[o.stage_log(s) for s in range(len(o.stages()))],
                seconds=time.time() - t0)


if __name__ == "__main__":
    path = os.path.join(HERE, "baseline_size_oracle.json")
    only = sys.argv[1:]                  # e.g. `make_stage_golden.py cfg2` regenerates one entry of an existing file
    out = json.load(open(path)) if (only and os.path.exists(path)) else {}
    for name, spec in SPECS.items():
        if only and name not in only:
            continue
        r = run_cascade(spec, orc.SORT_STABLE)
        r2 = run_cascade(spec, orc.SORT_STD)       # the reference's std::sort tie order must not change anything
        assert r["stages"] == r2["stages"] and r["reports"] == r2["reports"], name
        out[name] = dict(spec={k: v for k, v in spec.items()}, **r)
        print(name, "%.1fs" % r["seconds"], [len(s["trees"]) for s in r["stages"]], [x["neg_consumed"] for x in r["reports"]], flush=True)
    if only and "cfg3_stage0" not in only:
        json.dump(out, open(path, "w"))
        sys.exit(0)
    # cfg3, first stage, every feature
    import bench
    imgs, labels = bench.make_training_set_oracle()
    N = len(imgs)
    cores = os.cpu_count() or 8
    o = orc.Evaluator(orc.HAAR, orc.ALL, 24, 24, N)
    o.set_images(imgs, labels, threads=cores)
    t0 = time.time()
    td = orc.TrainData(o, N, orc.SORT_STD, cores)
    res = orc.boost_train(td, orc.boost_params(), cat_sort=orc.SORT_STD, record=True)
    w, m = res.snapshot(len(res.trees) - 1)
    out["cfg3_stage0"] = dict(N=N, F=td.F, trees=res.trees, threshold=res.threshold, log=res.log, seconds=time.time() - t0,
                              weights_sum=float(w.sum()), weights_head=[float(x) for x in w[:8]], mask_count=int((m != 0).sum()),
                              images_crc=int(np.bitwise_xor.reduce(imgs.astype(np.uint32).reshape(-1) * np.arange(1, imgs.size + 1, dtype=np.uint32))))
    print("cfg3_stage0", "%.1fs" % out["cfg3_stage0"]["seconds"], len(res.trees), res.trees[0][0], flush=True)
    with open(path, "w") as f:
        json.dump(out, f)
