#!/usr/bin/env python
"""bench.py — one "step" = one complete stage training (CvCascadeBoost::train) of BASELINE.json configs[2]:
Haar ALL 24x24, numPos=5000 numNeg=10000 (N=15000 samples, F=261600 features), reference default boosting
flags, on the stage-0 sample set (synthetic positives + the first 10000 windows of generated backgrounds).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  : feature evals/s through a whole stage = F*N / step time, samples already resident in HBM.
`e2e`    : same, through the C ABI with HOST buffers (H2D of the windows + D2H of the trained stage timed).
N > 1    : feature-sharded (each rank owns F/N variables; one NCCL all-gather of 64-byte candidates per node),
           total work fixed => "strong" scaling.
--impl reference : the CPU oracle (port of the reference; OpenCV C++ is not installable here) on a bounded
           sample (every k-th feature, k from the time budget or --ref-stride; 1 = the full F), all host threads for
           precalc/sort, serial split search like the reference; same `config` as the GPU arm.
`mining` : BASELINE configs[4]: the trained cascade over every window of 4096 distinct generated 1920x1080 backgrounds.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_POS, NUM_NEG, WIN = 5000, 10000, 24
HAAR_MODE_ALL = 2
WORKLOAD = "traincascade Haar ALL 24x24 stage train, numPos=5000 numNeg=10000 (BASELINE configs[2])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_sources():
    from cascadeclassifier_b200 import synth
    pos = synth.make_positives(NUM_POS, WIN, WIN, seed=12345)
    bgs = synth.make_backgrounds(6, 640, 480, seed0=1000, clutter=2.5)
    return pos, bgs


def make_training_set_gpu(ctx, ev):
    """Stage-0 sample set built by the product's own mining path (K5): positives from the .vec block, negatives =
    the first NUM_NEG windows of the NegReader stream over the generated backgrounds.  Returned to the host only
    for the e2e / cpu_baseline legs."""
    from cascadeclassifier_b200 import _lib
    pos, bgs = make_sources()
    lib, h = ctx.lib, ctx.h
    _lib.check(lib.cc_bg_clear(h), "cc_bg_clear")
    for b in bgs:
        b = np.ascontiguousarray(b)
        _lib.check(lib.cc_bg_register(h, b.ctypes.data_as(C.c_void_p), b.shape[1], b.shape[0]), "cc_bg_register")
    _lib.check(lib.cc_cascade_set_stages(h, None, 0), "cc_cascade_set_stages")
    consumed, got, vec_pos = C.c_int64(0), C.c_int32(0), C.c_int32(0)
    _lib.check(lib.cc_mine_positives(h, pos.ctypes.data_as(C.c_void_p), len(pos), NUM_POS, C.byref(vec_pos), C.byref(consumed), C.byref(got)), "mine_pos")
    st = _lib.ReaderState()
    _lib.check(lib.cc_mine_negatives(h, NUM_POS, NUM_NEG, 0.0, C.byref(st), C.byref(consumed), C.byref(got)), "mine_neg")
    assert got.value == NUM_NEG
    imgs = ev.images(0, NUM_POS + NUM_NEG)
    labels = np.concatenate([np.ones(NUM_POS, np.uint8), np.zeros(NUM_NEG, np.uint8)])
    return imgs, labels


def make_training_set_oracle():
    """Same set for the CPU arm, from the oracle's own PosReader / NegReader (bit-identical stream; see tests)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    pos, bgs = make_sources()
    rd = orc.NegReader(WIN, WIN, bgs)
    neg = np.stack([rd.get() for _ in range(NUM_NEG)])
    imgs = np.concatenate([pos[:NUM_POS], neg])
    labels = np.concatenate([np.ones(NUM_POS, np.uint8), np.zeros(NUM_NEG, np.uint8)])
    return imgs, labels


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    # BASELINE's other configs first (N = 1): whole cascades in their own contexts, before the headline config takes its 40 GB
    other = other_configs_leg(local) if world == 1 and not args.no_configs else None

    N = NUM_POS + NUM_NEG
    stream = torch.cuda.Stream(device=dev)
    ctx = ccb.Context(local, stream=stream.cuda_stream)
    lib, h = ctx.lib, ctx.h
    ev = ccb.CvFeatureEvaluator.create(ccb.CvFeatureParams.HAAR, ctx)
    ev.init(ccb.CvHaarFeatureParams(HAAR_MODE_ALL), N, (WIN, WIN))
    F = ev.getNumFeatures()
    imgs, labels = make_training_set_gpu(ctx, ev)
    if world > 1:
        from cascadeclassifier_b200.dist import init_comm
        init_comm(ctx, dist, dev)

    params = ccb.CvCascadeBoostParams()          # reference defaults: GAB, 0.995 / 0.5, trim 0.95, depth 1, 100 weak
    cparams = params.to_c()
    cap = params.weak_count
    trees = (_lib.Tree * cap)()
    logrows = (_lib.LogRow * cap)()
    stage = _lib.Stage(0.0, 0, 0, 0, trees, logrows)

    d_imgs = torch.from_numpy(imgs).to(dev)
    d_labels = torch.from_numpy(labels).to(dev)
    h_imgs = torch.from_numpy(imgs).pin_memory()
    h_labels = torch.from_numpy(labels).pin_memory()
    torch.cuda.synchronize(dev)

    def step_device():
        _lib.check(lib.cc_eval_set_images_device(h, C.c_void_p(d_imgs.data_ptr()), C.c_void_p(d_labels.data_ptr()), 0, N), "set_images_device")
        _lib.check(lib.cc_boost_train(h, N, C.byref(cparams), C.byref(stage)), "cc_boost_train")

    def step_host():
        _lib.check(lib.cc_eval_set_images(h, C.c_void_p(h_imgs.data_ptr()), C.c_void_p(h_labels.data_ptr()), 0, N), "set_images")
        _lib.check(lib.cc_boost_train(h, N, C.byref(cparams), C.byref(stage)), "cc_boost_train")
        return stage.num_weak, stage.threshold          # the trained stage is already on the host (D2H inside the call)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ctx.timing_reset()
    st0 = ctx.search_stats()
    launches0 = ctx.launch_count()
    ms_total = timed(step_device, args.steps)
    launches = ctx.launch_count() - launches0
    st1 = ctx.search_stats()
    searches_per_step = (st1["searches"] - st0["searches"]) / args.steps
    exact_rows_per_search = (st1["seed_rows"] + st1["flagged_rows"] - st0["seed_rows"] - st0["flagged_rows"]) / max(1, st1["searches"] - st0["searches"])
    fam = {}
    for name in ["integral", "precalc", "sort", "split_search", "boost_small"]:
        m, k = ctx.timing(name)
        fam[name] = {"ms_per_step": m / args.steps, "launches_per_step": k / args.steps}
    nweak = stage.num_weak
    clocks = sampler.finish() if sampler else None
    for _ in range(2):
        step_host()
    ms_e2e = timed(step_host, args.steps)

    mining = None
    if not args.no_mining:
        mining, mcc = mining_leg(local, dist if world > 1 else None, dev, world, nimg=args.mining_images)      # image-sharded across the ranks when N > 1
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            mining["cpu_baseline"] = mining_cpu_baseline(mcc.stageClassifiers)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = F * N / (ms_step * 1e-3)
    e2e_value = F * N / (ms_e2e / args.steps * 1e-3)
    peak, peak_kind = peaks()
    rows_local = F // world if world > 1 else F
    cells = float(rows_local) * N                      # (feature, sample) pairs of the local shard
    n_search = max(searches_per_step, 1e-9)
    ss_ms = fam["split_search"]["ms_per_step"] / n_search
    pre_ms, sort_ms = fam["precalc"]["ms_per_step"], fam["sort"]["ms_per_step"]
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    # algorithmic bytes per (feature, sample) pair (SURVEY.md 8d / DESIGN.md 4): precalc 4 B written; sort 4 B key read +
    # 2 B index written; split search: the pruned search reads the 2 B index row once per tree node (the f32 value row
    # is only read for the ~0.1 % of rows searched exactly)
    def roof(kernel, bytes_per_cell, ms, tkey):
        ach = bytes_per_cell * cells / (ms * 1e-3) / 1e9
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "traffic": (traffic[tkey] * rows_local / F) if tkey in traffic else None, "algorithmic_bytes_per_launch": bytes_per_cell * cells, "ms_per_launch": ms}
    rooflines = {"precalc": roof("k_precalc_haar_sep + k_precalc_haar", 4.0, pre_ms, "precalc_dram_bytes_per_step"),
                 "sort": roof("k_sort_rows_bucket (+ k_sort_rows on the rows it defers)", 6.0, sort_ms, "sort_dram_bytes_per_launch"),
                 "split_search": roof("k_ss_bound (+ exact search of seed / flagged rows)", 2.0, ss_ms, "split_search_dram_bytes_per_search")}
    dominant = max(("precalc", pre_ms), ("sort", sort_ms), ("split_search", fam["split_search"]["ms_per_step"]), key=lambda kv: kv[1])[0]
    out = {
        "metric": "Haar stage-train feature evals/s (F*N per complete stage train)", "value": value, "unit": "evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32 integral taps, f32 values, f64 split criteria",
        "data": "synthetic (generated glyph positives, generated backgrounds with distractors)",
        "config": bench_config(N, F), "parallelism": "feature-sharded x%d" % world, "weak_trees_in_stage": int(nweak),
        "first_split": [int(trees[0].nodes[0].var_idx), float(trees[0].nodes[0].c)],
        "stage_train_s": ms_step * 1e-3,
        "kernels": {"precalc_evals_per_s": cells / (pre_ms * 1e-3), "precalc_frac_of_hbm": rooflines["precalc"]["frac"],
                    "sort_ms": sort_ms, "sort_rows_deferred_to_radix": int(ctx.sort_deferred_rows()), "split_search_ms_per_node": ss_ms, "split_searches_per_step": searches_per_step,
                    "rows_searched_exactly_per_node": exact_rows_per_search, "families": fam, "rooflines": rooflines},
        "roofline": rooflines[dominant],
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(imgs.nbytes + labels.nbytes),
                "d2h_bytes_per_step": int(C.sizeof(_lib.Tree) * nweak + C.sizeof(_lib.LogRow) * nweak + 16), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if mining is not None:
        out["mining"] = mining
    if other is not None:
        out["other_configs"] = other
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(imgs, labels, F, budget_s=20.0, ref_trees=int(nweak))
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


OTHER_CONFIGS = {
    # BASELINE configs[0], [1], [3]: whole cascades through the host mirror of CvCascadeClassifier::train, reference default flags
    "configs[0]": dict(name="traincascade Haar BASIC 24x24, 2 stages, numPos=500 numNeg=1000", ft=0, haar_mode=0, W=24, H=24, npos=500, nneg=1000, nst=2),
    "configs[1]": dict(name="traincascade LBP 24x24, 10 stages, numPos=2000 numNeg=4000", ft=1, haar_mode=None, W=24, H=24, npos=2000, nneg=4000, nst=10),
    "configs[3]": dict(name="traincascade HOG 48x96, 15 stages, numPos=4000 numNeg=8000", ft=2, haar_mode=None, W=48, H=96, npos=4000, nneg=8000, nst=15),
}


def run_other_config(spec, device=0, nst=None):
    """One of BASELINE's other configs, end to end on synthetic data: wall time of train() (every stage until numStages or the
    required leaf false-alarm rate), split into mining and stage training, and the NegReader windows consumed."""
    import io
    import tempfile
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import synth
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    W, H = spec["W"], spec["H"]
    t0 = time.time()
    pos = synth.make_positives(int(spec["npos"] * 1.2) + 64, W, H, seed=12345)
    bgs = [synth.make_background(960, 720, 4000 + i, clutter=1.5) for i in range(24)]
    t_data = time.time() - t0
    ctx = ccb.Context(device)
    cc = CvCascadeClassifier(ctx, out=io.StringIO())
    fp = ccb.CvHaarFeatureParams(spec["haar_mode"]) if spec["ft"] == 0 else ccb.CvFeatureParams.create(spec["ft"])
    t0 = time.time()
    cc.train(tempfile.mkdtemp(), "pos.vec", "bg.txt", spec["npos"], spec["nneg"], 1024, 1024, nst or spec["nst"],
             CvCascadeParams(0, spec["ft"], (W, H)), fp, ccb.CvCascadeBoostParams(), positives=pos, backgrounds=bgs)
    ctx.synchronize()
    wall = time.time() - t0
    res = {"workload": spec["name"] + ", 24 generated 960x720 backgrounds", "stages_trained": len(cc.stageClassifiers),
           "weak_per_stage": [int(t["weak"]) for t in cc.timings], "train_wall_s": wall, "mining_s": sum(t["mine_s"] for t in cc.timings),
           "stage_training_s": sum(t["train_s"] for t in cc.timings), "neg_windows_consumed": int(sum(r["neg_consumed"] for r in cc.reports)),
           "gpu_launches": int(ctx.launch_count()), "data_generation_s": t_data, "timings": cc.timings, "reports": cc.reports}
    kern = {}
    for name in ["integral", "precalc", "sort", "split_search", "boost_small", "mine", "mine_host_assemble", "mine_host_eval", "mine_host_cut", "mine_host_alloc"]:
        ms, k = ctx.timing(name)
        kern[name] = {"ms": ms, "launches": k}
    res["families"] = kern
    ctx.close()
    return res


def other_configs_leg(device):
    out = {}
    for key, spec in OTHER_CONFIGS.items():
        r = run_other_config(spec, device)
        r.pop("timings"); r.pop("reports")
        out[key] = r
    return out


MINING_IMAGES, MINING_W, MINING_H = 4096, 1920, 1080
CASCADE_FIXTURE = os.path.join(ROOT, "tests", "golden", "cfg3_cascade_stages.json")


def load_fixture_cascade():
    """The cascade trained by BASELINE configs[2] (tools/full_cascade.py 20 5000 10000 48 --save-stages ...), kept as a fixture so
    that the mining sweep of configs[4] runs "a trained 20-stage cascade" without training one inside the bench."""
    from cascadeclassifier_b200.boost import CvCascadeBoost
    if not os.path.exists(CASCADE_FIXTURE):
        return None
    with open(CASCADE_FIXTURE) as f:
        d = json.load(f)
    return [CvCascadeBoost.from_trees(st["threshold"], st["trees"]) for st in d["stages"]], d.get("note", "")


def mining_leg(device, dist=None, dev=None, world=1, nimg=MINING_IMAGES, reps=3):
    """neg-mining windows/s on BASELINE configs[4]: fillPassedSamples' cascade evaluation (cc_mine_scan: every window evaluated,
    survivor cache not involved) of the trained Haar-ALL cascade over every window of 4096 distinct generated 1920x1080 backgrounds
    (13584 windows per image at round 0 = 5.56e7 windows per sweep), image-sharded across the ranks.  Backgrounds are generated on
    the GPU and stay resident in HBM (the product keeps decoded backgrounds resident).  Timed with CUDA events on the mining stream
    (max over ranks) around `reps` sweeps; wall clock reported next to it."""
    import io
    import torch
    import cascadeclassifier_b200 as ccb
    from cascadeclassifier_b200 import synth, _lib
    from cascadeclassifier_b200.cascade import CvCascadeClassifier, CvCascadeParams
    stream = torch.cuda.Stream(device=dev)
    ctx = ccb.Context(device, stream=stream.cuda_stream)
    lib, h = ctx.lib, ctx.h
    if dist is not None:
        from cascadeclassifier_b200.dist import init_comm
        init_comm(ctx, dist, dev)
    fixture = load_fixture_cascade()
    cc = CvCascadeClassifier(ctx, out=io.StringIO())
    if fixture is not None:
        stages, note = fixture
        cc.cascadeParams = CvCascadeParams(0, 0, (WIN, WIN))
        cc.featureEvaluator = ccb.CvFeatureEvaluator.create(ccb.CvFeatureParams.HAAR, ctx)
        cc.featureEvaluator.init(ccb.CvHaarFeatureParams(HAAR_MODE_ALL), 64, (WIN, WIN))
        cc.stageClassifiers = stages
        source = "fixture tests/golden/cfg3_cascade_stages.json (%s)" % note
    else:       # no fixture in the tree: a small cascade trained on the spot
        import tempfile
        pos = synth.make_positives(1400, WIN, WIN)
        small = synth.make_backgrounds(12, 480, 360, seed0=500, clutter=2.5)
        cc.train(tempfile.mkdtemp(prefix="bench_mining_cascade_r%d_" % device), "pos.vec", "bg.txt", 1000, 2000, 1024, 1024, 8, CvCascadeParams(0, 0, (WIN, WIN)),
                 ccb.CvHaarFeatureParams(HAAR_MODE_ALL), ccb.CvCascadeBoostParams(), positives=pos, backgrounds=small)
        source = "trained on the spot (fixture missing)"
    _lib.check(lib.cc_bg_clear(h), "cc_bg_clear")
    t_gen = time.time()
    with torch.cuda.stream(stream):
        for c0 in range(0, nimg, 256):
            n = min(256, nimg - c0)
            block = synth.make_backgrounds_device(n, MINING_W, MINING_H, 9000 + c0, dev)       # same seeds on every rank: same images
            for i in range(n):
                _lib.check(lib.cc_bg_register_device(h, C.c_void_p(block[i].data_ptr()), MINING_W, MINING_H, MINING_W), "cc_bg_register_device")
            ctx.synchronize()
            del block
    t_gen = time.time() - t_gen
    cc._sync_cascade()
    win, surv = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.cc_mine_scan(h, 0, nimg, C.byref(win), C.byref(surv)), "cc_mine_scan")      # warm-up (allocations)
    ctx.synchronize()
    ctx.timing_reset()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(reps):
        _lib.check(lib.cc_mine_scan(h, 0, nimg, C.byref(win), C.byref(surv)), "cc_mine_scan")
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()          # every sweep ends with a collective, so all ranks finish together; the barrier makes it explicit
    wall = (time.time() - t0) / reps
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_sweep = float(ms.item())
    ms_dev, k = ctx.timing("mine")
    # algorithmic bytes: every source pixel read once per image + every pyramid-level pixel produced once
    lvl_px = level_pixels(MINING_W, MINING_H, WIN, WIN)
    alg_bytes = float(nimg) * (MINING_W * MINING_H + lvl_px)
    peak, peak_kind = peaks()
    ach = alg_bytes / (ms_sweep * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("mining_dram_bytes_per_window")
    res = {"metric": "neg-mining windows/s (BASELINE configs[4]: %d generated %dx%d backgrounds, trained %d-stage Haar ALL cascade)" % (nimg, MINING_W, MINING_H, len(cc.stageClassifiers)),
           "windows_per_s": win.value / (ms_sweep * 1e-3), "windows": int(win.value), "survivors": int(surv.value), "images": nimg, "distinct_images": nimg,
           "image_size": "%dx%d" % (MINING_W, MINING_H), "stages": len(cc.stageClassifiers), "weak_per_stage": [len(s.trees) for s in cc.stageClassifiers],
           "cascade": source, "ms_per_sweep": ms_sweep, "wall_ms_per_sweep": wall * 1e3, "device_kernel_ms_per_sweep": ms_dev / reps,
           "launches_per_sweep": (ctx.launch_count() - launches0) / reps, "sweeps": reps, "background_generation_s": t_gen, "n_gpus": world,
           "roofline": {"kernel": "k_lvl_build + k_lvl_cascade (whole sweep)", "bound": "hbm", "achieved": ach / world, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s per GPU",
                        "frac": ach / world / peak, "algorithmic_bytes_per_window": alg_bytes / max(1, win.value),
                        "traffic": (traffic * win.value / world) if traffic else None,
                        "note": "reported, not gated: the miner is bound by instruction issue / shared memory, not by HBM (DESIGN.md K5)"},
           "sharding": "background images round-robin across ranks, survivor bitmap summed over NCCL" if world > 1 else "single GPU"}
    ctx.close()
    return res, cc


def level_pixels(w, h, W, H):
    """pixels of all pyramid levels of one w x h image at round 0 (NegReader::nextImg / get, src/imagestorage.cpp:57-126; f32 arithmetic)"""
    f = np.float32
    scale = max(f(W) / f(w), f(H) / f(h))
    total, first = 0, True
    while scale <= f(1.0):
        lw = int(scale * f(w) + f(0.5)) if first else int(scale * f(w))
        lh = int(scale * f(h) + f(0.5)) if first else int(scale * f(h))
        total += lw * lh
        first = False
        scale = f(scale * f(1.41421356))
    return total


def mining_cpu_baseline(stages, budget_windows=30000):
    """fillPassedSamples of the CPU oracle (serial, like the reference: one window at a time) on a bounded sample of the same
    workload: the first windows of one generated 1920x1080 background."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from cascadeclassifier_b200 import synth
    bg = synth.make_background(MINING_W, MINING_H, 9000, clutter=0.0)
    oc = orc.Cascade(orc.HAAR, orc.ALL, WIN, WIN, 1, 64, len(stages) + 1, orc.boost_params())
    oc.add_bg(bg)
    for st in stages:
        oc.push_stage(st.threshold, st.trees)
    t0 = time.time()
    got, consumed = oc.fill(0, 64, False, 1.0 / budget_windows)     # stops when (got + 1) / consumed <= 1 / budget
    dt = time.time() - t0
    return {"value": consumed / dt, "unit": "windows/s", "cores": 1, "kind": "port",
            "sample": "%d windows of one generated 1920x1080 background through the oracle's serial fillPassedSamples (%d accepted), %.2f s" % (consumed, got, dt)}


def bench_config(N, F):
    """identical in both arms (the driver compares it)"""
    return {"workload": WORKLOAD, "N": N, "F": F, "boost": "GAB, minHitRate 0.995, maxFalseAlarm 0.5, weightTrimRate 0.95, maxDepth 1, maxWeakCount 100",
            "sample_set": "stage 0: 5000 generated positives + the first 10000 NegReader windows of 6 generated 640x480 backgrounds",
            "timing": "inputs larger than L2 (fval 15.7 GB, sorted rows 23.5 GB per pass)"}


def cpu_stage_sample(o, N, F, stride, cores, ref_trees=None):
    """One stage training of the oracle (a port of the reference's CPU path) on every `stride`-th feature of the same workload:
    precalc + sort on all host threads (the cv::parallel_for_ shape of the reference, std::sort + LessThanIdx), serial split
    search (the reference's parallel_reduce is a serial stub, include/o_blockedrange.h:45-47).  The time is normalised to the number of
    weak trees the GPU arm grows on the full feature set (`ref_trees`), because a feature subsample can need another tree."""
    import orc
    rows = np.arange(0, F, stride, dtype=np.int32)
    t0 = time.time()
    td = orc.TrainData(o, N, orc.SORT_STD, cores, subset=rows if stride > 1 else None)
    t_pre = time.time() - t0
    res = orc.boost_train(td, orc.boost_params(), cat_sort=orc.SORT_STD)
    t_boost = res.seconds
    trees = len(res.trees)
    t_boost_norm = t_boost * (ref_trees / trees) if (ref_trees and trees) else t_boost
    del td
    return dict(rows=len(rows), t_pre=t_pre, t_boost=t_boost, t_boost_norm=t_boost_norm, trees=trees, first_var=int(rows[res.trees[0][0]["var_idx"]]) if stride > 1 else res.trees[0][0]["var_idx"],
                first_c=res.trees[0][0]["c"])


def cpu_baseline(imgs, labels, F, budget_s=20.0, ref_trees=None, stride=0):
    """cpu_baseline leg of the GPU arm: bounded sample sized to ~budget_s of CPU work."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    N = len(imgs)
    cores = os.cpu_count() or 1
    o = orc.Evaluator(orc.HAAR, orc.ALL, WIN, WIN, N)
    t0 = time.time()
    o.set_images(imgs, labels, threads=cores)
    t_int = time.time() - t0
    if stride <= 0:
        probe = cpu_stage_sample(o, N, F, 256, cores, ref_trees)
        per_row = (probe["t_pre"] + probe["t_boost"]) / probe["rows"]
        stride = 1
        while stride < 256 and per_row * (F / stride) > budget_s:
            stride *= 2
    r = cpu_stage_sample(o, N, F, stride, cores, ref_trees)
    total = t_int + r["t_pre"] + r["t_boost_norm"]
    return {"value": r["rows"] * N / total, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": "all %d samples, every %dth feature (%d of %d); integral %.2fs + precalc/std::sort %.2fs (%d threads) + boosting %.2fs serial for %d weak trees%s"
                      % (N, stride, r["rows"], F, t_int, r["t_pre"], cores, r["t_boost"], r["trees"],
                         "" if not ref_trees or r["trees"] == ref_trees else " (normalised to the %d trees of the full feature set: %.2fs)" % (ref_trees, r["t_boost_norm"])),
            "seconds": total, "stride": stride, "weak_trees": r["trees"], "first_split": [r["first_var"], r["first_c"]]}


GPU_ARM_TREES = 3      # weak trees the stage needs on the full feature set (GPU arm, and the oracle at stride 1: profiles/r02_reference_full_F.json)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference needs OpenCV C++ (absent from the image
    and from the GPU box: profiles/r02_opencv_probe.log), so this times the oracle port, with all host threads where the reference
    is threaded.  Each step is a bounded sample (every k-th feature, all samples); k is chosen so that the warmup + steps fit ~6 min."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    imgs, labels = make_training_set_oracle()
    N = len(imgs)
    import orc
    F = orc.lib().orc_catalog_count(orc.HAAR, orc.ALL, WIN, WIN)
    nsteps = max(1, args.warmup + args.steps)
    stride = args.ref_stride
    if stride <= 0:
        probe = cpu_baseline(imgs, labels, F, stride=128, ref_trees=GPU_ARM_TREES)
        per_full = probe["seconds"] * 128.0
        stride = 1
        while stride < 128 and per_full / stride * nsteps > 360.0:
            stride *= 2
    best, times = None, []
    for i in range(nsteps):
        cb = cpu_baseline(imgs, labels, F, stride=stride, ref_trees=GPU_ARM_TREES)
        if i >= args.warmup:
            times.append(cb["seconds"])
            best = cb
    sec = float(np.mean(times))
    value = best["value"] * best["seconds"] / sec
    best["value"] = value
    # ms_per_step is the time of one step = one bounded sample (F / stride features); value is the rate it sustains
    out = {"impl": "reference", "metric": "Haar stage-train feature evals/s (F*N per complete stage train)", "value": value, "unit": "evals/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "int32 integral taps, f32 values, f64 split criteria", "data": "synthetic (generated glyph positives, generated backgrounds with distractors)",
           "config": bench_config(N, F), "full_stage_ms_at_this_rate": F * N / value * 1e3,
           "cpu_baseline": best, "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mining", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[0] / [1] / [3] cascades (N = 1 only)")
    ap.add_argument("--mining-images", type=int, default=MINING_IMAGES)
    ap.add_argument("--ref-stride", type=int, default=0, help="--impl reference: every k-th feature (0 = chosen from the time budget, 1 = the full F)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
