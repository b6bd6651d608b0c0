#!/usr/bin/env python
"""bench.py -- scan pairs registered per second on the pairwise-registration hot path
(mutual-NN -> filtering network -> weighted Kabsch), BASELINE.json's metric on its configs[1]:
a 3DMatch-style scene of 60 fragments -> 1770 pairs at 5000 keypoints, 32-d features, per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scans S] [--points n]

One "step" = one pass of the whole path over the rank's pairs.  N>1: launched by torchrun, one rank per GPU; every
rank holds the scene, registers its contiguous slice of the lexicographic pair list (1770 pairs per rank: weak
scaling) and the step ends with the NCCL all-gather of the 16-float pose records.  Synthetic data
(synthdata.synth_scene, seed 41) and seeded random weights -- there is no network for datasets.

--impl reference: the reference's CPU implementation of the same path on the host cores, on a bounded sample of the same
workload: the UNMODIFIED reference code (Soft_NN('hard') x 2 -> construct_filtering_input_data -> OANet.forward)
when oracle/make_ref.sh has copied it to the git-ignored oracle/_ref (kind "reference"), else the oracle port (kind "port").

--workload dense50k: BASELINE configs[2] (50k x 50k keypoints per pair).  --scaling strong --scans 200: BASELINE configs[3]
(19,900 pairs split over the ranks in contiguous, padded shards; total work fixed).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scan pairs registered/sec (5k pts, 32-d FCGF)"
UNIT = "pairs/s"
NN_FLOP_PER_PAIR = lambda n, d: 2.0 * n * n * d                       # SURVEY.md 8d (one contraction, both directions)
FILTER_FLOP_PER_PAIR = lambda n: 2 * (n * 1005312.0 + 290.3e6) + n * 512.0  # SURVEY.md 8d, 2 blocks


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_sustained": d["bf16_tflops_sustained"], "tf_burst": d["bf16_tflops"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_sustained": 1400.0, "tf_burst": 1590.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def lex_pairs(S):
    """All scan pairs i < j in lexicographic order (lib/utils.py:873 itertools.combinations)."""
    return np.array([(i, j) for i in range(S) for j in range(i + 1, S)], dtype=np.int32).reshape(-1, 2)


def make_workload(S, n, seed=41):
    import synthdata                                     # neutral numpy generator (no oracle, no CUDA)
    feats, xyz, _ = synthdata.synth_scene(S, n, seed=seed)
    sd = synthdata.synth_state_dict(seed)
    return feats, xyz, sd


def cpu_port_pairs(feats, xyz, sd, pairs):
    """The oracle port of the path on the host cores for the given pairs.  Returns seconds."""
    from oracle import lmpcr_oracle as O
    from oracle import nn_c
    t0 = time.perf_counter()
    for a, b in pairs:
        i_st, _ = nn_c.nn_argmin(feats[a], feats[b])
        i_ts, _ = nn_c.nn_argmin(feats[b], feats[a])
        O.mutual_index(i_st, i_ts)
        xs = O.construct_xs(xyz[a], xyz[b][i_st])[None]
        O.oanet_forward(xs, sd, dtype=np.float32, guard="pair")
    return time.perf_counter() - t0


WORKLOAD = "configs[1]: %d scans -> %d pairs x %d keypoints x 32-d per GPU"


def ref_dir():
    d = os.path.join(ROOT, "oracle", "_ref")
    return d if os.path.isdir(os.path.join(d, "lib")) else None


class ReferencePath:
    """The reference's own functions for the path (imported from oracle/_ref, never from the product package)."""

    def __init__(self, sd):
        import torch
        import warnings
        warnings.filterwarnings("ignore", category=SyntaxWarning)          # the reference's docstrings carry "\m" escapes
        os.environ["LMPCR_REFERENCE_ROOT"] = ref_dir()
        from oracle import refimport
        import importlib
        importlib.reload(refimport)
        self.lib = refimport.import_reference()
        self.torch = torch
        torch.set_num_threads(os.cpu_count())
        cfg = self.lib.utils.load_config(os.path.join(ref_dir(), "configs", "pairwise_registration", "eval", "RegBlock.yaml"))
        cfg["misc"]["use_gpu"] = False
        self.net = self.lib.filtering.filtering_dict["oanet"](cfg).eval()
        self.net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            self.hard = self.lib.layers.Soft_NN(corr_type="hard", device="cpu")
        self.threads = torch.get_num_threads()

    def pairs(self, feats, xyz, pairs):
        torch = self.torch
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None]
        t0 = time.perf_counter()
        with torch.no_grad():
            for a, b in pairs:
                c_st = self.hard(t(feats[a]), t(feats[b]), t(xyz[b]))        # lib/pairwise/__init__.py:110
                self.hard(t(feats[b]), t(feats[a]), t(xyz[a]))               # :111
                fd = self.lib.utils.construct_filtering_input_data(t(xyz[a]), c_st, {}, False)   # :120
                self.net(fd)                                                 # lib/filtering/oanet.py:218 (both blocks + Kabsch)
        return time.perf_counter() - t0


def cpu_arm(sd):
    """(callable(feats, xyz, pairs) -> seconds, kind, threads, description)"""
    if ref_dir():
        try:
            r = ReferencePath(sd)
            return r.pairs, "reference", r.threads, "unmodified reference code from oracle/_ref (torch CPU, %d threads)" % r.threads
        except Exception as e:                                            # the copy is broken: say so and use the port
            sys.stderr.write("reference import failed (%s); using the oracle port\n" % e)
    from oracle import nn_c
    return (lambda f, x, p: cpu_port_pairs(f, x, sd, p)), "port", nn_c.threads(), "oracle port: C NN on %d threads + numpy network" % nn_c.threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, n = args.scans, args.points
    feats, xyz, sd = make_workload(S, n)
    all_pairs = lex_pairs(S)
    fn, kind, threads, how = cpu_arm(sd)
    t_pair = fn(feats, xyz, all_pairs[:1])                 # also the first warm-up
    budget = 150.0                                         # seconds for the whole --steps / --warmup run
    m = int(max(2, min(24, budget / ((args.steps + args.warmup) * max(t_pair, 1e-3)))))
    if args.ref_pairs > 0:
        m = args.ref_pairs
    for w in range(args.warmup):
        fn(feats, xyz, all_pairs[w * m % len(all_pairs):][:m])
    t = [fn(feats, xyz, all_pairs[(k * m) % (len(all_pairs) - m):][:m]) for k in range(args.steps)]
    sec = float(np.mean(t))
    val = m / sec
    sample = "%d consecutive pairs of the %d-scan x %d-point scene per step, %d steps (%s)" % (m, S, n, args.steps, how)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (S, S * (S - 1) // 2, n), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("3d_multiview_reg_b200")
    cabi = pkg._cabi
    scene = importlib.import_module("3d_multiview_reg_b200.scene")
    oanet = importlib.import_module("3d_multiview_reg_b200.lib.filtering.oanet")

    S, n = args.scans, args.points
    if args.workload == "dense50k":                      # BASELINE configs[2]: a few pairs of full voxelised fragments
        S, n = max(2, min(S, 4)) if args.scans != 60 else 3, (50000 if args.points == 5000 else args.points)
    strong = args.scaling == "strong"
    if strong:
        # BASELINE configs[3]: the scene's S(S-1)/2 pairs are FIXED and split over the ranks in contiguous, equally sized shards
        # (scene.partition_pairs: the last shard is short and padded in the all-gather)
        S_glob = S
        feats, xyz, sd = make_workload(S_glob, n)
        all_pairs = lex_pairs(S_glob)
        per_rank, ranges = scene.partition_pairs(len(all_pairs), world)
        my_pairs = all_pairs[ranges[rank][0]:ranges[rank][1]]
        pairs_total = len(all_pairs)
    else:
        per_rank = S * (S - 1) // 2                      # weak scaling: every rank registers this many pairs
        # the scene grows with the number of ranks so that the global lexicographic pair list has world*per_rank pairs
        S_glob = S
        while S_glob * (S_glob - 1) // 2 < world * per_rank:
            S_glob += 1
        feats, xyz, sd = make_workload(S_glob, n)
        all_pairs = lex_pairs(S_glob)[: world * per_rank]
        my_pairs = all_pairs[rank * per_rank:(rank + 1) * per_rank]
        pairs_total = world * per_rank
    n_mine = len(my_pairs)

    net = oanet.OANet({"misc": dict(iter_num=1, net_depth=12, net_channel=128, clusters=500, normalize_weights=True, use_gpu=True,
                                    gemm_algo=args.gemm_algo), "data": {"use_mutuals": 0}}).eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    net = net.to(dev)
    reg = scene.SceneRegistrar(net, nn_algo=args.nn_algo, pair_chunk=args.pair_chunk)

    f_host = torch.from_numpy(feats).pin_memory()
    x_host = torch.from_numpy(xyz).pin_memory()
    p_host = torch.from_numpy(np.ascontiguousarray(my_pairs)).pin_memory()
    rec_host = torch.empty((pairs_total, 16), dtype=torch.float32).pin_memory()
    f_dev, x_dev, p_dev = f_host.to(dev), x_host.to(dev), p_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_resident():
        mine, _ = reg.register_pairs(f_dev, x_dev, p_dev)
        if world > 1:
            return scene.all_gather_records(mine, per_rank, pairs_total, world)
        return mine

    def step_e2e():
        f = f_host.to(dev, non_blocking=True)
        x = x_host.to(dev, non_blocking=True)
        p = p_host.to(dev, non_blocking=True)
        mine, _ = reg.register_pairs(f, x, p)
        full = scene.all_gather_records(mine, per_rank, pairs_total, world) if world > 1 else mine
        rec_host.copy_(full, non_blocking=True)
        return full

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_timers=False):
        per = []
        if with_timers:
            reg.timers = {}
        barrier()
        for _ in range(steps):
            flush.zero_()                                              # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            per.append((e0, e1))
        barrier()
        ms = [a.elapsed_time(b) for a, b in per]
        stages = None
        if with_timers:
            stages = {k: float(np.sum([a.elapsed_time(b) for a, b in v])) / steps for k, v in reg.timers.items()}
            reg.timers = None
        tot = torch.tensor([float(np.sum(ms))], device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / steps, stages

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    l0 = cabi.launch_count()
    step_resident()
    launches_per_step = cabi.launch_count() - l0          # kernels of liblmpcr_b200 launched by one step (library counter)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    cabi.ktime_enable(True)          # CUDA event pairs around the dominant kernels' launches, on the launching stream, inside the timed region
    ms_step, stages = timed(step_resident, args.steps, with_timers=True)
    ktimes = {k: cabi.ktime_read(k) for k in ("pcn_stack_kernel", "pool_fused_kernel", "embed_fused_kernel", "unpool_fused_kernel", "oaf_stack_kernel", "nn_sweep_kernel", "nn_rescore_kernel")}
    cabi.ktime_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _ = timed(step_e2e, max(1, args.steps))

    if rank == 0:
        pk = peaks()
        value = pairs_total / (ms_step * 1e-3)
        nn_ms = stages["nn"]
        filt_ms = stages["filter"]
        nn_tf = n_mine * NN_FLOP_PER_PAIR(n, 32) / (nn_ms * 1e-3) / 1e12
        filt_tf = n_mine * FILTER_FLOP_PER_PAIR(n) / (filt_ms * 1e-3) / 1e12
        dom = live_pcn_roofline(ktimes["pcn_stack_kernel"], n_mine, n, args.steps, pk)
        if dom is None:              # the pair-resident kernels did not run (point count / group size outside their range): the per-layer GEMM instead
            dom = tcgemm_roofline(cabi, dev, n if n <= 8192 else 5000, pk)
        if args.workload == "dense50k":
            workload = "configs[2]: dense keypoints, %d scans -> %d pairs x %d keypoints x 32-d" % (S, pairs_total, n)
        elif strong:
            workload = "configs[3]: %d scans -> %d pairs x %d keypoints x 32-d, split over %d GPU(s) (shard %d, last shard %d)" % (
                S, pairs_total, n, world, per_rank, ranges[-1][1] - ranges[-1][0])
        else:
            workload = WORKLOAD % (S, per_rank, n)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "scene": "rank-0 scene of %d scans, %d pairs in total over %d rank(s)" % (S_glob, pairs_total, world),
                       "nn_algo": "tcgen05+rescore" if args.nn_algo == 1 else "exact_simt", "gemm_algo": "tcgen05 split-bf16" if args.gemm_algo == 1 else "fp32 simt",
                       "pair_chunk": args.pair_chunk, "arithmetic": "results in f32; NN screening fp16 operands -> f32 TMEM accumulators + exact f32 rescoring; GEMMs split-bf16 (hi+lo) -> f32", "l2": "256 MiB flush buffer written between timed iterations",
                       "parallelism": "pairs x%d" % world},
            # dominant kernel = pcn_stack_kernel (the pair-resident PointCN stacks, ~23 % of a step): every launch inside the timed steps
            # is bracketed by a CUDA event pair on the launching stream (lmpcr_debug_ktime_*); the per-layer GEMM it replaced is timed
            # alone beside it (roofline_tcgemm), the fused diff_pool kernel live like the dominant one (roofline_pool_fused)
            "roofline": dom,
            "roofline_tcgemm": tcgemm_roofline(cabi, dev, n if n <= 8192 else 5000, pk),
            "roofline_pool_fused": live_pool_roofline(ktimes["pool_fused_kernel"], n_mine, n, args.steps, pk),
            "roofline_embed_fused": live_embed_roofline(ktimes["embed_fused_kernel"], n_mine, n, args.steps, pk),
            "roofline_unpool_fused": live_unpool_roofline(ktimes["unpool_fused_kernel"], n_mine, n, args.steps, pk),
            "roofline_oaf": live_oaf_roofline(ktimes["oaf_stack_kernel"], n_mine, args.steps, pk),
            "roofline_filter_stage": {"bound": "tensor", "achieved": filt_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": filt_tf / pk["tf_sustained"],
                                      "algorithmic_flop_per_pair": FILTER_FLOP_PER_PAIR(n), "ms_per_step": filt_ms, "peak_source": pk["src"] + " bf16 sustained"},
            "roofline_nn": {"bound": "tensor", "achieved": nn_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": nn_tf / pk["tf_sustained"],
                            "frac_of_burst": nn_tf / pk["tf_burst"], "traffic": None, "kernel": "nn stage (both directions)",
                            "algorithmic_flop_per_pair": NN_FLOP_PER_PAIR(n, 32), "ms_per_step": nn_ms, "peak_source": pk["src"] + " bf16 sustained",
                            "sweep_ms_per_step": ktimes["nn_sweep_kernel"][1] / args.steps, "rescore_ms_per_step": ktimes["nn_rescore_kernel"][1] / args.steps,
                            "sweep_frac_of_sustained": (n_mine * NN_FLOP_PER_PAIR(n, 32) / max(ktimes["nn_sweep_kernel"][1] / args.steps, 1e-9) / 1e9) / pk["tf_sustained"]},
            "stages_ms": stages,
            "us_per_pair": {"nn": 1e3 * nn_ms / max(n_mine, 1), "filter": 1e3 * filt_ms / max(n_mine, 1), "total": 1e3 * ms_step / max(n_mine, 1)},
            "e2e": {"value": pairs_total / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(f_host.numel() * 4 + x_host.numel() * 4 + p_host.numel() * 4),
                    "d2h_bytes_per_step": int(rec_host.numel() * 4), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches_per_step * args.steps),        # kernels of liblmpcr_b200 launched in the timed region (library counter)
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            fn, kind, threads, how = cpu_arm(sd)
            cs = min(S_glob, 8)
            cp = lex_pairs(cs)[: args.cpu_pairs if n <= 8192 else 1]
            if n > 8192:                                   # the CPU paths materialise N x N matrices: bounded to a 5000-point subsample there
                cf, cx = feats[:cs, :5000], xyz[:cs, :5000]
            else:
                cf, cx = feats[:cs], xyz[:cs]
            fn(cf, cx, cp[:1])
            sec = fn(cf, cx, cp)
            line["cpu_baseline"] = {"value": len(cp) / sec, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": "%d pairs of the first %d scans of the same scene%s (%s), %.1f s"
                                              % (len(cp), cs, " at 5000 of the %d points" % n if n > 8192 else "", how, sec)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


PCN_PASSES_PER_PAIR = 2 * (3 * 3 + (3 * 2 - 1))      # per pair: 2 blocks x (l1_1: 3 layers x 3 passes + l1_2 tail: 2 layers, the last one's tile is not stored)


def live_pcn_roofline(kt, pairs, n, steps, pk, C=128):
    """pcn_stack_kernel as it ran inside the timed steps: kt = (launches, summed device milliseconds) from the event pairs around every
    launch.  HBM-bound: a layer streams its pair three times (statistics pass reads x; main pass reads x and writes the output), so the
    algorithmic bytes of a step are PCN_PASSES_PER_PAIR x pairs x C x n x 4 -- DESIGN.md 4.3.  `traffic` scales the ncu DRAM counters of
    one captured launch (profiles/r2_traffic.json: bytes moved / algorithmic bytes of that launch) to the average launch."""
    launches, ms = kt
    if launches == 0:
        return None
    byts = float(PCN_PASSES_PER_PAIR) * pairs * C * n * 4 * steps
    per_launch = byts / launches
    ms_launch = ms / launches
    gbs = byts / (ms * 1e-3) / 1e9
    traffic, note = None, None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get("pcn_stack_kernel")
        if t:
            traffic = per_launch * t["dram_bytes"] / t["algorithmic_bytes"]
            note = "ncu dram bytes / algorithmic bytes = %.3f on the captured launch (%s)" % (t["dram_bytes"] / t["algorithmic_bytes"], t["shape"])
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": traffic, "traffic_note": note,
            "kernel": "pcn_stack_kernel (pair-resident PointCN stacks: 3 fused layers after conv1, 2 + output head before the pose), timed live in the step",
            "algorithmic_bytes_per_launch": per_launch, "ms_per_launch": ms_launch, "launches_per_step": launches / steps,
            "ms_per_step": ms / steps, "peak_source": pk["src"] + " hbm copy"}


def live_pool_roofline(kt, pairs, n, steps, pk, C=128, K=500):
    """pool_fused_kernel (embedding conv + softmax + weighted sum of diff_pool) as it ran inside the timed steps.  Tensor-bound by its
    own arithmetic: 2 GEMMs of 2*K*C*n FLOP per pair and block, each executed as three bf16 products."""
    launches, ms = kt
    if launches == 0:
        return None
    flop = 2.0 * (2.0 * K * C * n) * 2 * pairs * steps          # algorithmic: 2 GEMMs x 2 blocks per pair
    tf = flop / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
            "executed_frac": 3.0 * tf / pk["tf_sustained"], "kernel": "pool_fused_kernel, timed live in the step",
            "ms_per_launch": ms / launches, "launches_per_step": launches / steps, "ms_per_step": ms / steps,
            "note": "executed tensor work is 3x the algorithmic FLOPs (split-bf16: hi.hi + hi.lo + lo.hi)", "peak_source": pk["src"] + " bf16 sustained"}


def live_embed_roofline(kt, pairs, n, steps, pk, C=128, K=500):
    """The `up` embedding conv on the pair-resident kernel (pool_fused.cu, POOL_EMBED) as it ran inside the timed steps.  HBM-bound by its
    output: it writes the [K, n] logits of every pair (K*n*4 bytes per pair and block) and reads the pair's tiles once from HBM."""
    launches, ms = kt
    if launches == 0:
        return None
    byts = (float(K) * n * 4 + float(C) * n * 4) * 2 * pairs * steps
    gbs = byts / (ms * 1e-3) / 1e9
    flop = 2.0 * K * C * n * 2 * pairs * steps
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "kernel": "pool_fused_kernel in embedding-conv mode, timed live in the step", "ms_per_launch": ms / launches,
            "launches_per_step": launches / steps, "ms_per_step": ms / steps, "tensor_tflops_algorithmic": flop / (ms * 1e-3) / 1e12,
            "note": "algorithmic bytes = the logits written + the pair's tiles read once (the other cluster blocks of a pair find them in L2)",
            "peak_source": pk["src"] + " hbm copy"}


def live_unpool_roofline(kt, pairs, n, steps, pk, C=128, K=500):
    """diff_unpool's product on the pair-resident kernel (unpool_fused.cu) as it ran inside the timed steps.  HBM-bound by its input: it
    reads the [K, n] logits of every pair once, x_down [C, K] once, and writes the [C, n] output."""
    launches, ms = kt
    if launches == 0:
        return None
    byts = (float(K) * n + float(C) * K + float(C) * n + n) * 4 * 2 * pairs * steps
    gbs = byts / (ms * 1e-3) / 1e9
    flop = 2.0 * K * C * n * 2 * pairs * steps
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "kernel": "unpool_fused_kernel, timed live in the step", "ms_per_launch": ms / launches,
            "launches_per_step": launches / steps, "ms_per_step": ms / steps, "tensor_tflops_algorithmic": flop / (ms * 1e-3) / 1e12,
            "note": "algorithmic bytes = logits + x_down + column maxima read, output written; 2 blocks per pair", "peak_source": pk["src"] + " hbm copy"}


def live_oaf_roofline(kt, pairs, steps, pk, C=128, K=500, layers=3):
    """The OAFilter stage on the pair-resident kernel (oaf.cu) as it ran inside the timed steps: per pair, block and layer two
    128 x 128 x K convolutions and one 128 x K x K cluster-mixing product.  Tensor-bound by executed work (three bf16
    products per fp32 product); its operands live in L2 / on chip, HBM sees the stack's input and output only."""
    launches, ms = kt
    if launches == 0:
        return None
    flop = (2.0 * 2 * C * C * K + 2.0 * C * K * K) * layers * 2 * pairs * steps
    tf = flop / (ms * 1e-3) / 1e12
    w2_bytes = float(512 * 512 * 4) * layers * 2 * pairs * steps
    return {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
            "executed_frac": 3 * tf / pk["tf_sustained"], "kernel": "oaf_stack_kernel, timed live in the step",
            "ms_per_launch": ms / launches, "launches_per_step": launches / steps, "ms_per_step": ms / steps,
            "l2_w2_stream_gbs": w2_bytes / (ms * 1e-3) / 1e9,
            "note": "executed tensor work is 3x the algorithmic FLOPs (split-bf16); l2_w2_stream_gbs = the bf16 hi/lo image of W2 (1 MB) streamed L2 -> SM once per pair and layer",
            "peak_source": pk["src"] + " bf16 sustained"}


def tcgemm_roofline(cabi, dev, n, pk, P=148, C=128, sets=3, iters=12):
    """The fused conv layer (tcgemm_kernel) alone: out = W * relu(x*scale+shift) + bias + residual over P pairs.
    Buffer sets are rotated (3 x 1.1 GB, far larger than L2) so that every launch streams from HBM.
    Algorithmic bytes per launch: read x + read residual + write out = 3 * P*C*n*4 (+ weights);
    algorithmic FLOPs: 2*C*C*n*P."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(41)
    w = (torch.randn(C, C, generator=g) / C ** 0.5).to(dev)
    b = torch.randn(C, generator=g).to(dev)
    sc = torch.rand(P, C, generator=g).to(dev) + 0.5
    sh = (0.3 * torch.randn(P, C, generator=g)).to(dev)
    bufs = [(torch.randn(P, C, n, device=dev), torch.randn(P, C, n, device=dev), torch.empty(P, C, n, device=dev)) for _ in range(sets)]
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    for x, r, o in bufs:
        cabi.conv1x1(x, w, b, sc, sh, r, gemm_algo=1, out=o, workspace=ws)
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        x, r, o = bufs[i % sets]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cabi.conv1x1(x, w, b, sc, sh, r, gemm_algo=1, out=o, workspace=ws)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b_) for a, b_ in evs]))
    byts = 3.0 * P * C * n * 4 + 2 * C * C * 4
    flops = 2.0 * C * C * n * P
    gbs = byts / (ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("tcgemm_conv128_res_148pairs_dram_bytes")
    del bufs
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": traffic,
            "kernel": "tcgemm_kernel<0> (fused 1x1 conv layer 128->128 + affine/ReLU prologue + residual), %d pairs x %d pts" % (P, n),
            "algorithmic_bytes_per_launch": byts, "ms_per_launch": ms, "tensor_tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
            "tensor_frac_of_sustained": flops / (ms * 1e-3) / 1e12 / pk["tf_sustained"], "peak_source": pk["src"] + " hbm copy"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans", type=int, default=60)
    ap.add_argument("--points", type=int, default=5000)
    ap.add_argument("--pair-chunk", type=int, default=296)
    ap.add_argument("--nn-algo", type=int, default=int(os.environ.get("LMPCR_NN_ALGO", "1")))
    ap.add_argument("--gemm-algo", type=int, default=int(os.environ.get("LMPCR_GEMM_ALGO", "1")))
    ap.add_argument("--ref-pairs", type=int, default=0, help="pairs per step of the reference arm (0: sized to a ~150 s run)")
    ap.add_argument("--workload", default="scene", choices=["scene", "dense50k"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-pairs", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
