#!/usr/bin/env python
"""bench.py -- scan pairs registered per second on the pairwise-registration hot path
(mutual-NN -> filtering network -> weighted Kabsch), BASELINE.json's metric on its configs[1]:
a 3DMatch-style scene of 60 fragments -> 1770 pairs at 5000 keypoints, 32-d features, per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scans S] [--points n]

One "step" = one pass of the whole path over the rank's pairs.  N>1: launched by torchrun, one rank per GPU; every
rank holds the scene, registers its contiguous slice of the lexicographic pair list (1770 pairs per rank: weak
scaling) and the step ends with the NCCL all-gather of the 16-float pose records.  Synthetic data
(synthdata.synth_scene, seed 41) and seeded random weights -- there is no network for datasets.

--impl reference: the reference's CPU implementation of the same path.  The reference is pure Python/torch and
cannot travel to the GPU box, so the arm runs the oracle port (oracle/: C for the NN arithmetic, numpy for the
network and Kabsch) on the host cores, on a bounded sample of the same workload.
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import lmpcr_oracle as O
    from oracle import nn_c
    S, n = args.scans, args.points
    feats, xyz, sd = make_workload(min(S, 8), n)         # a bounded sample only touches the first scans
    pairs = lex_pairs(min(S, 8))[: args.ref_pairs]
    for _ in range(args.warmup):
        cpu_port_pairs(feats, xyz, sd, pairs[:1])
    t = [cpu_port_pairs(feats, xyz, sd, pairs) for _ in range(args.steps)]
    sec = float(np.mean(t))
    val = len(pairs) / sec
    cores = os.cpu_count()
    sample = "%d pairs of the %d-scan x %d-point scene per step (oracle port: C NN on %d threads + numpy network)" % (len(pairs), S, n, nn_c.threads())
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: %d scans -> %d pairs x %d keypoints x 32-d per GPU" % (S, S * (S - 1) // 2, n), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    per_rank = S * (S - 1) // 2                          # weak scaling: every rank registers this many pairs
    # the scene grows with the number of ranks so that the global lexicographic pair list has world*per_rank pairs
    S_glob = S
    while S_glob * (S_glob - 1) // 2 < world * per_rank:
        S_glob += 1
    feats, xyz, sd = make_workload(S_glob, n)
    all_pairs = lex_pairs(S_glob)[: world * per_rank]
    my_pairs = all_pairs[rank * per_rank:(rank + 1) * per_rank]

    net = oanet.OANet({"misc": dict(iter_num=1, net_depth=12, net_channel=128, clusters=500, normalize_weights=True, use_gpu=True,
                                    gemm_algo=args.gemm_algo), "data": {"use_mutuals": 0}}).eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    net = net.to(dev)
    reg = scene.SceneRegistrar(net, nn_algo=args.nn_algo, pair_chunk=args.pair_chunk)

    f_host = torch.from_numpy(feats).pin_memory()
    x_host = torch.from_numpy(xyz).pin_memory()
    p_host = torch.from_numpy(np.ascontiguousarray(my_pairs)).pin_memory()
    rec_host = torch.empty((world * per_rank, 16), dtype=torch.float32).pin_memory()
    f_dev, x_dev, p_dev = f_host.to(dev), x_host.to(dev), p_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_resident():
        mine, _ = reg.register_pairs(f_dev, x_dev, p_dev)
        if world > 1:
            return scene.all_gather_records(mine, per_rank, world * per_rank, world)
        return mine

    def step_e2e():
        f = f_host.to(dev, non_blocking=True)
        x = x_host.to(dev, non_blocking=True)
        p = p_host.to(dev, non_blocking=True)
        mine, _ = reg.register_pairs(f, x, p)
        full = scene.all_gather_records(mine, per_rank, world * per_rank, world) if world > 1 else mine
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
    ms_step, stages = timed(step_resident, args.steps, with_timers=True)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _ = timed(step_e2e, max(1, args.steps))

    if rank == 0:
        pk = peaks()
        pairs_total = world * per_rank
        value = pairs_total / (ms_step * 1e-3)
        n_chunks = (per_rank + args.pair_chunk - 1) // args.pair_chunk
        nn_ms = stages["nn"]
        filt_ms = stages["filter"]
        nn_tf = per_rank * NN_FLOP_PER_PAIR(n, 32) / (nn_ms * 1e-3) / 1e12
        filt_tf = per_rank * FILTER_FLOP_PER_PAIR(n) / (filt_ms * 1e-3) / 1e12
        dom = dominant_kernel_roofline(cabi, dev, n, pk)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: %d scans -> %d pairs x %d keypoints x 32-d per GPU (rank-0 scene of %d scans, %d pairs total)"
                                   % (S, per_rank, n, S_glob, pairs_total),
                       "nn_algo": "tcgen05+rescore" if args.nn_algo == 1 else "exact_simt", "gemm_algo": "tcgen05 split-bf16" if args.gemm_algo == 1 else "fp32 simt",
                       "pair_chunk": args.pair_chunk, "arithmetic": "results in f32; NN screening fp16 operands -> f32 TMEM accumulators + exact f32 rescoring; GEMMs split-bf16 (hi+lo) -> f32", "l2": "256 MiB flush buffer written between timed iterations",
                       "parallelism": "pairs x%d" % world},
            # dominant kernel = tcgemm_kernel (the fused 1x1-conv layer; ~77 % of the step in the ncu launch list under
            # profiles/): timed alone, live, with CUDA events on a 128->128-channel layer with residual over 148 pairs.
            "roofline": dom,
            "roofline_filter_stage": {"bound": "tensor", "achieved": filt_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": filt_tf / pk["tf_sustained"],
                                      "algorithmic_flop_per_pair": FILTER_FLOP_PER_PAIR(n), "ms_per_step": filt_ms, "peak_source": pk["src"] + " bf16 sustained"},
            "roofline_nn": {"bound": "tensor", "achieved": nn_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": nn_tf / pk["tf_sustained"],
                            "frac_of_burst": nn_tf / pk["tf_burst"], "traffic": None, "kernel": "nn stage (both directions)",
                            "algorithmic_flop_per_pair": NN_FLOP_PER_PAIR(n, 32), "ms_per_step": nn_ms, "peak_source": pk["src"] + " bf16 sustained"},
            "stages_ms": stages,
            "e2e": {"value": pairs_total / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(f_host.numel() * 4 + x_host.numel() * 4 + p_host.numel() * 4),
                    "d2h_bytes_per_step": int(rec_host.numel() * 4), "ms_per_step": ms_e2e},
            "gpu_launches": None,
            "clocks": clocks,
        }
        # launches of OUR kernels per step (counted from the call structure: see DESIGN.md "launch count")
        line["gpu_launches"] = int(launches_per_step * args.steps)
        if world == 1 and not args.no_cpu_baseline:
            cf, cx, csd = feats[:8], xyz[:8], sd
            cp = lex_pairs(8)[: args.cpu_pairs]
            cpu_port_pairs(cf, cx, csd, cp[:1])
            sec = cpu_port_pairs(cf, cx, csd, cp)
            from oracle import nn_c
            line["cpu_baseline"] = {"value": len(cp) / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "%d pairs of the same scene (oracle port: C NN on %d threads + numpy network), %.1f s"
                                              % (len(cp), nn_c.threads(), sec)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(cabi, dev, n, pk, P=148, C=128, sets=3, iters=12):
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


def launches_per_chunk(args, per_rank):
    """Kernel launches of liblmpcr_b200 per pair chunk (one register_pairs iteration)."""
    nn = 2 * (2 if args.nn_algo == 0 else 4)           # per direction: sqnorm + argmin | prep + sweep + rescore (+memset)
    half = 3
    per_block = 1 + 1 + half * 4 + 2 + 1 + 1 + half * 6 + 2 + 1 + 1 + (5 + (half - 1) * 4) + 1   # see filter_net.cu
    filt = 2 * per_block + 2 * 2                        # + guard/kabsch per block
    return nn + 1 + filt + 1


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
    ap.add_argument("--ref-pairs", type=int, default=12)
    ap.add_argument("--cpu-pairs", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
