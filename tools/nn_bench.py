"""NN stage only: 60 scans x 5000 pts, one chunk of pairs, both directions (used under ncu for per-kernel times)."""
import sys, os, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import lmpcr_oracle as O
import synthdata
from util import cabi
ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=60); ap.add_argument("--points", type=int, default=5000)
ap.add_argument("--pairs", type=int, default=256); ap.add_argument("--algo", type=int, default=1); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
feats, xyz, _ = synthdata.synth_scene(a.scans, a.points, seed=41)
f = torch.from_numpy(feats).cuda()
pairs = torch.from_numpy(O.enumerate_pairs(a.scans)[: a.pairs]).cuda()
jobs = torch.cat([pairs, pairs.flip(1)], 0).contiguous()
for _ in range(2):
    cabi.nn_argmin(f, f, jobs, algo=a.algo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cabi.ktime_enable(True)
e0.record()
for _ in range(a.iters):
    idx = cabi.nn_argmin(f, f, jobs, algo=a.algo)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
kt = {k: cabi.ktime_read(k) for k in ("nn_sweep_kernel", "nn_rescore_kernel")}
cabi.ktime_enable(False)
if kt["nn_sweep_kernel"][0]:
    tiles = 2.0 * a.pairs * ((a.points + 127) // 128) * ((a.points + 255) // 256) / 148
    sw = kt["nn_sweep_kernel"][1] / kt["nn_sweep_kernel"][0]
    print("  sweep %.3f ms per launch (%.0f ns per 128x256 tile and SM), rescore %.3f ms" % (sw, 1e6 * sw / tiles, kt["nn_rescore_kernel"][1] / max(kt["nn_rescore_kernel"][0], 1)))
print("algo %d: %.3f ms per call, %.2f us/pair, %.1f TFLOP/s algorithmic" % (a.algo, ms, 1e3 * ms / a.pairs, a.pairs * 2.0 * a.points ** 2 * 32 / ms / 1e9))
