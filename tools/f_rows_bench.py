"""Throughput of the SURVEY 8f kernels that are not on the measured path: 2-NN / Lowe-ratio extraction (lmpcr_nn_top2, scripts/extract_data.py)
and soft correspondences (lmpcr_nn_soft, the demo configuration), both exact CUDA-core kernels, next to the hard NN of the hot path.
python tools/f_rows_bench.py [--scans 16] [--points 5000] [--pairs 64]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synthdata
from util import cabi
ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=16); ap.add_argument("--points", type=int, default=5000); ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
feats, xyz, _ = synthdata.synth_scene(a.scans, a.points, seed=41)
f, x = torch.from_numpy(feats).cuda(), torch.from_numpy(xyz).cuda()
allp = [(i, j) for i in range(a.scans) for j in range(i + 1, a.scans)][: a.pairs]
pairs = torch.tensor(allp, dtype=torch.int32).cuda()
jobs = torch.cat([pairs, pairs.flip(1)], 0).contiguous()


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


P = len(allp)
for name, fn in (("hard NN, tcgen05 screening + exact rescoring (both directions)", lambda: cabi.nn_argmin(f, f, jobs, algo=cabi.NN_TENSOR)),
                 ("hard NN, exact CUDA-core kernel (both directions)", lambda: cabi.nn_argmin(f, f, jobs, algo=cabi.NN_EXACT_SIMT)),
                 ("2-NN + distances for the Lowe ratio, tcgen05 screening + exact rescoring (both directions)", lambda: cabi.nn_top2(f, f, jobs)),
                 ("2-NN + distances for the Lowe ratio, exact CUDA-core kernel (both directions)", lambda: cabi.nn_top2(f, f, jobs, algo=cabi.NN_EXACT_SIMT)),
                 ("soft correspondences, exact CUDA-core online softmax (both directions)", lambda: cabi.nn_soft(f, f, x, jobs, 0.09))):
    ms = timed(fn)
    print("%-84s %8.2f ms per %d pairs x %d pts = %7.1f us/pair, %8.0f pairs/s" % (name, ms, P, a.points, 1e3 * ms / P, P / ms * 1e3))
