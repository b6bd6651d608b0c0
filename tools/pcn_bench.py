"""The PointCN stack kernel alone: P pairs x N points x `layers` layers per launch (time per launch, per layer-pass and HBM GB/s
against the 3-pass algorithmic traffic).  python tools/pcn_bench.py [--pairs 296] [--points 5000] [--layers 3] [--iters 5]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util import cabi, cu
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=296); ap.add_argument("--points", type=int, default=5000)
ap.add_argument("--layers", type=int, default=3); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
rng = np.random.default_rng(0)
C = 128
def lp():
    bn = lambda: [rng.uniform(0.5, 1.5, C), 0.3 * rng.standard_normal(C), 0.2 * rng.standard_normal(C), rng.uniform(0.5, 1.5, C)]
    w = lambda: rng.standard_normal((C, C)) / np.sqrt(C)
    return [cu(t.astype(np.float32)) for t in bn() + [w(), 0.1 * rng.standard_normal(C)] + bn() + [w(), 0.1 * rng.standard_normal(C)]]
layers = [lp() for _ in range(a.layers)]
bufs = [(torch.randn(a.pairs, C, a.points, device="cuda"), torch.empty(a.pairs, C, a.points, device="cuda")) for _ in range(2)]
for x, o in bufs:
    cabi.pointcn_stack(x, layers, out=o)
torch.cuda.synchronize()
ev = []
for i in range(a.iters):
    x, o = bufs[i % 2]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cabi.pointcn_stack(x, layers, out=o); e1.record(); ev.append((e0, e1))
torch.cuda.synchronize()
ms = float(np.median([p.elapsed_time(q) for p, q in ev]))
byts = 3.0 * a.layers * a.pairs * C * a.points * 4
print("pcn stack: %d pairs x %d pts x %d layers: %.3f ms per launch (incl. weight split + input statistics), %.1f us per layer, %.0f GB/s of 3-pass traffic"
      % (a.pairs, a.points, a.layers, ms, 1e3 * ms / a.layers, byts / ms / 1e6))

if int(os.environ.get("LMPCR_PCN_DEBUG", "0")):
    import ctypes
    buf = (ctypes.c_ulonglong * 40)()
    cabi.load().lmpcr_debug_pcn_profile(buf, 1)
    cabi.pointcn_stack(bufs[0][0], layers, out=bufs[0][1])
    cabi.load().lmpcr_debug_pcn_profile(buf, 1)
    names = {0: "loadA wait XREAD", 1: "loadA issue", 2: "loadB wait OUTRDY", 3: "loadB store+wait_read", 4: "loadB issue", 5: "mma wait H1FULL", 6: "mma wait YEMPTY",
             7: "mma issue1(A) / after-issue1(B)", 8: "mma wait H2FULL", 9: "mma wait ZEMPTY", 10: "mma issue2", 11: "statsA wait YFULL", 12: "statsA work",
             13: "cvtB wait YFULL", 14: "cvtB ld+f2", 15: "cvtB wait H2EMPTY", 16: "cvtB split+store+fence", 17: "epiB wait ZFULL", 19: "epiB work",
             20: "prodA wait XFULL", 21: "prodA ld+f1", 22: "prodA wait H1EMPTY", 23: "prodA split+store+fence", 24: "prodB wait XFULL", 25: "prodB ld+f1",
             26: "prodB wait H1EMPTY", 27: "prodB split+store+fence", 28: "loader: pass A tail", 29: "loader: pass B tail"}
    tiles = ((a.points + 31) // 32) * a.layers * ((a.pairs + 147) // 148)
    for i in sorted(names):
        print("%-34s %9.1f cycles per tile" % (names[i], buf[i] / tiles))
