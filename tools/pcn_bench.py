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
    names = {0: "A load: issue", 1: "A load: wait XREAD", 2: "A mma: issue", 3: "A mma: wait H1FULL", 4: "A mma: wait YEMPTY", 5: "A stats: work",
             6: "A stats: wait YFULL", 7: "A prod(w10): produce+fence", 8: "A prod: wait XFULL", 9: "A prod: wait H1EMPTY", 10: "A loader tail (drain)",
             16: "B load: issue", 17: "B load: wait OUTRDY", 18: "B load: store+wait_read", 19: "B mma: issue GEMM2", 20: "B mma: wait H1FULL",
             21: "B mma: wait YEMPTY", 22: "B mma: issue GEMM1", 23: "B mma: wait H2FULL", 24: "B mma: wait ZEMPTY", 25: "B cvt: work", 26: "B cvt: wait YFULL",
             27: "B cvt: wait H2EMPTY", 28: "B epi(w7): work", 29: "B epi: wait ZFULL", 30: "B prod: produce+fence", 31: "B prod: wait XFULL",
             32: "B prod: wait H1EMPTY", 33: "B loader tail (drain)"}
    tiles = ((a.points + 63) // 64) * a.layers * ((a.pairs + 147) // 148)
    for i in sorted(names):
        print("%-34s %9.1f cycles per 64-point tile" % (names[i], buf[i] / tiles))
