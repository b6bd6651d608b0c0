"""diff_unpool's product alone: the pair-resident kernel (lmpcr_softmax_unpool mode 2) next to the generic GEMM (mode 1), kernel times
from the library's own CUDA-event brackets / torch events around the whole call.
python tools/unpool_bench.py [--pairs 296] [--points 2000] [--clusters 500] [--iters 5]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util import cabi
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=296); ap.add_argument("--points", type=int, default=2000)
ap.add_argument("--clusters", type=int, default=500); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--fused-only", action="store_true", help="skip the generic path (ncu captures)")
a = ap.parse_args()
C, P, N, K = 128, a.pairs, a.points, a.clusters
g = torch.Generator(device="cuda"); g.manual_seed(0)
xd = torch.randn(P, C, K, device="cuda", generator=g)
Es = [3 * torch.randn(P, K, N, device="cuda", generator=g) for _ in range(2)]


def timed(mode):
    cabi.softmax_unpool(xd, Es[0], mode); torch.cuda.synchronize()
    ev = []
    for i in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cabi.softmax_unpool(xd, Es[i % 2], mode); e1.record(); ev.append((e0, e1))
    torch.cuda.synchronize()
    return float(np.median([p.elapsed_time(q) for p, q in ev]))


cabi.ktime_enable(True)
ms2 = timed(2)
kt = cabi.ktime_read("unpool_fused_kernel")
print("ktime unpool_fused_kernel:", kt)
bytes_alg = P * (K * N + C * K + C * N + N) * 4.0
print("pair-resident: %d pairs x %d pts x %d clusters: %.3f ms per call (statistics pass included)" % (P, N, K, ms2))
if kt and kt[0]:
    us = 1e3 * kt[1] / kt[0]
    print("  kernel alone %.1f us per launch: %.0f GB/s algorithmic (E + x_down read, out written), %.1f TFLOP/s (3 bf16 products)"
          % (us, bytes_alg / us / 1e3, 3 * 2.0 * C * K * N * P / us / 1e6))
if not a.fused_only:
    ms1 = timed(1)
    print("generic GEMM (split_weights + deferred-softmax tcgemm): %.3f ms per call" % ms1)
