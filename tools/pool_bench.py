"""The fused diff_pool kernel alone: P pairs x N points x K clusters per launch, next to the GEMM path it replaces (convert_b +
embedding GEMM + row maxima + pooling GEMM, measured through lmpcr_conv1x1 + lmpcr_softmax_pool).
python tools/pool_bench.py [--pairs 296] [--points 5000] [--clusters 500] [--iters 5]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util import cabi, cu
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=296); ap.add_argument("--points", type=int, default=5000)
ap.add_argument("--clusters", type=int, default=500); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--single-only", action="store_true", help="skip the two-pass timing (ncu captures)")
a = ap.parse_args()
C, P, N, K = 128, a.pairs, a.points, a.clusters
g = torch.Generator(device="cuda"); g.manual_seed(0)
xs = [torch.randn(P, C, N, device="cuda", generator=g) for _ in range(2)]
sc = torch.rand(P, C, device="cuda", generator=g) + 0.3
sh = 0.5 * torch.randn(P, C, device="cuda", generator=g)
w = torch.randn(K, C, device="cuda", generator=g) * (3 / np.sqrt(C))


def timed(fn):
    fn(xs[0]); torch.cuda.synchronize()
    ev = []
    for i in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(xs[i % 2]); e1.record(); ev.append((e0, e1))
    torch.cuda.synchronize()
    return float(np.median([p.elapsed_time(q) for p, q in ev]))


if not a.single_only:
    ms2 = timed(lambda x: cabi.diff_pool_fused(x, sc, sh, w, 1))
    print("two-pass mode: %.3f ms per launch" % ms2)
ms = timed(lambda x: cabi.diff_pool_fused(x, sc, sh, w))
flop = 2.0 * 2 * K * C * N * P            # embedding conv + weighted sum
print("fused diff_pool: %d pairs x %d pts x %d clusters: %.3f ms per launch, %.2f us/pair, %.1f TFLOP/s algorithmic, x read at %.0f GB/s (2 passes)"
      % (P, N, K, ms, 1e3 * ms / P, flop / ms / 1e9, 2.0 * P * C * N * 4 / ms / 1e6))
if int(os.environ.get("LMPCR_POOL_DEBUG", "0")):
    import ctypes
    buf = (ctypes.c_ulonglong * 32)()
    cabi.load().lmpcr_debug_pool_profile(buf, 1)
    cabi.diff_pool_fused(xs[0], sc, sh, w)
    cabi.load().lmpcr_debug_pool_profile(buf, 1)
    names = {0: "load: issue", 1: "load: wait XFREE", 2: "A mma: issue", 3: "A mma: wait HFULL+EEMPTY", 4: "A max: work", 5: "A max: wait EFULL",
             8: "A prod: produce", 9: "A prod: wait XFULL+HEMPTY", 10: "B mma: issue GEMM2 (+commit)", 11: "B mma: wait HFULL+EEMPTY", 12: "B mma: issue GEMM1",
             13: "B mma: wait PFULL", 14: "B exp: st + arrive", 15: "B exp: wait EFULL", 16: "B exp: ld + exp + split", 17: "B exp: wait PEMPTY",
             18: "B exp: Z exchange", 19: "B exp: wait ACCFULL", 20: "B prod: produce", 21: "B prod: wait XFULL", 22: "B prod: wait HEMPTY+XBEMPTY",
             23: "loader: rest of item"}
    items = -(-(P * ((K + 127) // 128)) // 148)
    tiles = ((N + 63) // 64) * items
    for i in sorted(names):
        print("%-34s %9.1f cycles per 64-point tile" % (names[i], buf[i] / tiles))
bias = torch.randn(K, device="cuda", generator=g)
ms_e = timed(lambda x: cabi.embed_fused(x, sc, sh, w, bias, want_colmax=True))
print("embedding conv on the same kernel (E written, column maxima): %.3f ms per launch, E written at %.0f GB/s" % (ms_e, 1.0 * P * K * N * 4 / ms_e / 1e6))
if int(os.environ.get("LMPCR_POOL_DEBUG", "0")):
    import ctypes
    buf = (ctypes.c_ulonglong * 32)()
    cabi.load().lmpcr_debug_pool_profile(buf, 1)
    cabi.embed_fused(xs[0], sc, sh, w, bias, want_colmax=True)
    cabi.load().lmpcr_debug_pool_profile(buf, 1)
    names = {0: "load: issue", 1: "load: wait XFREE", 20: "prod: produce", 21: "prod: wait XFULL", 22: "prod: wait HEMPTY", 24: "mma: issue", 25: "mma: wait HFULL",
             26: "mma: wait EEMPTY", 27: "rd: loop tail", 28: "rd: wait EFULL", 29: "rd: ld + bias + colmax", 30: "rd: wait SFREE", 31: "rd: stage + fence"}
    items = -(-(P * ((K + 127) // 128)) // 148)
    tiles = ((N + 63) // 64) * items
    print("embedding-conv mode:")
    for i in sorted(names):
        print("%-34s %9.1f cycles per 64-point tile" % (names[i], buf[i] / tiles))
if os.environ.get("POOL_BENCH_REF", "1") == "1":
    # the path it replaces, through the stand-alone entry points: fused conv (affine + ReLU prologue) -> E, then the deferred-softmax pooling GEMM
    ms_conv = timed(lambda x: cabi.conv1x1(x, w, None, scale=sc, shift=sh, gemm_algo=1))
    E = cabi.conv1x1(xs[0], w, None, scale=sc, shift=sh, gemm_algo=1)
    ms_pool = timed(lambda x: cabi.softmax_pool(x, E, 1))
    print("GEMM path: embedding conv %.3f ms + softmax/pool %.3f ms = %.3f ms per launch (%.2fx)" % (ms_conv, ms_pool, ms_conv + ms_pool, (ms_conv + ms_pool) / ms))
