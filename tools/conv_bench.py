"""Class-A layer alone: out = W relu(x*scale+shift) + bias (+ residual) through lmpcr_conv1x1 (tcgemm_kernel).
LMPCR_TC_DEBUG=256 additionally prints the per-role cycle counters of CTA 0."""
import sys, os, argparse, ctypes, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
cabi = importlib.import_module("3d_multiview_reg_b200")._cabi
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=296); ap.add_argument("--points", type=int, default=5000)
ap.add_argument("--cin", type=int, default=128); ap.add_argument("--cout", type=int, default=128)
ap.add_argument("--iters", type=int, default=9)
a = ap.parse_args()
dev = torch.device("cuda:0")
P, n, Ci, Co = a.pairs, a.points, a.cin, a.cout
w = (torch.randn(Co, Ci) / Ci ** 0.5).to(dev); b = torch.randn(Co).to(dev)
sc = torch.rand(P, Ci).to(dev) + 0.5; sh = (0.3 * torch.randn(P, Ci)).to(dev)
sets = 3
bufs = [(torch.randn(P, Ci, n, device=dev), torch.randn(P, Co, n, device=dev), torch.empty(P, Co, n, device=dev)) for _ in range(sets)]
ws = torch.empty(1 << 22, dtype=torch.uint8, device=dev)
names = ["mma wait T_EMPTY", "mma wait FULL", "mma issue", "prod wait EMPTY", "prod convert", "prod fence+arrive", "prod fetch/params",
         "epi wait T_FULL", "epi wait residual", "epi phase1", "epi stats", "epi phase2 stores"]
for use_res in (False, True):
    for x, r, o in bufs:
        cabi.conv1x1(x, w, b, sc, sh, r if use_res else None, gemm_algo=1, out=o, workspace=ws)
    torch.cuda.synchronize()
    if int(os.environ.get("LMPCR_TC_DEBUG", "0")) & 256:
        buf = (ctypes.c_ulonglong * 16)()
        cabi.load().lmpcr_debug_tc_profile(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        x, r, o = bufs[i % sets]
        cabi.conv1x1(x, w, b, sc, sh, r if use_res else None, gemm_algo=1, out=o, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    byts = ((Ci + Co + (Co if use_res else 0)) * P * n * 4)
    print("residual=%d: %.1f us per launch, %.0f GB/s algorithmic (%.1f%% of 6549)" % (use_res, 1e3 * ms, byts / ms / 1e6, byts / ms / 1e6 / 65.49))
    if int(os.environ.get("LMPCR_TC_DEBUG", "0")) & 256:
        cabi.load().lmpcr_debug_tc_profile(buf, 1)
        for i, nm in enumerate(names):
            print("   %-20s %10.3f Mcycles" % (nm, buf[i] / 1e6 / a.iters))
