"""diff_pool's weighted sum alone against fp64, both normalisation modes (lmpcr_softmax_pool)."""
import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
cabi = importlib.import_module("3d_multiview_reg_b200")._cabi
for N in (2000, 5000, 50000):
    rng = np.random.default_rng(N)
    x = (rng.standard_normal((1, 128, N)) * 2 + 0.5).astype(np.float32)
    E = (rng.standard_normal((1, 500, N)) * 3).astype(np.float32)
    e64 = E.astype(np.float64)
    S = np.exp(e64 - e64.max(2, keepdims=True)); S /= S.sum(2, keepdims=True)
    ref = np.matmul(x.astype(np.float64), S.transpose(0, 2, 1))
    xd, Ed = torch.from_numpy(x).cuda(), torch.from_numpy(E).cuda()
    for mode in (0, 1):
        o = cabi.softmax_pool(xd, Ed, mode).cpu().numpy().astype(np.float64)
        d = o - ref
        print("N=%d mode %d: max %.2e mean %.2e signed %.2e rel-max %.2e" % (N, mode, np.abs(d).max(), np.abs(d).mean(), d.mean(), np.abs(d).max() / np.abs(ref).max()))
