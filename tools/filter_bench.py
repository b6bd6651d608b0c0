"""Filter stage only (used under ncu for per-kernel times)."""
import sys, os, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import lmpcr_oracle as O
import synthdata
from util import cabi, load_oanet
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=64); ap.add_argument("--points", type=int, default=5000)
ap.add_argument("--algo", type=int, default=1); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
sd = synthdata.synth_state_dict(41)
xs, _, _ = synthdata.synth_xs(8, a.points, seed=41)
xs = torch.from_numpy(np.tile(xs, (a.pairs // 8, 1, 1, 1))).cuda()
net = load_oanet(sd, gemm_algo=a.algo)
cfg, params = net.cabi_cfg(), net.param_table()
cfg.guard_mode = cabi.GUARD_PAIR
for _ in range(2):
    cabi.filter_forward(xs, params, cfg, want_latent=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    cabi.filter_forward(xs, params, cfg, want_latent=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print("gemm_algo %d: %.2f ms per call, %.1f us/pair, %.1f TFLOP/s algorithmic" % (a.algo, ms, 1e3 * ms / a.pairs, a.pairs * 10.636e9 / ms / 1e9))

import ctypes
if int(os.environ.get("LMPCR_TC_DEBUG", "0")) & 256:
    buf = (ctypes.c_ulonglong * 16)()
    cabi.load().lmpcr_debug_tc_profile(buf, 1)
    cabi.filter_forward(xs, params, cfg, want_latent=False)
    cabi.load().lmpcr_debug_tc_profile(buf, 1)
    names = ["mma wait T_EMPTY", "mma wait FULL", "mma issue", "prod wait EMPTY", "prod convert", "prod fence+arrive", "prod fetch/params",
             "epi wait T_FULL", "epi wait residual", "epi phase1", "epi stats", "epi phase2 stores"]
    tot = 1.0
    for i, n in enumerate(names):
        print("%-20s %10.3f Mcycles" % (n, buf[i] / 1e6))
