"""Times oaf_stack_kernel alone (CUDA event pairs around the launch: lmpcr_debug_ktime_*) on P pairs x 3 layers x 500 clusters.
Usage: python tools/oaf_bench.py [P] [layers] [reps]"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from test_gpu_oaf import _NAMES, _layer_sd  # noqa: E402

cabi = importlib.import_module("3d_multiview_reg_b200")._cabi
P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rng = np.random.default_rng(1)
K = 500
x = torch.from_numpy((rng.standard_normal((P, 128, K)) + 0.1).astype(np.float32)).cuda()
sds = [_layer_sd(rng, K) for _ in range(layers)]
params = [[torch.from_numpy(sd[m + "." + n]).cuda() for m, ns in _NAMES for n in ns] for sd in sds]
for _ in range(3):
    cabi.oafilter_stack(x, params)
torch.cuda.synchronize()
cabi.ktime_enable(True)
for _ in range(reps):
    cabi.oafilter_stack(x, params)
torch.cuda.synchronize()
n, ms = cabi.ktime_read("oaf_stack_kernel")
cabi.ktime_enable(False)
flop = (2.0 * 2 * 128 * 128 * K + 2.0 * 128 * K * K) * layers * P
print("oaf_stack_kernel: %d pairs x %d layers: %.1f us per launch (%d launches), %.1f us per pair and layer, %.1f TFLOP/s algorithmic"
      % (P, layers, 1e3 * ms / n, n, 1e3 * ms / n / layers / ((P + 147) // 148), flop / (ms / n * 1e-3) / 1e12))

if int(os.environ.get("LMPCR_OAF_DEBUG", "0")):
    import ctypes
    buf = (ctypes.c_ulonglong * 40)()
    cabi.load().lmpcr_debug_oaf_profile(buf, 1)
    cabi.oafilter_stack(x, params)
    cabi.load().lmpcr_debug_oaf_profile(buf, 1)
    names = {30: "phase_sync", 9: "conv1 start (first load, W1 -> TMEM)", 0: "conv1 wait x box", 1: "conv1 load + affine", 2: "conv1 wait MMADONE", 3: "conv1 operand image + fences",
             4: "conv1 set barrier + next load + arrive", 5: "conv1 epi: tc_ld + a = relu(bn_k(y)) -> smem / TMEM", 6: "conv1 epi: store drained + set barrier",
             7: "conv1 epi: stage + set barrier + TMA store", 8: "conv1 tail", 10: "conv2 wait D2FULL (= MMA phase)", 11: "conv2 epi: wait y box", 12: "conv2 epi: box + tc_ld + z + stats",
             13: "conv2 epi: store drained + set barrier + next load", 14: "conv2 epi: stage + set barrier + TMA store", 15: "conv2 epi: half end / stats",
             16: "conv2 ctl: wait WFULL", 17: "conv2 ctl: MMA issue + commit", 18: "conv2 ctl: wait WEMPTY + refill", 19: "conv2 ctl: wait EPIDONE",
             29: "conv3 start", 20: "conv3 wait z box", 21: "conv3 load + affine", 22: "conv3 wait MMADONE", 23: "conv3 operand image + fences", 24: "conv3 set barrier + next load + arrive",
             28: "conv3 epi: wait residual box", 25: "conv3 epi: box + tc_ld + out + stats", 26: "conv3 epi: store drained + set barrier + next load",
             27: "conv3 epi: stage + set barrier + TMA store", 31: "conv3 tail / stats", 32: "conv1/3 ctl: wait HFULL", 33: "conv1/3 ctl: MMA issue + commit"}
    per = layers * ((P + 147) // 148)
    tot_row = sum(buf[i] for i in names if i not in (16, 17, 18, 19, 32, 33))
    for i in names:
        print("%-58s %9.0f cycles per pair and layer" % (names[i], buf[i] / per))
    print("row thread total %.0f cycles per pair and layer" % (tot_row / per))
