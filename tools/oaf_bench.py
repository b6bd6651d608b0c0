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
