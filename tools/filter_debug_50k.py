import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import lmpcr_oracle as O
import synthdata
from util import cabi, cu, load_oanet
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sd = synthdata.synth_state_dict(SEED)
xs, _, _ = synthdata.synth_xs(1, N, seed=SEED)
o64 = O.oanet_forward(xs, sd, dtype=np.float64)
o32 = O.oanet_forward(xs, sd, dtype=np.float32)
print("numpy fp32 vs fp64: logits %.2e" % np.abs(o32["logits"][-1] - o64["logits"][-1]).max())
for algo in (0, 1):
    net = load_oanet(sd, gemm_algo=algo)
    with torch.no_grad():
        out = net({"xs": torch.from_numpy(xs)})
    for it in range(2):
        d = np.abs(out["logits"][it].cpu().numpy() - o64["logits"][it])
        print("algo %d block %d: logits max %.2e  p99.9 %.2e mean %.2e | rot %.2e trans %.2e" % (algo, it, d.max(), np.quantile(d, 0.999), d.mean(),
              O.chordal_angle(out["rot_est"][it].cpu().numpy(), o64["rot_est"][it]).max(), np.abs(out["trans_est"][it].cpu().numpy() - o64["trans_est"][it]).max()))
