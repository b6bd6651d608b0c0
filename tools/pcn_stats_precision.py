"""Error of the network against the fp64 oracle with the pair-resident kernels on (74 pairs x 2000 points): used to compare builds of pcn.cu whose
statistics pass uses one, two or three bf16 products (LMPCR_B200_LIB selects the build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import lmpcr_oracle as O
import synthdata
from util import cabi, cu, load_oanet
for seed in (3, 5):
    sd = synthdata.synth_state_dict(seed)
    xs, _, _ = synthdata.synth_xs(74, 2000, seed=seed)
    net = load_oanet(sd, gemm_algo=1)
    out = net({"xs": cu(xs)})
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    err = [np.abs(out["logits"][it].cpu().numpy() - o64["logits"][it]) for it in range(2)]
    rot = O.chordal_angle(out["rot_est"][-1].cpu().numpy(), o64["rot_est"][-1])
    print("seed %d: logits block 0 max %.2e mean %.2e | block 1 max %.2e mean %.2e | rot max %.2e rad median %.2e" %
          (seed, err[0].max(), err[0].mean(), err[1].max(), err[1].mean(), rot.max(), np.median(rot)))
