"""Small end-to-end run for compute-sanitizer (memcheck): 3 scans x 512 points, both NN paths, both GEMM paths, soft mode."""
import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import lmpcr_oracle as O
import synthdata
from util import cabi, cu, load_oanet
scene = importlib.import_module("3d_multiview_reg_b200.scene")
feats, xyz, _ = synthdata.synth_scene(3, 516, seed=1)
sd = synthdata.synth_state_dict(1)
f, x = cu(feats), cu(xyz)
for g in (0, 1):
    net = load_oanet(sd, gemm_algo=g)
    for a in (0, 1):
        rec = scene.SceneRegistrar(net, nn_algo=a).register_scene(f, x)
        torch.cuda.synchronize()
        print("gemm", g, "nn", a, float(rec.abs().sum()))
jobs = torch.tensor([[0, 1]], dtype=torch.int32, device="cuda")
print("soft", float(cabi.nn_soft(f, f, x, jobs, 0.09).sum()))
xs, _, _ = synthdata.synth_xs(2, 301, seed=3)     # ragged N: generic epilogue paths
out = load_oanet(sd, gemm_algo=1)({"xs": torch.from_numpy(xs)})
torch.cuda.synchronize()
print("ragged ok", float(out["logits"][-1].sum()))
