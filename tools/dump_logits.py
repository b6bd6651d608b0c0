import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synthdata
from util import load_oanet
N, seed, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
sd = synthdata.synth_state_dict(seed)
xs, _, _ = synthdata.synth_xs(1, N, seed=seed)
net = load_oanet(sd, gemm_algo=1)
with torch.no_grad():
    o = net({"xs": torch.from_numpy(xs)})
np.savez(out, l0=o["logits"][0].cpu().numpy(), l1=o["logits"][1].cpu().numpy(), lat=o["latent features"].cpu().numpy()[0, :, ::50, 0])
