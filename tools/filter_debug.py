"""Compare the filter network's two GEMM back ends against the fp64 oracle, layer group by layer group (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import lmpcr_oracle as O
import synthdata
from util import cabi, cu, load_oanet
for (P, N, kw, seed) in [(3, 64, dict(net_channel=32, clusters=16), 9), (2, 2000, {}, 7)]:
    sd = synthdata.synth_state_dict(seed, **kw)
    xs, _, _ = synthdata.synth_xs(P, N, seed=seed)
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    for algo in (0, 1):
        net = load_oanet(sd, gemm_algo=algo, **kw)
        with torch.no_grad():
            out = net({"xs": torch.from_numpy(xs)})
        torch.cuda.synchronize()
        for it in range(2):
            dl = np.abs(out["logits"][it].cpu().numpy() - o64["logits"][it]).max()
            dr = O.chordal_angle(out["rot_est"][it].cpu().numpy(), o64["rot_est"][it]).max()
            dt = np.abs(out["trans_est"][it].cpu().numpy() - o64["trans_est"][it]).max()
            print("P=%d N=%d algo=%d block %d: logits %.2e  rot %.2e rad  trans %.2e m" % (P, N, algo, it, dl, dr, dt))
        lat = np.abs(out["latent features"][..., 0].cpu().numpy() - o64["latent features"][..., 0]).max()
        print("   latent max err %.2e" % lat)
