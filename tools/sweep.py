"""BASELINE.json configs[4]: pairs/s over keypoints-per-scan x pairs-per-batch on one GPU.
Writes gpurun_out/sweep.json; one warm-up + two timed passes per cell (CUDA events), scene resident in HBM."""
import argparse, importlib, json, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, nargs="+", default=[1000, 2000, 5000, 10000, 20000])
ap.add_argument("--pairs", type=int, nargs="+", default=[64, 256, 1024, 4096])
ap.add_argument("--out", default="gpurun_out/sweep.json")
a = ap.parse_args()
pkg = importlib.import_module("3d_multiview_reg_b200")
scene = importlib.import_module("3d_multiview_reg_b200.scene")
oanet = importlib.import_module("3d_multiview_reg_b200.lib.filtering.oanet")
net = oanet.OANet({"misc": dict(iter_num=1, net_depth=12, net_channel=128, clusters=500, normalize_weights=True, use_gpu=True, gemm_algo=1),
                   "data": {"use_mutuals": 0}}).eval()
net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synthdata.synth_state_dict(41).items()}, strict=True)
net = net.cuda()
reg = scene.SceneRegistrar(net, nn_algo=pkg._cabi.NN_TENSOR)
rows = []
Pmax = max(a.pairs)
S = int(math.ceil((1 + math.sqrt(1 + 8 * Pmax)) / 2))
for n in a.points:
    feats, xyz, _ = synthdata.synth_scene(S, n, seed=41)
    feats, xyz = torch.from_numpy(feats).cuda(), torch.from_numpy(xyz).cuda()
    allp = torch.tensor([(i, j) for i in range(S) for j in range(i + 1, S)], dtype=torch.int32).cuda()
    for P in a.pairs:
        pairs = allp[:P].contiguous()
        reg.register_pairs(feats, xyz, pairs)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rec, _x = reg.register_pairs(feats, xyz, pairs); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ok = int((scene.unpack_records(rec)["status"] == 0).sum())
        rows.append({"points": n, "pairs": P, "ms": best, "pairs_per_s": P / best * 1e3, "us_per_pair": best / P * 1e3, "status_ok": ok})
        print(rows[-1], flush=True)
    del feats, xyz
    torch.cuda.empty_cache()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(a.out, "w"), indent=1)
