"""Generates tests/golden/oanet_train_golden.npz: the UNMODIFIED reference OANet run in TRAINING mode (BatchNorm with batch
statistics -- the state scripts/benchmark_pairwise_registration.py leaves the model in, it never calls .eval()) on seeded
inputs: outputs of the forward pass and every BatchNorm buffer after it.

    python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refimport  # noqa: E402
import synthdata  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [("full_p4_n1000", 4, 1000, 26, False), ("small_p3_n96", 3, 96, 22, True), ("full_p1_n2000", 1, 2000, 23, False)]


def main():
    lib = refimport.import_reference()
    cfg = lib.utils.load_config(os.path.join(refimport.REFERENCE_ROOT, "configs/pairwise_registration/eval/RegBlock.yaml"))
    cfg["misc"]["use_gpu"] = False
    torch.set_num_threads(os.cpu_count())
    out = {}
    for name, P, N, seed, small in CASES:
        c = {k: dict(v) if isinstance(v, dict) else v for k, v in cfg.items()}
        kw = {}
        if small:
            c["misc"].update(net_channel=32, clusters=16)
            kw = dict(net_channel=32, clusters=16)
        net = lib.filtering.oanet.OANet(c)
        sd = synthdata.synth_state_dict(seed, **kw)
        net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        net.train()
        xs, _, _ = synthdata.synth_xs(P, N, seed=seed)
        with torch.no_grad():
            o = net({"xs": torch.from_numpy(xs)})
        out[name + "_cfg"] = np.array([P, N, seed, int(small)])
        for it in range(2):
            out["%s_logits%d" % (name, it)] = o["logits"][it].numpy()
            out["%s_R%d" % (name, it)] = o["rot_est"][it].numpy()
            out["%s_t%d" % (name, it)] = o["trans_est"][it].numpy()
        after = net.state_dict()
        keys = [k for k in after if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")]
        # buffers are stored concatenated in state_dict order to keep the file small
        out[name + "_bn_keys"] = np.array(keys)
        out[name + "_bn_vals"] = np.concatenate([after[k].numpy().astype(np.float64).reshape(-1) for k in keys])
        print(name, len(keys), out[name + "_bn_vals"].shape)
    np.savez_compressed(os.path.join(OUT, "oanet_train_golden.npz"), **out)
    print(os.path.getsize(os.path.join(OUT, "oanet_train_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
