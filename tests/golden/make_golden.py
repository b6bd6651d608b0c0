"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (imported from /root/reference through
oracle/refimport.py) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are not stored when they can be regenerated bit-identically from a seed with numpy's PCG64
(oracle.lmpcr_oracle.synth_scene / synth_xs / synth_state_dict); reference OUTPUTS are stored.
The reference has no golden vectors of its own, so these files are what pins the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lmpcr_oracle as O  # noqa: E402
from oracle import refimport  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def main():
    lib = refimport.import_reference()
    torch.manual_seed(41)
    torch.set_num_threads(os.cpu_count())

    # ---------------- stage 1: hard NN both ways, mutuals, xs ----------------
    nn = {}
    cases = [("s300x700", 300, 700, 11), ("s1000", 1000, 1000, 12), ("s5000", 5000, 5000, 41), ("s2049x777", 2049, 777, 13)]
    hard = lib.layers.Soft_NN(corr_type="hard", device="cpu")
    soft = lib.layers.Soft_NN(corr_type="soft", st=False, device="cpu")     # demo config: temp 0.3 -> T = 0.09
    for name, n, m, seed in cases:
        big = max(n, m)
        feats, xyz, _ = O.synth_scene(2, big, seed=seed)
        fs, ft, xs_, xt = feats[0, :n], feats[1, :m], xyz[0, :n], xyz[1, :m]
        with torch.no_grad():
            d_st = lib.utils.pairwise_distance(t(fs)[None], t(ft)[None])
            d_ts = lib.utils.pairwise_distance(t(ft)[None], t(fs)[None])
            nn[name + "_idx_st"] = d_st.min(dim=2)[1][0].numpy().astype(np.int32)
            nn[name + "_idx_ts"] = d_ts.min(dim=2)[1][0].numpy().astype(np.int32)
            nn[name + "_min_st"] = d_st.min(dim=2)[0][0].numpy()
            c_st = hard(t(fs)[None], t(ft)[None], t(xt)[None])            # lib/layers.py:44
            c_ts = hard(t(ft)[None], t(fs)[None], t(xs_)[None])
            assert np.array_equal(c_st[0].numpy(), xt[nn[name + "_idx_st"]])
            if n <= 1000:
                nn[name + "_soft_st"] = soft(t(fs)[None], t(ft)[None], t(xt)[None])[0].numpy()
                nn[name + "_soft_T"] = np.array(float(soft.get_temp().item()))
            if n == m and n <= 1000:
                mut = lib.utils.extract_mutuals(t(xs_)[None], t(xt)[None], c_st, c_ts)   # lib/utils.py:822
                nn[name + "_mutual_geo"] = mut[0].numpy().astype(np.uint8)
            fd = lib.utils.construct_filtering_input_data(t(xs_)[None], c_st, {}, False)  # lib/utils.py:888
            nn[name + "_xs_sum"] = np.array(fd["xs"].double().sum().item())
            assert tuple(fd["xs"].shape) == (1, 1, n, 6)
        nn[name + "_shape"] = np.array([n, m, seed])
    # exact ties: duplicated target rows -> first minimum must win
    feats, _, _ = O.synth_scene(2, 256, seed=5)
    ft = np.concatenate([feats[1][:128], feats[1][:128]], axis=0)
    d = lib.utils.pairwise_distance(t(feats[0])[None], t(ft)[None])
    nn["ties_idx"] = d.min(dim=2)[1][0].numpy().astype(np.int32)
    # un-normalised features (norm terms matter)
    rng = np.random.default_rng(99)
    fa = (rng.standard_normal((400, 32)) * 2).astype(np.float32)
    fb = (rng.standard_normal((600, 32)) * 0.5 + 0.3).astype(np.float32)
    d = lib.utils.pairwise_distance(t(fa)[None], t(fb)[None])
    nn["unnorm_idx"] = d.min(dim=2)[1][0].numpy().astype(np.int32)
    nn["unnorm_dist"] = d[0].numpy()[:64, :64].copy()
    np.savez_compressed(os.path.join(OUT, "nn_golden.npz"), **nn)

    # ---------------- stage 3: Kabsch ----------------
    kb = {}
    for name, P, N, seed, wmode in [("n3", 4, 3, 1, "ones"), ("n4", 4, 4, 2, "rand"), ("n50", 8, 50, 3, "rand"),
                                    ("n5000", 3, 5000, 4, "sparse"), ("n777", 5, 777, 5, "relu")]:
        xs, Rs, ts = O.synth_xs(P, N, inlier_frac=0.5, seed=seed)
        rng = np.random.default_rng(seed + 100)
        if wmode == "ones":
            w = np.ones((P, N), np.float32)
        elif wmode == "rand":
            w = rng.uniform(0, 1, (P, N)).astype(np.float32)
        elif wmode == "sparse":
            w = (rng.uniform(0, 1, (P, N)) * (rng.uniform(0, 1, (P, N)) < 0.1)).astype(np.float32)
        else:
            w = np.maximum(np.tanh(rng.standard_normal((P, N))), 0).astype(np.float32)
        R, tt, res, flag = lib.utils.kabsch_transformation_estimation(t(xs[:, 0, :, :3]), t(xs[:, 0, :, 3:]), t(w))
        kb[name + "_cfg"] = np.array([P, N, seed])
        kb[name + "_w"] = w
        kb[name + "_R"] = R.numpy()
        kb[name + "_t"] = tt.numpy()
        kb[name + "_res"] = res.numpy()
        r2 = lib.utils.transformation_residuals(t(xs[:, 0, :, :3]), t(xs[:, 0, :, 3:]), R, tt)
        assert np.array_equal(r2.numpy(), res.numpy())
    np.savez_compressed(os.path.join(OUT, "kabsch_golden.npz"), **kb)

    # ---------------- stage 2+3: OANet forward (eval mode) ----------------
    cfg = lib.utils.load_config(os.path.join(refimport.REFERENCE_ROOT, "configs/pairwise_registration/eval/RegBlock.yaml"))
    cfg["misc"]["use_gpu"] = False
    oa = {}
    for name, P, N, seed, small, guard in [("full_p2_n2000", 2, 2000, 7, False, False), ("full_p1_n5000", 1, 5000, 8, False, False),
                                           ("small_p3_n64", 3, 64, 9, True, False), ("guard_p2_n500", 2, 500, 10, False, True)]:
        c = {k: dict(v) if isinstance(v, dict) else v for k, v in cfg.items()}
        kw = {}
        if small:
            c["misc"].update(net_channel=32, clusters=16)
            kw = dict(net_channel=32, clusters=16)
        net = lib.filtering.oanet.OANet(c).eval()
        sd = O.synth_state_dict(seed, **kw)
        if guard:   # all logits negative in the first block -> sum(w) == 0 -> +1/N guard (oanet.py:177-178)
            sd["reg_init.output.bias"] = np.full((1,), -50.0, np.float32)
        net.load_state_dict({k: t(np.asarray(v)) for k, v in sd.items()}, strict=True)
        xs, _, _ = O.synth_xs(P, N, seed=seed)
        with torch.no_grad():
            out = net({"xs": t(xs)})
        oa[name + "_cfg"] = np.array([P, N, seed, int(small), int(guard)])
        for it in range(2):
            oa["%s_logits%d" % (name, it)] = out["logits"][it].numpy()
            oa["%s_scores%d" % (name, it)] = out["scores"][it].numpy()
            oa["%s_R%d" % (name, it)] = out["rot_est"][it].numpy()
            oa["%s_t%d" % (name, it)] = out["trans_est"][it].numpy()
        oa[name + "_latent_absmean"] = np.array(out["latent features"].abs().mean().item())
        oa[name + "_flag"] = np.array(bool(out["gradient_flag"]))
    np.savez_compressed(os.path.join(OUT, "oanet_golden.npz"), **oa)
    for f in ("nn_golden.npz", "kabsch_golden.npz", "oanet_golden.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
