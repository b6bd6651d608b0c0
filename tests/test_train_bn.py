"""Training-mode BatchNorm (scripts/benchmark_pairwise_registration.py never calls .eval()):
oracle against a golden written by the reference OANet in .train() mode (CPU), CUDA path against both (GPU)."""
import os

import numpy as np
import pytest

from oracle import lmpcr_oracle as O
import synthdata

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oanet_train_golden.npz")
LOGIT_TOL, ROT_TOL, TRANS_TOL, BUF_TOL = 5e-4, 5e-4, 1e-3, 2e-5      # same gates as the eval-mode network tests; buffers: abs


def _cases():
    g = np.load(GOLD)
    for name in ("full_p4_n1000", "small_p3_n96", "full_p1_n2000"):
        P, N, seed, small = [int(v) for v in g[name + "_cfg"]]
        kw = dict(net_channel=32, clusters=16) if small else {}
        sd = synthdata.synth_state_dict(seed, **kw)
        xs, _, _ = synthdata.synth_xs(P, N, seed=seed)
        keys = [str(k) for k in g[name + "_bn_keys"]]
        vals, bufs, o = g[name + "_bn_vals"], {}, 0
        for k in keys:
            n = int(np.asarray(sd[k]).size)
            bufs[k] = vals[o:o + n].reshape(np.asarray(sd[k]).shape)
            o += n
        yield name, g, sd, xs, kw, bufs


def _check_outputs(name, g, logits, R, t):
    for it in range(2):
        assert np.max(np.abs(logits[it] - g["%s_logits%d" % (name, it)])) < LOGIT_TOL
        ang = O.chordal_angle(R[it], g["%s_R%d" % (name, it)])
        assert np.max(ang) < ROT_TOL
        assert np.max(np.abs(t[it] - g["%s_t%d" % (name, it)])) < TRANS_TOL


def _check_buffers(bufs, got):
    for k, want in bufs.items():
        if k.endswith("num_batches_tracked"):
            assert int(got[k]) == int(want) == 1
        else:
            assert np.max(np.abs(np.asarray(got[k], np.float64) - want)) < BUF_TOL, k


def test_oracle_train_mode_matches_reference():
    for name, g, sd, xs, kw, bufs in _cases():
        out = O.oanet_forward(xs, sd, dtype=np.float32, train=True)
        _check_outputs(name, g, out["logits"], out["rot_est"], out["trans_est"])
        assert set(out["bn_updates"]) == set(bufs)
        _check_buffers(bufs, out["bn_updates"])
        # the input state_dict is left untouched and eval mode differs from train mode
        assert int(np.asarray(sd[next(k for k in bufs if k.endswith("num_batches_tracked"))])) == 0
        ev = O.oanet_forward(xs, sd, dtype=np.float32)
        assert np.max(np.abs(ev["logits"][0] - out["logits"][0])) > 1e-3


@pytest.mark.gpu
def test_gpu_train_mode_matches_reference_and_updates_buffers():
    import torch
    from util import load_oanet
    for algo in (0, 1):
        for name, g, sd, xs, kw, bufs in _cases():
            net = load_oanet(sd, gemm_algo=algo, **kw)
            net.train()
            out = net({"xs": torch.from_numpy(xs)})
            logits = [v.cpu().numpy() for v in out["logits"]]
            R = [v.cpu().numpy() for v in out["rot_est"]]
            t = [v.cpu().numpy() for v in out["trans_est"]]
            _check_outputs(name, g, logits, R, t)
            after = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
            _check_buffers(bufs, after)
            # a second pass keeps counting and eval mode afterwards uses the updated running statistics
            net({"xs": torch.from_numpy(xs)})
            assert int(net.state_dict()[next(k for k in bufs if k.endswith("num_batches_tracked"))]) == 2
            net.eval()
            ev = net({"xs": torch.from_numpy(xs)})
            sd2 = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
            want = O.oanet_forward(xs, sd2, dtype=np.float64)
            assert np.max(np.abs(ev["logits"][-1].cpu().numpy() - want["logits"][-1])) < LOGIT_TOL
