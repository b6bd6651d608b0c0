"""GPU parity, stage 2(+3): the filtering network through lmpcr_filter_forward vs the fp64 oracle and the reference
goldens.  Gates (DESIGN.md "tolerances"): logits/scores 5e-4, rotation 5e-4 rad (chordal), translation 1e-3 m --
5x the reference's own fp32-vs-fp64 spread on these inputs; the Kabsch stage alone is gated at 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from util import cabi, cu, load_oanet

pytestmark = pytest.mark.gpu
LOGIT_TOL, ROT_TOL, TRANS_TOL = 5e-4, 5e-4, 1e-3


def _run(net, xs):
    with torch.no_grad():
        return net({"xs": torch.from_numpy(xs)})          # CPU input, moved inside like oanet.py:234


GEMM_ALGOS = [0, 1]      # 0: fp32 CUDA cores, 1: tcgen05 split-bf16 tensor cores


@pytest.mark.parametrize("algo", GEMM_ALGOS)
@pytest.mark.parametrize("name", ["full_p2_n2000", "full_p1_n5000", "small_p3_n64", "guard_p2_n500"])
def test_oanet_vs_reference_golden(golden_dir, name, algo):
    g = np.load(os.path.join(golden_dir, "oanet_golden.npz"))
    P, N, seed, small, guard = [int(v) for v in g[name + "_cfg"]]
    kw = dict(net_channel=32, clusters=16) if small else {}
    sd = O.synth_state_dict(seed, **kw)
    if guard:
        sd["reg_init.output.bias"] = np.full((1,), -50.0, np.float32)
    xs, _, _ = O.synth_xs(P, N, seed=seed)
    out = _run(load_oanet(sd, gemm_algo=algo, **kw), xs)
    assert set(["logits", "scores", "rot_est", "trans_est", "latent features", "gradient_flag"]) <= set(out.keys())
    assert len(out["logits"]) == 2 and tuple(out["rot_est"][-1].shape) == (P, 3, 3) and tuple(out["trans_est"][-1].shape) == (P, 3, 1)
    assert tuple(out["latent features"].shape) == (P, kw.get("net_channel", 128), N, 1)
    for it in range(2):
        assert np.abs(out["logits"][it].cpu().numpy() - g["%s_logits%d" % (name, it)]).max() < LOGIT_TOL
        assert np.abs(out["scores"][it].cpu().numpy() - g["%s_scores%d" % (name, it)]).max() < LOGIT_TOL
        assert O.chordal_angle(out["rot_est"][it].cpu().numpy(), g["%s_R%d" % (name, it)]).max() < ROT_TOL
        assert np.abs(out["trans_est"][it].cpu().numpy() - g["%s_t%d" % (name, it)]).max() < TRANS_TOL
    assert out["gradient_flag"] == bool(g[name + "_flag"])
    assert abs(float(out["latent features"].abs().mean()) - float(g[name + "_latent_absmean"])) < 1e-3
    if guard:
        assert np.allclose(out["scores"][0].cpu().numpy(), 1.0 / N)


@pytest.mark.parametrize("algo", GEMM_ALGOS)
@pytest.mark.parametrize("P,N,seed", [(3, 777, 31), (2, 1001, 32), (1, 130, 33)])
def test_oanet_vs_fp64_oracle_ragged_sizes(P, N, seed, algo):
    """Point counts that are not multiples of 4 / of the tile sizes (mutual-filtered inputs have arbitrary N)."""
    sd = O.synth_state_dict(seed)
    xs, _, _ = O.synth_xs(P, N, seed=seed)
    out = _run(load_oanet(sd, gemm_algo=algo), xs)
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    for it in range(2):
        assert np.abs(out["logits"][it].cpu().numpy() - o64["logits"][it]).max() < LOGIT_TOL
        assert O.chordal_angle(out["rot_est"][it].cpu().numpy(), o64["rot_est"][it]).max() < ROT_TOL
        assert np.abs(out["trans_est"][it].cpu().numpy() - o64["trans_est"][it]).max() < TRANS_TOL
    # stage 3 on the net's own weights: 1e-5 gate
    w = out["scores"][-1].cpu().numpy()
    Ro, to, reso, _ = O.kabsch(xs[:, 0, :, :3], xs[:, 0, :, 3:], w, dtype=np.float64)
    assert O.chordal_angle(out["rot_est"][-1].cpu().numpy(), Ro).max() < 1e-5
    assert np.abs(out["trans_est"][-1].cpu().numpy() - to).max() < 1e-5
    assert np.abs(out["residuals"].cpu().numpy() - reso).max() < 1e-5


@pytest.mark.parametrize("algo", GEMM_ALGOS)
def test_group_size_does_not_change_results(algo):
    """Pairs are processed in workspace-sized groups; per-pair arithmetic must not depend on the grouping."""
    sd = O.synth_state_dict(5)
    xs, _, _ = O.synth_xs(5, 512, seed=5)
    net = load_oanet(sd, gemm_algo=algo)
    cfg = net.cabi_cfg()
    x = cu(xs)
    full = cabi.filter_forward(x, net.param_table(), cfg)
    one_pair = cabi.filter_workspace_bytes(cfg, 1, 512)
    small = cabi.filter_forward(x, net.param_table(), cfg,
                                workspace=torch.empty(one_pair + 5 * 512 * 4 + 8192, dtype=torch.uint8, device="cuda"))
    for k in ("logits", "scores", "R", "t", "residuals", "latent", "conf"):
        assert torch.equal(full[k], small[k]), k


def test_side_channel_and_iter0():
    """use_mutuals == 2 (7-channel xs) and iter_num = 0 (single block of depth 12)."""
    sd = O.synth_state_dict(9, side_channel=1)
    xs, _, _ = O.synth_xs(2, 300, seed=9)
    xs7 = np.concatenate([xs, (np.random.default_rng(1).uniform(size=(2, 1, 300, 1)) > 0.5).astype(np.float32)], axis=3)
    import importlib
    oanet = importlib.import_module("3d_multiview_reg_b200.lib.filtering.oanet")
    net = oanet.OANet({"misc": dict(iter_num=1, net_depth=12, net_channel=128, clusters=500, normalize_weights=True, use_gpu=True),
                       "data": {"use_mutuals": 2}}).eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    out = _run(net.cuda(), xs7)
    o64 = O.oanet_forward(xs7, sd, dtype=np.float64)
    assert np.abs(out["logits"][-1].cpu().numpy() - o64["logits"][-1]).max() < LOGIT_TOL
    sd0 = O.synth_state_dict(10, iter_num=0)
    out0 = _run(load_oanet(sd0, iter_num=0), xs)
    o0 = O.oanet_forward(xs, sd0, iter_num=0, dtype=np.float64)
    assert len(out0["logits"]) == 1
    assert np.abs(out0["logits"][0].cpu().numpy() - o0["logits"][0]).max() < LOGIT_TOL


def test_empty_batch_and_bad_args():
    sd = O.synth_state_dict(1)
    net = load_oanet(sd)
    out = cabi.filter_forward(torch.zeros((0, 1, 100, 6), device="cuda"), net.param_table(), net.cabi_cfg())
    assert out["logits"].shape == (2, 0, 100)
    with pytest.raises(AssertionError):
        net({"xs": torch.zeros(2, 100, 6)})
    with pytest.raises(cabi.LmpcrError):
        cabi.filter_forward(torch.zeros((1, 1, 100, 6), device="cuda"), net.param_table()[:-1], net.cabi_cfg())


@pytest.mark.parametrize("algo", GEMM_ALGOS)
@pytest.mark.parametrize("P,cin,cout,N,res", [(3, 128, 128, 1000, True), (2, 256, 128, 777, False), (2, 128, 500, 640, False), (1, 8, 128, 130, False)])
def test_conv1x1_layer(P, cin, cout, N, res, algo):
    """The fused layer alone (lmpcr_conv1x1) against fp64 numpy: relative error of a split-bf16 product is ~2^-16."""
    rng = np.random.default_rng(cin + N)
    x = rng.standard_normal((P, cin, N)).astype(np.float32)
    w = (rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, (P, cin)).astype(np.float32)
    sh = (0.3 * rng.standard_normal((P, cin))).astype(np.float32)
    r = rng.standard_normal((P, cout, N)).astype(np.float32) if res else None
    out = cabi.conv1x1(cu(x), cu(w), cu(b), cu(sc), cu(sh), cu(r) if res else None, gemm_algo=algo).cpu().numpy()
    h = np.maximum(x.astype(np.float64) * sc[:, :, None] + sh[:, :, None], 0)
    ref = np.einsum("oc,pcn->pon", w.astype(np.float64), h) + b[None, :, None] + (r if res else 0)
    assert np.abs(out - ref).max() < (2e-5 if algo == 0 else 1e-4) * max(1.0, np.abs(ref).max())
    plain = cabi.conv1x1(cu(x), cu(w), gemm_algo=algo).cpu().numpy()
    assert np.abs(plain - np.einsum("oc,pcn->pon", w.astype(np.float64), x.astype(np.float64))).max() < 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("N", [200, 2000, 20000])
def test_softmax_pool_modes_against_fp64(N):
    """diff_pool's weighted sum alone (oanet.py:107-109) through lmpcr_softmax_pool: normalised operand vs deferred normalisation.
    Gate 5e-5 relative to the largest output: split-bf16 products carry 2^-17 per operand, and reductions longer than 8192 points
    are re-summed in round-to-nearest segments (tcgen05 accumulates with round-toward-zero)."""
    rng = np.random.default_rng(N)
    x = (rng.standard_normal((2, 128, N)) * 2 + 0.5).astype(np.float32)
    E = (rng.standard_normal((2, 500, N)) * 3).astype(np.float32)
    e64 = E.astype(np.float64)
    S = np.exp(e64 - e64.max(2, keepdims=True))
    S /= S.sum(2, keepdims=True)
    ref = np.matmul(x.astype(np.float64), S.transpose(0, 2, 1))
    for mode in (0, 1):
        got = cabi.softmax_pool(cu(x), cu(E), mode).cpu().numpy()
        assert np.abs(got - ref).max() < 5e-5 * np.abs(ref).max(), (N, mode)
