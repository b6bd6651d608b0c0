"""GPU: the drop-in boundary runs the measured path: a reference-shaped config without any extra key takes the tcgen05
GEMMs, weights are packed once (no split kernels from the second call on), workspaces are reused, and the ADVICE edge cases hold."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from util import cabi, cu, load_oanet

pytestmark = pytest.mark.gpu

# configs/pairwise_registration/eval/RegBlock.yaml, the keys OANet reads (oanet.py:203-215) -- nothing of ours added
STOCK_CFG = {"method": {"task": "pairwise", "descriptor_module": None, "filter_module": "oanet"},
             "misc": {"run_mode": "train", "net_depth": 12, "clusters": 500, "iter_num": 1, "net_channel": 128, "use_gpu": True,
                      "normalize_weights": True, "inlier_weight_threshold": 0.5},
             "train": {"samp_type": "rand", "corr_type": "soft", "st_grad_flag": True},
             "data": {"max_num_points": 5000, "use_mutuals": 0}}


def test_stock_config_runs_tensor_path_and_packs_weights_once():
    oanet = importlib.import_module("3d_multiview_reg_b200.lib.filtering.oanet")
    sd = O.synth_state_dict(11)
    net = oanet.OANet(STOCK_CFG).eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    net = net.cuda()
    xs, _, _ = O.synth_xs(3, 640, seed=11)
    c = cabi.launch_count_named
    t0, s0, p0, g0 = c("tcgemm_kernel"), c("split_weights_kernel"), c("pcn_pack_weights_kernel"), c("gemm_fused_kernel")
    out1 = net({"xs": torch.from_numpy(xs)})                      # CPU input, moved inside like oanet.py:234
    assert c("tcgemm_kernel") > t0 and c("gemm_fused_kernel") == g0, "the stock config must take the tcgen05 GEMMs"
    s1, p1 = c("split_weights_kernel"), c("pcn_pack_weights_kernel")
    assert s1 > s0                                                 # packed at the first call ...
    ws1 = net._workspace
    out2 = net({"xs": torch.from_numpy(xs)})
    # ... and from then on only the per-pair activation operands (x_down for diff_unpool) are split: 1 launch per block
    assert c("split_weights_kernel") - s1 == 2 and c("pcn_pack_weights_kernel") == p1
    assert net._workspace is ws1, "the scratch tensor is reused across calls"
    assert torch.equal(out1["logits"][-1], out2["logits"][-1])
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    assert np.abs(out1["logits"][-1].cpu().numpy() - o64["logits"][-1]).max() < 5e-4
    # a changed weight repacks
    with torch.no_grad():
        dict(net.named_parameters())["reg_init.l1_1.0.conv.3.weight"].mul_(1.0)
    net({"xs": torch.from_numpy(xs)})
    assert c("pcn_pack_weights_kernel") > p1


def test_scene_path_forces_eval_batchnorm_and_keeps_running_stats():
    """A forgotten .eval(): the scene path must not switch to batch statistics (results would depend on pair_chunk) nor touch
    running_mean / running_var."""
    scene = importlib.import_module("3d_multiview_reg_b200.scene")
    S, n = 4, 512
    feats, xyz, _ = O.synth_scene(S, n, seed=21)
    net = load_oanet(O.synth_state_dict(21))
    f, x = cu(feats), cu(xyz)
    rec_eval = scene.SceneRegistrar(net, pair_chunk=2).register_scene(f, x)
    before = {k: v.clone() for k, v in net.named_buffers()}
    net.train()
    rec_a = scene.SceneRegistrar(net, pair_chunk=2).register_scene(f, x)
    rec_b = scene.SceneRegistrar(net, pair_chunk=5).register_scene(f, x)
    assert torch.equal(rec_a, rec_eval) and torch.equal(rec_b, rec_eval)
    for k, v in net.named_buffers():
        assert torch.equal(v, before[k]), k


@pytest.mark.parametrize("algo", [cabi.NN_EXACT_SIMT, cabi.NN_TENSOR])
def test_nan_feature_row_gives_a_valid_index(algo):
    """ADVICE: a query row of NaN used to return 0x7fffffff on the tensor path and the gather kernels read out of bounds."""
    rng = np.random.default_rng(3)
    f = rng.standard_normal((2, 300, 32)).astype(np.float32)
    f /= np.linalg.norm(f, axis=2, keepdims=True)
    f[0, 17] = np.nan
    f[0, 40, 3] = np.inf
    jobs = torch.tensor([[0, 1], [1, 0]], dtype=torch.int32).cuda()
    idx = cabi.nn_argmin(cu(f), cu(f), jobs, algo=algo)
    assert int(idx.min()) >= 0 and int(idx.max()) < 300
    xyz = cu(rng.uniform(0, 1, (2, 300, 3)).astype(np.float32))
    _, xs = cabi.mutual_xs(xyz, torch.tensor([[0, 1]], dtype=torch.int32).cuda(), idx[:1], idx[1:])
    assert torch.isfinite(xs).all()
    # a corrupt index handed to the gather / mutual kernels is clamped, not dereferenced
    bad = idx.clone(); bad[0, 5] = 2 ** 31 - 1; bad[1, 7] = -3
    g = cabi.gather_xyz(xyz, jobs, bad)
    assert torch.isfinite(g).all()
    m, xs2 = cabi.mutual_xs(xyz, torch.tensor([[0, 1]], dtype=torch.int32).cuda(), bad[:1], bad[1:])
    assert torch.isfinite(xs2).all()


def test_two_devices_in_one_process():
    """ADVICE: the opt-in to large dynamic shared memory is per device; a process that drives a second GPU must set it there too."""
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 128, 512)).astype(np.float32)
    w = (rng.standard_normal((128, 128)) / 11).astype(np.float32)
    outs = []
    for d in (0, 1):
        with torch.cuda.device(d):
            outs.append(cabi.conv1x1(torch.from_numpy(x).cuda(d), torch.from_numpy(w).cuda(d), gemm_algo=1).cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("N", [640, 5000])
def test_softmax_unpool_against_fp64(N):
    """diff_unpool's weighted sum alone (oanet.py:126-128) through lmpcr_softmax_unpool: softmax over the CLUSTER axis."""
    rng = np.random.default_rng(N)
    xd = (rng.standard_normal((2, 128, 500)) * 2 + 0.5).astype(np.float32)
    E = (rng.standard_normal((2, 500, N)) * 3).astype(np.float32)
    e64 = E.astype(np.float64)
    S = np.exp(e64 - e64.max(1, keepdims=True))
    S /= S.sum(1, keepdims=True)
    ref = np.matmul(xd.astype(np.float64), S)
    for mode in (0, 1):
        got = cabi.softmax_unpool(cu(xd), cu(E), mode).cpu().numpy()
        assert np.abs(got - ref).max() < 5e-5 * np.abs(ref).max(), (N, mode)
    # a softmax over the wrong axis would be far off
    Sw = np.exp(e64 - e64.max(2, keepdims=True)); Sw /= Sw.sum(2, keepdims=True)
    assert np.abs(np.matmul(xd.astype(np.float64), Sw) - ref).max() > 1e-2 * np.abs(ref).max()


def test_grouping_invariance_at_bench_size():
    """The bench workload's shape (5000 points, tcgen05 GEMMs): records of a group of pairs equal the records of the same pairs
    run one by one -- per-pair arithmetic does not depend on the grouping, with or without the pair-resident kernels."""
    sd = O.synth_state_dict(7)
    xs, _, _ = O.synth_xs(74, 5000, seed=7)
    net = load_oanet(sd, gemm_algo=1)
    cfg = net.cabi_cfg()
    cfg.guard_mode = cabi.GUARD_PAIR
    x = cu(xs)
    params, packed = net.param_table(), net.packed_weights()
    full = cabi.filter_forward(x, params, cfg, want_latent=False, packed=packed)
    for p in (0, 36, 73):
        one = cabi.filter_forward(x[p:p + 1].contiguous(), params, cfg, want_latent=False, packed=packed)
        # the group took pcn_stack_kernel and pool_fused_kernel, the single pair the per-layer GEMMs: equal within the two evaluations' fp32 noise
        assert (one["logits"][0][0] - full["logits"][0][p]).abs().max().item() < 5e-4
        assert O.chordal_angle(one["R"][-1].cpu().numpy(), full["R"][-1][p:p + 1].cpu().numpy()).max() < 1e-3
    os.environ["LMPCR_PCN"] = "0"
    os.environ["LMPCR_POOL_FUSED"] = "0"
    os.environ["LMPCR_EMBED_FUSED"] = "0"
    os.environ["LMPCR_CONV_WIDE"] = "0"
    os.environ["LMPCR_UNPOOL_FUSED"] = "0"
    os.environ["LMPCR_OAF"] = "0"
    try:
        full_l = cabi.filter_forward(x, params, cfg, want_latent=False, packed=packed)
        for p in (0, 36, 73):
            one = cabi.filter_forward(x[p:p + 1].contiguous(), params, cfg, want_latent=False, packed=packed)
            for k in ("logits", "scores", "R", "t"):
                assert torch.equal(one[k][:, 0], full_l[k][:, p]), (k, p)      # same kernels: bit-identical
    finally:
        del os.environ["LMPCR_PCN"]
        del os.environ["LMPCR_POOL_FUSED"]
        del os.environ["LMPCR_EMBED_FUSED"]
        del os.environ["LMPCR_CONV_WIDE"]
        del os.environ["LMPCR_UNPOOL_FUSED"]
        del os.environ["LMPCR_OAF"]
    # the latent-feature variant of the pair-resident tail (tiles stored) gives the same logits as the on-chip variant
    lat = cabi.filter_forward(x, params, cfg, want_latent=True, packed=packed)
    assert torch.equal(lat["logits"], full["logits"]) and torch.equal(lat["R"], full["R"])
