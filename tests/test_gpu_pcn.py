"""GPU: the pair-resident PointCN stack kernel (csrc/pcn.cu) against an fp64 numpy restatement of lib/filtering/oanet.py:18-43,
alone (lmpcr_pointcn_stack) and inside lmpcr_filter_forward (groups of >= 64 pairs take it; LMPCR_PCN=0 switches it off)."""
import os

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from util import cabi, cu, load_oanet

pytestmark = pytest.mark.gpu


def _layer_params(rng, C=128):
    def bn():
        return [rng.uniform(0.5, 1.5, C), 0.3 * rng.standard_normal(C), 0.2 * rng.standard_normal(C), rng.uniform(0.5, 1.5, C)]
    w1, w2 = rng.standard_normal((C, C)) / np.sqrt(C), rng.standard_normal((C, C)) / np.sqrt(C)
    return [a.astype(np.float32) for a in bn() + [w1, 0.1 * rng.standard_normal(C)] + bn() + [w2, 0.1 * rng.standard_normal(C)]]


def _pointcn_ref(x, lp):
    """x + conv.7(relu(bn5(in(conv.3(relu(bn1(in(x))))))))   (oanet.py:24-43), fp64, InstanceNorm eps 1e-5, biased variance"""
    g1, b1, rm1, rv1, w1, c1, g2, b2, rm2, rv2, w2, c2 = [np.asarray(a, np.float64) for a in lp]

    def f(t, g, b, rm, rv):
        m, v = t.mean(2, keepdims=True), t.var(2, keepdims=True)
        y = (t - m) / np.sqrt(v + 1e-5)
        y = (y - rm[None, :, None]) / np.sqrt(rv[None, :, None] + 1e-5) * g[None, :, None] + b[None, :, None]
        return np.maximum(y, 0)

    y = np.einsum("oc,pcn->pon", w1, f(x, g1, b1, rm1, rv1)) + c1[None, :, None]
    return x + np.einsum("oc,pcn->pon", w2, f(y, g2, b2, rm2, rv2)) + c2[None, :, None]


@pytest.mark.parametrize("P,N,layers", [(3, 1000, 1), (2, 64, 3), (5, 1332, 3), (150, 96, 2), (1, 5000, 2)])
def test_pointcn_stack_against_fp64(P, N, layers):
    rng = np.random.default_rng(P * 1000 + N)
    x = (rng.standard_normal((P, 128, N)) * 2 + 0.5).astype(np.float32)
    lps = [_layer_params(rng) for _ in range(layers)]
    ref = x.astype(np.float64)
    for lp in lps:
        ref = _pointcn_ref(ref, lp)
    out, stats = cabi.pointcn_stack(cu(x), [[cu(a) for a in lp] for lp in lps], want_stats=True)
    out, stats = out.cpu().numpy(), stats.cpu().numpy()
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(out - ref).max() < 1e-4 * layers * scale, np.abs(out - ref).max()
    # fused statistics of the output: what the next layer's InstanceNorm needs
    assert np.abs(stats[..., 0] - ref.mean(2)).max() < 1e-4 * scale
    assert np.abs(stats[..., 1] / N - ref.var(2)).max() < 2e-4 * max(1.0, ref.var(2).max())
    # in place (the network runs l1_2's middle layers over the buffer they read)
    xin = cu(x)
    out2 = cabi.pointcn_stack(xin, [[cu(a) for a in lp] for lp in lps], out=xin)
    assert out2.data_ptr() == xin.data_ptr() and np.array_equal(out2.cpu().numpy(), out)


def test_network_takes_the_stack_kernel_and_agrees_with_the_layer_path():
    """74 pairs in one call (two whole waves of 37: the group is not cut): l1_1 and the middle of l1_2 run through pcn_stack_kernel; same logits as the per-layer GEMM path
    (LMPCR_PCN=0) within the tensor-path noise, and within the 5e-4 gate of the fp64 oracle."""
    sd = O.synth_state_dict(3)
    xs, _, _ = O.synth_xs(74, 500, seed=3)
    net = load_oanet(sd, gemm_algo=1)
    x = cu(xs)
    n0 = cabi.launch_count_named("pcn_stack_kernel")
    out = net({"xs": x})
    assert cabi.launch_count_named("pcn_stack_kernel") - n0 == 4          # 2 blocks x (l1_1 stack + l1_2 tail with the fused head)
    os.environ["LMPCR_PCN"] = "0"
    try:
        n1 = cabi.launch_count_named("pcn_stack_kernel")
        ref = net({"xs": x})
        assert cabi.launch_count_named("pcn_stack_kernel") == n1
    finally:
        del os.environ["LMPCR_PCN"]
    # two evaluations of the same fp32 network against the fp64 oracle, every pair and both blocks: the stack kernel is as accurate
    # as the per-layer path (the maximum over 74 pairs sits above the 5e-4 gate the 2..5-pair tests use: the second block amplifies
    # the first block's pose through its residual input)
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    err = {name: [np.abs(res["logits"][it].cpu().numpy() - o64["logits"][it]).max() for it in range(2)] for name, res in (("pcn", out), ("layers", ref))}
    print("max |logit - fp64| per block:", err)
    for it in range(2):
        assert err["pcn"][it] < max(5e-4, 1.5 * err["layers"][it]), err
    assert err["pcn"][0] < 5e-4 and err["layers"][0] < 5e-4, err
    assert O.chordal_angle(out["rot_est"][-1].cpu().numpy(), o64["rot_est"][-1]).max() < 1e-3
    assert (out["logits"][0] - ref["logits"][0]).abs().max().item() < 5e-4
