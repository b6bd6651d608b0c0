"""GPU: the pair-resident OAFilter stack (csrc/oaf.cu: conv1 -> BatchNorm over the clusters -> cluster-mixing conv2 -> conv3 + shortcut of
all OAFilter layers of a block in one launch) against the fp64 oracle restatement of lib/filtering/oanet.py:56-93 (oracle._oafilter,
pinned by the network goldens), alone (lmpcr_oafilter_stack) and inside lmpcr_filter_forward (groups of >= 64 pairs take it;
LMPCR_OAF=0 switches it off)."""
import os

import numpy as np
import pytest

from oracle import lmpcr_oracle as O
from util import cabi, cu, load_oanet

pytestmark = pytest.mark.gpu

_NAMES = [("conv1.1", ("weight", "bias", "running_mean", "running_var")), ("conv1.3", ("weight", "bias")),
          ("conv2.0", ("weight", "bias", "running_mean", "running_var")), ("conv2.2", ("weight", "bias")),
          ("conv3.2", ("weight", "bias", "running_mean", "running_var")), ("conv3.4", ("weight", "bias"))]


def _layer_sd(rng, K, scale_w2=1.0):
    """One OAFilter's tensors under the reference's names, with non-trivial BatchNorm statistics."""
    sd = {}
    for mod, ch in (("conv1.1", 128), ("conv2.0", K), ("conv3.2", 128)):
        sd[mod + ".weight"] = rng.uniform(0.5, 1.5, ch).astype(np.float32)
        sd[mod + ".bias"] = (0.3 * rng.standard_normal(ch)).astype(np.float32)
        sd[mod + ".running_mean"] = (0.2 * rng.standard_normal(ch)).astype(np.float32)
        sd[mod + ".running_var"] = rng.uniform(0.5, 1.5, ch).astype(np.float32)
    for mod, ch in (("conv1.3", 128), ("conv3.4", 128)):
        sd[mod + ".weight"] = (rng.standard_normal((ch, ch, 1, 1)) / np.sqrt(ch)).astype(np.float32)
        sd[mod + ".bias"] = (0.1 * rng.standard_normal(ch)).astype(np.float32)
    sd["conv2.2.weight"] = (scale_w2 * rng.standard_normal((K, K, 1, 1)) / np.sqrt(K)).astype(np.float32)
    sd["conv2.2.bias"] = (0.1 * rng.standard_normal(K)).astype(np.float32)
    return sd


def _ref(x, layers):
    out = x.astype(np.float64)
    for sd in layers:
        out = O._oafilter(out, O._SD(sd, "", np.float64))
    return out


@pytest.mark.parametrize("P,K,n_layers", [(3, 500, 1), (2, 500, 2), (5, 500, 3), (150, 500, 3), (2, 512, 2), (2, 484, 3), (1, 500, 4)])
def test_oafilter_stack_against_fp64(P, K, n_layers):
    rng = np.random.default_rng(P * 10000 + K * 10 + n_layers)
    x = (rng.standard_normal((P, 128, K)) * 1.5 + 0.3).astype(np.float32)
    layers = [_layer_sd(rng, K) for _ in range(n_layers)]
    ref = _ref(x, layers)
    params = [[cu(sd[m + "." + n]) for m, ns in _NAMES for n in ns] for sd in layers]
    n0 = cabi.launch_count_named("oaf_stack_kernel")
    got = cabi.oafilter_stack(cu(x), params).cpu().numpy()
    assert cabi.launch_count_named("oaf_stack_kernel") - n0 == 1
    assert got.shape == ref.shape and np.isfinite(got).all()
    err = np.abs(got - ref).max()
    # split-bf16 products with fp32 accumulation and fp32 statistics: 5e-5 of the largest value per layer
    assert err < 5e-5 * n_layers * np.abs(ref).max(), (err, np.abs(ref).max())


def test_oafilter_stack_is_deterministic_and_per_pair():
    """Two runs give the same bits, and a pair's result does not depend on which CTA or position in the batch it had."""
    rng = np.random.default_rng(77)
    K = 500
    x = (rng.standard_normal((160, 128, K)) + 0.1).astype(np.float32)
    layers = [_layer_sd(rng, K) for _ in range(3)]
    params = [[cu(sd[m + "." + n]) for m, ns in _NAMES for n in ns] for sd in layers]
    a = cabi.oafilter_stack(cu(x), params)
    b = cabi.oafilter_stack(cu(x), params)
    assert (a == b).all().item()
    c = cabi.oafilter_stack(cu(x[149:152]), params)
    assert (c == a[149:152]).all().item()


def test_network_takes_the_oafilter_stack_and_agrees_with_the_gemm_path():
    """74 pairs in one call: both blocks run their OAFilter stage through oaf_stack_kernel (one launch per block); same logits as the
    nine-GEMM path (LMPCR_OAF=0) within the tensor-path noise, and within the gate of the fp64 oracle."""
    sd = O.synth_state_dict(9)
    xs, _, _ = O.synth_xs(74, 500, seed=9)
    net = load_oanet(sd, gemm_algo=1)
    x = cu(xs)
    n0 = cabi.launch_count_named("oaf_stack_kernel")
    out = net({"xs": x})
    assert cabi.launch_count_named("oaf_stack_kernel") - n0 == 2
    os.environ["LMPCR_OAF"] = "0"
    try:
        n1 = cabi.launch_count_named("oaf_stack_kernel")
        ref = net({"xs": x})
        assert cabi.launch_count_named("oaf_stack_kernel") == n1
    finally:
        del os.environ["LMPCR_OAF"]
    for it in range(2):
        assert (out["logits"][it] - ref["logits"][it]).abs().max().item() < 5e-4
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    err = {name: [np.abs(res["logits"][it].cpu().numpy() - o64["logits"][it]).max() for it in range(2)] for name, res in (("fused", out), ("gemm", ref))}
    print("max |logit - fp64| per block:", err)
    for it in range(2):
        assert err["fused"][it] < max(5e-4, 1.5 * err["gemm"][it]), err
    assert err["fused"][0] < 5e-4, err
