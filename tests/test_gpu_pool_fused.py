"""GPU: the fused diff_pool kernel (csrc/pool_fused.cu: embedding conv + softmax over the points + weighted sum in one launch)
against an fp64 numpy restatement of lib/filtering/oanet.py:96-110, alone (lmpcr_diff_pool_fused) and inside lmpcr_filter_forward
(groups of >= 64 pairs take it; LMPCR_POOL_FUSED=0 switches it off)."""
import os

import numpy as np
import pytest

from oracle import lmpcr_oracle as O
from util import cabi, cu, load_oanet

pytestmark = pytest.mark.gpu


def _diff_pool_ref(x, sc, sh, w, b):
    """out[p,c,k] = sum_n x[p,c,n] softmax_n(w relu(x*sc+sh) + b)[k,n]  in fp64 (the bias is kept here: it must cancel)"""
    x64 = x.astype(np.float64)
    h = np.maximum(x64 * sc.astype(np.float64)[:, :, None] + sh.astype(np.float64)[:, :, None], 0)
    E = np.einsum("kc,pcn->pkn", w.astype(np.float64), h) + b.astype(np.float64)[None, :, None]
    S = np.exp(E - E.max(2, keepdims=True))
    S /= S.sum(2, keepdims=True)
    return np.matmul(x64, S.transpose(0, 2, 1))


@pytest.mark.parametrize("mode", [0, 1], ids=["single_pass", "two_pass"])
@pytest.mark.parametrize("P,N,K", [(3, 200, 500), (2, 2000, 500), (2, 5000, 500), (75, 264, 500), (3, 64, 16), (2, 1000, 130), (2, 8192, 300), (1, 28, 500)])
def test_diff_pool_fused_against_fp64(P, N, K, mode):
    rng = np.random.default_rng(P * 100000 + N * 10 + K)
    x = (rng.standard_normal((P, 128, N)) * 2 + 0.5).astype(np.float32)
    sc = rng.uniform(0.3, 1.2, (P, 128)).astype(np.float32)
    sh = (0.5 * rng.standard_normal((P, 128))).astype(np.float32)
    w = (rng.standard_normal((K, 128)) * 3 / np.sqrt(128)).astype(np.float32)          # logits of a few units: a peaked softmax
    b = rng.standard_normal(K).astype(np.float32)
    ref = _diff_pool_ref(x, sc, sh, w, b)
    got = cabi.diff_pool_fused(cu(x), cu(sc), cu(sh), cu(w), mode).cpu().numpy()
    assert got.shape == (P, 128, K)
    err = np.abs(got - ref).max()
    # 5e-5 relative to the largest output, the gate of lmpcr_softmax_pool (split-bf16 products, fp32 accumulation)
    assert err < 5e-5 * np.abs(ref).max(), (err, np.abs(ref).max())


@pytest.mark.parametrize("mode", [0, 1], ids=["single_pass", "two_pass"])
def test_diff_pool_fused_extreme_logits(mode):
    """Rows whose maximum is far from zero (both signs) and one dominating point in the first tile: exp(E - shift) stays in range."""
    rng = np.random.default_rng(7)
    P, N, K = 2, 1000, 500
    x = (rng.standard_normal((P, 128, N))).astype(np.float32)
    x[:, :, 17] *= 6.0                                        # one point with large features
    sc = np.ones((P, 128), np.float32)
    sh = np.zeros((P, 128), np.float32)
    w = (rng.standard_normal((K, 128)) * 8 / np.sqrt(128)).astype(np.float32)
    ref = _diff_pool_ref(x, sc, sh, w, np.zeros(K, np.float32))
    got = cabi.diff_pool_fused(cu(x), cu(sc), cu(sh), cu(w), mode).cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()


def test_diff_pool_fused_single_pass_overflow_takes_the_fallback():
    """A point far behind the first tile whose logits exceed every first-tile logit by hundreds: under the first tile's shift
    exp() overflows, the row sums reveal it, and the second launch redoes those (pair, cluster block) items with row maxima over
    all points.  Pair 1 is tame and must come out of the single pass untouched by the fallback (bit-identical to a tame-only call)."""
    rng = np.random.default_rng(11)
    P, N, K = 2, 1500, 500
    x = rng.standard_normal((P, 128, N)).astype(np.float32)
    x[0, :, 900] *= 30.0
    sc = np.ones((P, 128), np.float32)
    sh = np.zeros((P, 128), np.float32)
    w = (rng.standard_normal((K, 128)) * 8 / np.sqrt(128)).astype(np.float32)
    ref = _diff_pool_ref(x, sc, sh, w, np.zeros(K, np.float32))
    E0 = w.astype(np.float64) @ np.maximum(x[0].astype(np.float64), 0)
    assert (E0[:, 900] - E0[:, :64].max(1)).max() > 100, "the case must overflow the single pass"
    got = cabi.diff_pool_fused(cu(x), cu(sc), cu(sh), cu(w), 0).cpu().numpy()
    two = cabi.diff_pool_fused(cu(x), cu(sc), cu(sh), cu(w), 1).cpu().numpy()
    assert np.isfinite(got).all()
    # logits of several hundred carry an absolute fp32 error of ~1e-2, which rows where the outlier ties with ordinary points turn into
    # a relative output error of that size: the gate against fp64 is loose here, the one against the two-pass mode (same arithmetic,
    # same maxima for the redone items) is tight
    assert np.abs(got - ref).max() < 1e-3 * np.abs(ref).max()
    assert np.array_equal(got[0], two[0]), "flagged items are redone by exactly the two-pass code"
    assert np.abs(got[1] - two[1]).max() < 2e-5 * np.abs(ref[1]).max()
    tame = cabi.diff_pool_fused(cu(x[1:]), cu(sc[1:]), cu(sh[1:]), cu(w), 0).cpu().numpy()
    assert np.array_equal(tame[0], got[1])


@pytest.mark.parametrize("P,N,K", [(3, 200, 500), (2, 5000, 500), (75, 264, 500), (3, 64, 16), (2, 1000, 130), (1, 28, 500)])
def test_embed_fused_against_fp64(P, N, K):
    """The embedding conv of diff_unpool on the pair-resident kernel (lmpcr_embed_fused): logits and their column maxima (the shift of
    the softmax over the clusters) against fp64; rows / columns past the matrix edge must stay untouched (canary)."""
    rng = np.random.default_rng(P * 100000 + N * 10 + K + 1)
    x = (rng.standard_normal((P, 128, N)) * 2 + 0.5).astype(np.float32)
    sc = rng.uniform(0.3, 1.2, (P, 128)).astype(np.float32)
    sh = (0.5 * rng.standard_normal((P, 128))).astype(np.float32)
    w = (rng.standard_normal((K, 128)) * 3 / np.sqrt(128)).astype(np.float32)
    b = rng.standard_normal(K).astype(np.float32)
    h = np.maximum(x.astype(np.float64) * sc.astype(np.float64)[:, :, None] + sh.astype(np.float64)[:, :, None], 0)
    ref = np.einsum("kc,pcn->pkn", w.astype(np.float64), h) + b.astype(np.float64)[None, :, None]
    E, cm = cabi.embed_fused(cu(x), cu(sc), cu(sh), cu(w), cu(b), want_colmax=True)
    E, cm = E.cpu().numpy(), cm.cpu().numpy()
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(E - ref).max() < 2e-5 * scale, np.abs(E - ref).max()
    assert np.abs(cm / np.log2(np.e) - ref.max(1)).max() < 2e-5 * scale
    # without a bias, without the maxima
    E0 = cabi.embed_fused(cu(x), cu(sc), cu(sh), cu(w)).cpu().numpy()
    assert np.abs(E0 - (ref - b.astype(np.float64)[None, :, None])).max() < 2e-5 * scale


def test_network_takes_the_fused_pool_and_agrees_with_the_gemm_path():
    """74 pairs in one call: both blocks run diff_pool through pool_fused_kernel; same logits as the embedding-GEMM + pooling-GEMM
    path (LMPCR_POOL_FUSED=0) within the tensor-path noise, and the first block within the 5e-4 gate of the fp64 oracle."""
    sd = O.synth_state_dict(5)
    xs, _, _ = O.synth_xs(74, 500, seed=5)
    net = load_oanet(sd, gemm_algo=1)
    x = cu(xs)
    n0 = cabi.launch_count_named("pool_fused_kernel")
    out = net({"xs": x})
    assert cabi.launch_count_named("pool_fused_kernel") - n0 == 4          # 2 blocks x (single pass + fallback launch)
    n0e = cabi.launch_count_named("embed_fused_kernel")
    os.environ["LMPCR_POOL_FUSED"] = "0"
    try:
        n1 = cabi.launch_count_named("pool_fused_kernel")
        ref = net({"xs": x})
        assert cabi.launch_count_named("pool_fused_kernel") == n1
        assert cabi.launch_count_named("embed_fused_kernel") - n0e == 2       # the `up` embedding conv of both blocks, independent of the pool switch
        os.environ["LMPCR_EMBED_FUSED"] = "0"
        n2 = cabi.launch_count_named("embed_fused_kernel")
        ref2 = net({"xs": x})
        assert cabi.launch_count_named("embed_fused_kernel") == n2
        assert (ref2["logits"][0] - ref["logits"][0]).abs().max().item() < 5e-4
        # diff_unpool's product: pair-resident kernel (default, one launch per block) vs the generic GEMM
        os.environ["LMPCR_UNPOOL_FUSED"] = "0"
        n3 = cabi.launch_count_named("unpool_fused_kernel")
        assert n3 >= 6                                                        # the three forwards above took it in both blocks
        ref3 = net({"xs": x})
        assert cabi.launch_count_named("unpool_fused_kernel") == n3
        assert (ref3["logits"][0] - ref2["logits"][0]).abs().max().item() < 5e-4
    finally:
        del os.environ["LMPCR_POOL_FUSED"]
        os.environ.pop("LMPCR_EMBED_FUSED", None)
        os.environ.pop("LMPCR_UNPOOL_FUSED", None)
    o64 = O.oanet_forward(xs, sd, dtype=np.float64)
    err = {name: [np.abs(res["logits"][it].cpu().numpy() - o64["logits"][it]).max() for it in range(2)] for name, res in (("fused", out), ("gemm", ref))}
    print("max |logit - fp64| per block:", err)
    for it in range(2):
        assert err["fused"][it] < max(5e-4, 1.5 * err["gemm"][it]), err
    assert err["fused"][0] < 5e-4 and err["gemm"][0] < 5e-4, err
    # rotations: the worst of the 74 synthetic pairs is ill-conditioned (1e-4 on the logits moves it by ~1e-3 rad on either path), so the
    # gate on the worst pair is 2e-3 for both paths (measured 1.1e-3 / 6.4e-4); the median pair is two orders of magnitude tighter
    ang = {name: O.chordal_angle(res["rot_est"][-1].cpu().numpy(), o64["rot_est"][-1]) for name, res in (("fused", out), ("gemm", ref))}
    print("chordal angle vs fp64: max / median", {k: (float(v.max()), float(np.median(v))) for k, v in ang.items()})
    assert ang["fused"].max() < 2e-3 and ang["gemm"].max() < 2e-3
    assert np.median(ang["fused"]) < 5e-5 and (ang["fused"] > 2e-4).mean() < 0.1
    assert (out["logits"][0] - ref["logits"][0]).abs().max().item() < 5e-4


@pytest.mark.parametrize("P,N", [(3, 200), (2, 5000), (75, 264), (1, 28), (2, 1332)])
def test_conv_wide_against_fp64(P, N):
    """The two 256 -> 128 convolutions of l1_2's first PointCN in one launch (lmpcr_conv_wide, csrc/conv_wide.cu): shot_cut on the raw
    input, conv.3 behind the folded norm affine + ReLU; outputs and the fused row statistics against fp64."""
    rng = np.random.default_rng(P * 10000 + N)
    x = (rng.standard_normal((P, 256, N)) * 2 + 0.5).astype(np.float32)
    w0 = (rng.standard_normal((128, 256)) / 16).astype(np.float32)
    w1 = (rng.standard_normal((128, 256)) / 16).astype(np.float32)
    b0, b1 = rng.standard_normal(128).astype(np.float32), rng.standard_normal(128).astype(np.float32)
    sc = rng.uniform(0.3, 1.2, (P, 256)).astype(np.float32)
    sh = (0.5 * rng.standard_normal((P, 256))).astype(np.float32)
    x64 = x.astype(np.float64)
    ref0 = np.einsum("oc,pcn->pon", w0.astype(np.float64), x64) + b0.astype(np.float64)[None, :, None]
    h = np.maximum(x64 * sc.astype(np.float64)[:, :, None] + sh.astype(np.float64)[:, :, None], 0)
    ref1 = np.einsum("oc,pcn->pon", w1.astype(np.float64), h) + b1.astype(np.float64)[None, :, None]
    outs, stats = cabi.conv_wide(cu(x), [dict(weight=cu(w0), bias=cu(b0)), dict(weight=cu(w1), bias=cu(b1), scale=cu(sc), shift=cu(sh))], want_stats=True)
    for o, st, ref in zip(outs, stats, (ref0, ref1)):
        o, st = o.cpu().numpy(), st.cpu().numpy()
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(o - ref).max() < 2e-5 * scale, np.abs(o - ref).max()
        assert np.abs(st[..., 0] - ref.mean(2)).max() < 1e-4 * scale
        assert np.abs(st[..., 1] / N - ref.var(2)).max() < 2e-4 * max(1.0, ref.var(2).max())
    # a single convolution, no bias
    one = cabi.conv_wide(cu(x), [dict(weight=cu(w1), scale=cu(sc), shift=cu(sh))])[0].cpu().numpy()
    assert np.abs(one - (ref1 - b1.astype(np.float64)[None, :, None])).max() < 2e-5 * max(1.0, np.abs(ref1).max())


@pytest.mark.parametrize("P,K,N", [(2, 500, 2000), (3, 500, 644), (2, 512, 36), (5, 300, 1000), (2, 388, 4), (150, 500, 256), (300, 500, 128)])
def test_unpool_fused_against_fp64(P, K, N):
    """diff_unpool's product on the pair-resident kernel (lmpcr_softmax_unpool mode 2): x_down on chip, E by TMA, column sums in the
    producers.  Ragged last tiles (N % 64 != 0, N < 32), partial last cluster chunk, several CTAs per pair (few pairs) and several pairs
    per CTA (P > 148)."""
    rng = np.random.default_rng(P * 1000 + K + N)
    xd = (rng.standard_normal((P, 128, K)) * 2 + 0.5).astype(np.float32)
    E = (rng.standard_normal((P, K, N)) * 3).astype(np.float32)
    E[0, :, 0] += 40.0            # a column far from the others: the shift by the column maximum keeps it finite
    E[-1, K - 1, N - 1] = 25.0    # the last cluster row / last point dominate their column
    e64 = E.astype(np.float64)
    S = np.exp(e64 - e64.max(1, keepdims=True))
    S /= S.sum(1, keepdims=True)
    ref = np.matmul(xd.astype(np.float64), S)
    got = cabi.softmax_unpool(cu(xd), cu(E), 2).cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 5e-5 * np.abs(ref).max()
    # same arithmetic as the generic deferred path up to the summation order
    gen = cabi.softmax_unpool(cu(xd), cu(E), 1).cpu().numpy()
    assert np.abs(got - gen).max() < 2e-5 * np.abs(ref).max()
    # bit-identical from run to run (fixed-order column sums)
    again = cabi.softmax_unpool(cu(xd), cu(E), 2).cpu().numpy()
    assert np.array_equal(got, again)
