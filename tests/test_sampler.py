"""Keypoint sampler (lib/layers.py:90-154, 'rand').  CPU: the oracle restatement reproduces the reference's own output under the
same numpy seed (golden from the unmodified reference, tests/golden/make_sampler_golden.py).  GPU: the mirror's rng='numpy' mode is
index-for-index the reference; the device stream (Philox keys + segmented sort) satisfies the sampling contract, is reproducible
per seed and uniform (chi-square on inclusion and first-position counts)."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O

CASES = [("norepl", [700, 512, 901], 256, 32, 5, 1234), ("repl", [300, 100, 250], 128, 16, 6, 99)]


def _inputs(pts, dim, seed):
    g = np.random.default_rng(seed)
    total = int(sum(pts))
    return g.standard_normal((total, 3)).astype(np.float32), g.standard_normal((total, dim)).astype(np.float32)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_golden(golden_dir, case):
    name, pts, m, dim, seed, npseed = case
    G = np.load(os.path.join(golden_dir, "sampler_golden.npz"))
    C, F = _inputs(pts, dim, seed)
    np.random.seed(npseed)
    idx, sc, sf = O.sampler_rand(C, F, pts, m)
    assert np.array_equal(sc, G[name + "_C"]) and np.array_equal(sf, G[name + "_F"])
    assert O.sampler_contract(idx, pts, m) == []


def test_contract_checker_rejects_bad_samples():
    pts, m = [10, 12], 4
    assert O.sampler_contract(np.array([[0, 1, 2, 3], [10, 11, 12, 21]]), pts, m) == []
    assert O.sampler_contract(np.array([[0, 1, 2, 10], [10, 11, 12, 21]]), pts, m)          # leaves its cloud
    assert O.sampler_contract(np.array([[0, 1, 2, 2], [10, 11, 12, 21]]), pts, m)           # repeat without replacement
    assert O.sampler_contract(np.array([[0, 1, 2]]), pts, m)                                # shape


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_numpy_stream_mode_is_the_reference(golden_dir, case):
    layers = importlib.import_module("3d_multiview_reg_b200.lib.layers")
    name, pts, m, dim, seed, npseed = case
    G = np.load(os.path.join(golden_dir, "sampler_golden.npz"))
    C, F = _inputs(pts, dim, seed)
    s = layers.Sampler("rand", m, rng="numpy")
    np.random.seed(npseed)
    sc, sf = s(torch.from_numpy(C).cuda(), torch.from_numpy(F).cuda(), torch.tensor(pts))
    assert np.array_equal(sc.cpu().numpy(), G[name + "_C"]) and np.array_equal(sf.cpu().numpy(), G[name + "_F"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_device_sampler_contract_and_gather(case):
    from util import cabi
    name, pts, m, dim, seed, _ = case
    C, F = _inputs(pts, dim, seed)
    Cd, Fd = torch.from_numpy(C).cuda(), torch.from_numpy(F).cuda()
    replace = min(pts) < m
    idx, sc, sf = cabi.sample_keypoints(Cd, Fd, pts, m, replace, seed=77)
    idx_h = idx.cpu().numpy()
    assert O.sampler_contract(idx_h, pts, m) == []
    assert np.array_equal(sc.cpu().numpy(), C[idx_h]) and np.array_equal(sf.cpu().numpy(), F[idx_h])
    idx2, _, _ = cabi.sample_keypoints(Cd, Fd, pts, m, replace, seed=77)
    assert torch.equal(idx, idx2), "same seed, same sample"
    idx3, _, _ = cabi.sample_keypoints(Cd, Fd, pts, m, replace, seed=78)
    assert not torch.equal(idx, idx3)
    if not replace:          # not the identity order and not sorted: an ordered random subset
        assert not np.array_equal(np.sort(idx_h[0]), idx_h[0])


@pytest.mark.gpu
def test_device_sampler_module_and_errors():
    from util import cabi
    layers = importlib.import_module("3d_multiview_reg_b200.lib.layers")
    pts = [700, 512, 901]
    C, F = _inputs(pts, 32, 5)
    s = layers.Sampler("rand", 256)
    torch.manual_seed(3)
    a = s(torch.from_numpy(C).cuda(), torch.from_numpy(F).cuda(), torch.tensor(pts))
    torch.manual_seed(3)
    b = s(torch.from_numpy(C).cuda(), torch.from_numpy(F).cuda(), torch.tensor(pts))
    assert a[0].shape == (3, 256, 3) and a[1].shape == (3, 256, 32)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), "torch.manual_seed makes the device sample reproducible"
    with pytest.raises(cabi.LmpcrError):          # numpy: "Cannot take a larger sample than population when 'replace=False'"
        cabi.sample_keypoints(torch.from_numpy(C).cuda(), torch.from_numpy(F).cuda(), pts, 600, False, seed=1)
    with pytest.raises(NotImplementedError):
        layers.Sampler("fps", 256)(torch.from_numpy(C).cuda(), torch.from_numpy(F).cuda(), torch.tensor(pts))


@pytest.mark.gpu
@pytest.mark.parametrize("replace", [False, True])
def test_device_sampler_is_uniform(replace):
    """4000 clouds of 64 points, 16 samples each, one call: every point must be included ~1000 times and be first ~62.5 times."""
    from util import cabi
    b, n, m = 4000, 64, 16
    pts = [n] * b
    C = torch.zeros((b * n, 3), device="cuda")
    F = torch.zeros((b * n, 1), device="cuda")
    idx, _, _ = cabi.sample_keypoints(C, F, pts, m, replace, seed=2026)
    local = (idx.cpu().numpy() - (np.arange(b) * n)[:, None])
    assert local.min() >= 0 and local.max() < n
    incl = np.bincount(local.ravel(), minlength=n).astype(np.float64)
    first = np.bincount(local[:, 0], minlength=n).astype(np.float64)
    e_incl, e_first = b * m / n, b / n
    var_incl = e_incl * (1 - (m / n if not replace else 1.0 / n))
    chi_incl = ((incl - e_incl) ** 2 / var_incl).sum()
    chi_first = ((first - e_first) ** 2 / e_first).sum()
    # 63 degrees of freedom: P(chi2 > 120) < 2e-5
    assert chi_incl < 120 and chi_first < 120, (chi_incl, chi_first)
    # successive slots of one cloud are not correlated with the point order: mean |idx[j+1] - idx[j]| is (n+1)/3 for a random permutation, (n^2-1)/3n for independent draws
    step = np.abs(np.diff(local, axis=1)).mean()
    expect = (n * n - 1) / (3.0 * n) if replace else (n + 1) / 3.0
    assert abs(step - expect) < 0.25, (step, expect)
