"""Overlap ratio under an estimated pose (lib/utils.py:713-786): the oracle against a golden written by
the reference's own compute_overlap_ratio (CPU), the hash-grid CUDA path against both (GPU)."""
import os

import numpy as np
import pytest

from oracle import lmpcr_oracle as O
import synthdata

GOLD = os.path.join(os.path.dirname(__file__), "golden", "overlap_golden.npz")


def _cases():
    g = np.load(GOLD)
    for name in ("a", "b", "c", "d"):
        seed, ni, nj, ov = g[name + "_cfg"]
        pi, pj, T = synthdata.synth_cloud_pair(int(seed), int(ni), int(nj), float(ov))
        assert np.array_equal(T, g[name + "_T"])
        yield name, g, pi, pj, T


def test_oracle_matches_reference_overlap():
    for name, g, pi, pj, T in _cases():
        r, _, _ = O.compute_overlap_ratio(pi, pj, T, method="3DMatch")
        assert r == float(g[name + "_3dmatch"])                       # counts are integers: exact
        r, _, _ = O.compute_overlap_ratio(pi, pj, T, method="FCGF", voxel_size=0.025)
        assert r == float(g[name + "_fcgf"])


def test_oracle_voxel_down_sample_properties():
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1, 1, (4000, 3))
    d = O.voxel_down_sample(pts, 0.1)
    assert len(d) < len(pts) and np.all(d.min(0) >= pts.min(0)) and np.all(d.max(0) <= pts.max(0))
    np.testing.assert_allclose(d.mean(0) * 0 + pts.mean(0), pts.mean(0))
    # every point lies in the voxel of exactly one output point; weighted mean of outputs = mean of inputs
    origin = pts.min(0) - 0.05
    ijk = np.floor((pts - origin) / 0.1).astype(np.int64)
    _, counts = np.unique(ijk, axis=0, return_counts=True)
    assert len(counts) == len(d)
    np.testing.assert_allclose((d * counts[:, None]).sum(0) / len(pts), pts.mean(0), rtol=1e-12)
    assert len(O.voxel_down_sample(pts[:1], 0.1)) == 1


@pytest.mark.gpu
def test_gpu_overlap_matches_reference_golden():
    import importlib
    import torch
    utils = importlib.import_module("3d_multiview_reg_b200.lib.utils")
    cabi = importlib.import_module("3d_multiview_reg_b200._cabi")
    for name, g, pi, pj, T in _cases():
        before = cabi.launch_count()
        assert utils.compute_overlap_ratio(pi, pj, T, method="3DMatch") == float(g[name + "_3dmatch"])
        assert cabi.launch_count() > before
        assert utils.compute_overlap_ratio(torch.from_numpy(pi), torch.from_numpy(pj).cuda(), T, method="FCGF") == float(g[name + "_fcgf"])
        _, m01, m10 = O.compute_overlap_ratio(pi, pj, T)
        dev = lambda a: torch.from_numpy(a).cuda()
        assert cabi.overlap_count(dev(pi), dev(pj), dev(T), 0.05) == m01
        assert cabi.overlap_count(dev(pj), dev(pi), dev(np.linalg.inv(T)), 0.05) == m10
    with pytest.raises(ValueError):
        utils.compute_overlap_ratio(pi, pj, T, method="nope")


@pytest.mark.gpu
def test_gpu_voxel_downsample_and_edge_cases():
    import importlib
    import torch
    cabi = importlib.import_module("3d_multiview_reg_b200._cabi")
    rng = np.random.default_rng(5)
    for n, voxel in ((1, 0.1), (37, 0.5), (20000, 0.025), (50000, 0.2)):
        pts = rng.uniform(-3, 3, (n, 3))
        got = cabi.voxel_downsample(torch.from_numpy(pts).cuda(), voxel).cpu().numpy()
        want = O.voxel_down_sample(pts, voxel)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)     # same voxel order (x, y, z keys), means in fp64
    # identical clouds, identity pose -> everything overlaps; disjoint clouds -> nothing; empty target -> 0
    pts = torch.from_numpy(rng.uniform(0, 1, (3000, 3))).cuda()
    assert cabi.overlap_count(pts, pts, None, 0.05) == 3000
    assert cabi.overlap_count(pts, pts + 10.0, None, 0.05) == 0
    assert cabi.overlap_count(pts, pts[:0], None, 0.05) == 0
    # a point exactly `radius` away is NOT within (strict <, lib/utils.py:748)
    a = torch.tensor([[0.0, 0.0, 0.0]], dtype=torch.float64).cuda()
    assert cabi.overlap_count(a, a + torch.tensor([0.05, 0.0, 0.0], dtype=torch.float64).cuda(), None, 0.05) == 0
    assert cabi.overlap_count(a, a + torch.tensor([0.0499999, 0.0, 0.0], dtype=torch.float64).cuda(), None, 0.05) == 1
    with pytest.raises(cabi.LmpcrError):
        cabi.overlap_count(a * 1e9 + 1e9, a, None, 0.05)
