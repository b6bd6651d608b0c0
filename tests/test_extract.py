"""Correspondence extraction: oracle vs the golden written by the reference's own
scripts/extract_data.py:run_correspondence_extraction (CPU), and the CUDA top-2 path vs both (GPU)."""
import os

import numpy as np
import pytest

from oracle import lmpcr_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "extract_golden.npz")
RATIO_TOL = 1e-4        # sklearn evaluates float32 inputs through a float64 GEMM; ours is the torch fp32 formula


def _cases():
    g = np.load(GOLD)
    S, n, seed, _ = [int(v) for v in g["cfg"]]
    feats, xyz, _ = O.synth_scene(S, n, seed=seed)
    for a in range(S):
        for b in range(a + 1, S):
            k = "p%d_%d_" % (a, b)
            i1, i2 = g[k + "inds1"].astype(np.int64), g[k + "inds2"].astype(np.int64)
            yield k, g, feats[a][i1], xyz[a][i1], feats[b][i2], xyz[b][i2]


def _check(k, g, x, mutuals, ratios):
    gx, gm, gr = g[k + "x"], g[k + "mutuals"], g[k + "ratios"]
    assert x.shape == gx.shape and mutuals.shape == (gx.shape[0], 1) and ratios.shape == gr.shape
    assert str(mutuals.dtype) == str(g[k + "dtypes"][1]) == "float64" and str(ratios.dtype) == str(g[k + "dtypes"][2])
    assert np.array_equal(x, gx)                                   # same matches -> identical coordinates
    assert np.array_equal(mutuals[:, 0].astype(np.uint8), gm[:, 0])
    assert np.max(np.abs(ratios - gr)) < RATIO_TOL


def test_oracle_matches_reference_extraction():
    n = 0
    for k, g, f1, k1, f2, k2 in _cases():
        x, mutuals, ratios, _, _ = O.extract_correspondences(f1, k1, f2, k2)
        _check(k, g, x.astype(g[k + "x"].dtype), mutuals, ratios)
        n += 1
    assert n == 3


def test_oracle_top2_is_sorted_and_distinct():
    feats, _, _ = O.synth_scene(2, 400, seed=3)
    idx, d = O.nn_top2_f32(feats[0], feats[1])
    assert np.all(d[:, 0] <= d[:, 1]) and np.all(idx[:, 0] != idx[:, 1])
    i1, d1 = O.nn_argmin_f32(feats[0], feats[1])
    assert np.array_equal(idx[:, 0], i1) and np.array_equal(d[:, 0], d1)


@pytest.mark.gpu
def test_gpu_extraction_matches_reference_golden_and_oracle(tmp_path):
    import importlib
    import torch
    ext = importlib.import_module("3d_multiview_reg_b200.extract")
    cabi = importlib.import_module("3d_multiview_reg_b200._cabi")
    for k, g, f1, k1, f2, k2 in _cases():
        feats = torch.from_numpy(np.stack([f1, f2])).cuda()
        xyz = torch.from_numpy(np.stack([k1, k2])).cuda()
        before = cabi.launch_count()
        c = ext.scene_correspondences(feats, xyz)
        assert cabi.launch_count() > before
        _check(k, g, c["x"][0].cpu().numpy(), c["mutuals"][0].cpu().numpy(), c["ratios"][0].cpu().numpy())
        # bit-exact against the restatement (indices and squared distances of both searches)
        _, _, _, nn, nn1 = O.extract_correspondences(f1, k1, f2, k2)
        assert np.array_equal(c["nn_12"][0].cpu().numpy(), nn) and np.array_equal(c["nn_21"][0].cpu().numpy(), nn1)
        pairs = torch.tensor([[0, 1]], dtype=torch.int32)
        files = ext.save_scene_correspondences(str(tmp_path / k), "scene0", c, pairs)
        d = np.load(files[0])
        assert sorted(d.files) == ["mutuals", "ratios", "x"] and os.path.basename(files[0]) == "scene0_000_001.npz"
        assert ext.save_scene_correspondences(str(tmp_path / k), "scene0", c, pairs) == []      # existing files are skipped (:148-149)


@pytest.mark.gpu
def test_gpu_top2_exact_on_ties_and_ragged_sizes():
    import importlib
    import torch
    cabi = importlib.import_module("3d_multiview_reg_b200._cabi")
    feats, _, _ = O.synth_scene(2, 777, seed=17)
    fb = np.concatenate([feats[1][:300], feats[1][:300]], 0)               # every target row duplicated: d1 == d2 exactly
    q = torch.from_numpy(feats[0][None, :515].copy()).cuda()
    b = torch.from_numpy(fb[None]).cuda()
    jobs = torch.tensor([[0, 0]], dtype=torch.int32, device="cuda")
    oi, od = O.nn_top2_f32(feats[0][:515], fb)
    for algo in (None, cabi.NN_EXACT_SIMT):            # tcgen05 screening + exact rescoring (default for 32-d) and the exact CUDA-core kernel
        idx, dist = cabi.nn_top2(q, b, jobs, algo=algo)
        assert np.array_equal(idx[0].cpu().numpy(), oi) and np.array_equal(dist[0].cpu().numpy(), od), algo
    assert np.array_equal(oi[:, 1], oi[:, 0] + 300)
    with pytest.raises(cabi.LmpcrError):
        cabi.nn_top2(q, b[:, :1].contiguous(), jobs)


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(5000, 5000), (2049, 777), (130, 4100)])
def test_gpu_top2_tensor_path_is_bit_identical_to_the_exact_kernel(n, m):
    """Two nearest neighbours on the tcgen05 path (chunks within the margin of the SECOND smallest chunk minimum, exact rescoring) against the
    exact CUDA-core kernel: indices and squared distances identical, both directions of a pair, un-normalised features included."""
    import importlib
    import torch
    cabi = importlib.import_module("3d_multiview_reg_b200._cabi")
    rng = np.random.default_rng(n * 7 + m)
    fa = rng.standard_normal((1, n, 32)).astype(np.float32)
    fb = rng.standard_normal((1, m, 32)).astype(np.float32)
    fa /= np.linalg.norm(fa, axis=2, keepdims=True)
    fb *= (0.5 + rng.random((1, m, 1))).astype(np.float32) / np.linalg.norm(fb, axis=2, keepdims=True)      # norms between 0.5 and 1.5
    q, b = torch.from_numpy(fa).cuda(), torch.from_numpy(fb).cuda()
    jobs = torch.tensor([[0, 0]], dtype=torch.int32, device="cuda")
    it, dt = cabi.nn_top2(q, b, jobs, algo=cabi.NN_TENSOR)
    ie, de = cabi.nn_top2(q, b, jobs, algo=cabi.NN_EXACT_SIMT)
    assert torch.equal(it, ie) and torch.equal(dt, de)
    it2, dt2 = cabi.nn_top2(b, q, jobs, algo=cabi.NN_TENSOR)
    ie2, de2 = cabi.nn_top2(b, q, jobs, algo=cabi.NN_EXACT_SIMT)
    assert torch.equal(it2, ie2) and torch.equal(dt2, de2)
