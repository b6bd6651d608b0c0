"""GPU parity, stage 1: CUDA nearest-neighbour path (through the C ABI) vs the CPU oracle and vs the golden
vectors produced by the unmodified reference.  Gate: BIT-EXACT indices (and distances)."""
import os

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from oracle import nn_c
from util import cabi, cu

pytestmark = pytest.mark.gpu
ALGOS = [cabi.NN_EXACT_SIMT, cabi.NN_TENSOR]


def _jobs(pairs):
    return torch.tensor(pairs, dtype=torch.int32, device="cuda")


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", ["s300x700", "s1000", "s5000", "s2049x777"])
def test_nn_vs_reference_golden(golden_dir, name, algo):
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    n, m, seed = [int(v) for v in g[name + "_shape"]]
    feats, xyz, _ = O.synth_scene(2, max(n, m), seed=seed)
    fs, ft = cu(feats[0:1, :n]), cu(feats[1:2, :m])
    idx_st, d_st = cabi.nn_argmin(fs, ft, _jobs([[0, 0]]), algo=algo, return_dist=True)
    idx_ts = cabi.nn_argmin(ft, fs, _jobs([[0, 0]]), algo=algo)
    assert np.array_equal(idx_st[0].cpu().numpy(), g[name + "_idx_st"])
    assert np.array_equal(idx_ts[0].cpu().numpy(), g[name + "_idx_ts"])
    assert np.array_equal(d_st[0].cpu().numpy(), g[name + "_min_st"])


@pytest.mark.parametrize("algo", ALGOS)
def test_nn_scene_jobs_vs_oracle(algo):
    """4 scans -> 6 pairs x 2 directions in ONE call, features never copied per pair."""
    feats, xyz, _ = O.synth_scene(4, 1500, seed=3)
    pairs = O.enumerate_pairs(4)
    f = cu(feats)
    jobs = np.concatenate([pairs, pairs[:, ::-1]], 0)
    idx, dist = cabi.nn_argmin(f, f, _jobs(jobs.tolist()), algo=algo, return_dist=True)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for j, (a, b) in enumerate(jobs):
        ref_i, ref_d = nn_c.nn_argmin(feats[a], feats[b])
        assert np.array_equal(idx[j], ref_i) and np.array_equal(dist[j], ref_d)


@pytest.mark.parametrize("algo", ALGOS)
def test_nn_ties_and_unnormalised(golden_dir, algo):
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    feats, _, _ = O.synth_scene(2, 256, seed=5)
    ft = np.concatenate([feats[1][:128], feats[1][:128]], axis=0)
    idx = cabi.nn_argmin(cu(feats[0:1]), cu(ft[None]), _jobs([[0, 0]]), algo=algo)
    assert np.array_equal(idx[0].cpu().numpy(), g["ties_idx"])          # first minimum wins
    if True:   # un-normalised features: the screening margin adapts to the actual norms / rounding errors
        rng = np.random.default_rng(99)
        fa = (rng.standard_normal((400, 32)) * 2).astype(np.float32)
        fb = (rng.standard_normal((600, 32)) * 0.5 + 0.3).astype(np.float32)
        idx = cabi.nn_argmin(cu(fa[None]), cu(fb[None]), _jobs([[0, 0]]), algo=algo)
        assert np.array_equal(idx[0].cpu().numpy(), g["unnorm_idx"])


@pytest.mark.parametrize("dim", [8, 16, 64])
def test_nn_other_dims(dim):
    rng = np.random.default_rng(dim)
    a = rng.standard_normal((333, dim)).astype(np.float32)
    b = rng.standard_normal((500, dim)).astype(np.float32)
    idx, d = cabi.nn_argmin(cu(a[None]), cu(b[None]), _jobs([[0, 0]]), return_dist=True)
    ri, rd = nn_c.nn_argmin(a, b)
    assert np.array_equal(idx[0].cpu().numpy(), ri) and np.array_equal(d[0].cpu().numpy(), rd)


def test_edge_sizes():
    rng = np.random.default_rng(0)
    for n, m in [(1, 1), (1, 129), (127, 1), (129, 3)]:
        a = rng.standard_normal((n, 32)).astype(np.float32)
        b = rng.standard_normal((m, 32)).astype(np.float32)
        idx = cabi.nn_argmin(cu(a[None]), cu(b[None]), _jobs([[0, 0]]))
        assert np.array_equal(idx[0].cpu().numpy(), nn_c.nn_argmin(a, b)[0])
    # empty job list
    out = cabi.nn_argmin(cu(a[None]), cu(b[None]), torch.zeros((0, 2), dtype=torch.int32, device="cuda"))
    assert out.shape == (0, a.shape[0])


def test_pairwise_distance_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    rng = np.random.default_rng(99)
    fa = (rng.standard_normal((400, 32)) * 2).astype(np.float32)
    fb = (rng.standard_normal((600, 32)) * 0.5 + 0.3).astype(np.float32)
    d = cabi.pairwise_distance(cu(fa[None]), cu(fb[None]))[0].cpu().numpy()
    assert np.array_equal(d[:64, :64], g["unnorm_dist"])
    assert np.array_equal(d, nn_c.pairwise_distance(fa, fb))


def test_mutual_xs_gather_knn(golden_dir):
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    feats, xyz, _ = O.synth_scene(2, 1000, seed=12)
    f, x = cu(feats), cu(xyz)
    pairs = _jobs([[0, 1]])
    idx_st = cabi.nn_argmin(f, f, pairs)
    idx_ts = cabi.nn_argmin(f, f, pairs.flip(1).contiguous())
    o_st, o_ts, o_mut, o_xs = O.register_pair_stage1(feats[0], feats[1], xyz[0], xyz[1], mutual_mode="geometric")
    m_geo, xs = cabi.mutual_xs(x, pairs, idx_st, idx_ts, cabi.MUTUAL_GEOMETRIC, 0.05)
    assert np.array_equal(m_geo[0].cpu().numpy(), g["s1000_mutual_geo"])      # reference's extract_mutuals
    assert np.array_equal(xs[0].cpu().numpy(), o_xs)
    m_idx, xs7 = cabi.mutual_xs(x, pairs, idx_st, idx_ts, cabi.MUTUAL_INDEX, 0.05, xs_channels=7)
    assert np.array_equal(m_idx[0].cpu().numpy(), O.mutual_index(o_st, o_ts))
    assert np.array_equal(xs7[0, 0, :, 6].cpu().numpy(), O.mutual_index(o_st, o_ts).astype(np.float32))
    # Soft_NN('hard') output == gathered target coordinates (lib/layers.py:86)
    corr = cabi.gather_xyz(x, pairs, idx_st)
    assert np.array_equal(corr[0].cpu().numpy(), xyz[1][o_st])
    # knn_point(k=1) (lib/utils.py:274) and the module-level mirror of extract_mutuals
    sq, idx = cabi.knn3d_1(x[1:2], corr)
    assert np.array_equal(idx[0].cpu().numpy().astype(np.int64), O.knn_point_1(xyz[1], xyz[1][o_st]))
    import importlib
    U = importlib.import_module("3d_multiview_reg_b200.lib.utils")
    L = importlib.import_module("3d_multiview_reg_b200.lib.layers")
    back = cabi.gather_xyz(x, pairs.flip(1).contiguous(), idx_ts)
    mut = U.extract_mutuals(x[0:1], x[1:2], corr, back)
    assert np.array_equal(mut[0].cpu().numpy().astype(np.uint8), g["s1000_mutual_geo"])
    hard = L.Soft_NN(corr_type="hard", device="cuda")
    assert np.array_equal(hard(f[0:1], f[1:2], x[1:2])[0].cpu().numpy(), xyz[1][o_st])


def test_tensor_path_screening_scores_and_margin():
    """The raw tcgen05 scores equal |b|^2 - 2 a_hat.b_hat (fp16-rounded operands) to fp32 accumulation accuracy, and
    the fp16 range guard (features outside the operand range) still yields exact indices via full rescoring."""
    feats, _, _ = O.synth_scene(2, 700, seed=11)
    fs, ft = feats[0, :300], feats[1, :700]
    idx, dist, sc, amin = cabi.nn_tensor_debug(cu(fs[None]), cu(ft[None]), _jobs([[0, 0]]))
    sc = sc[0].cpu().numpy()[:, :700]
    ah, bh = fs.astype(np.float16).astype(np.float64), ft.astype(np.float16).astype(np.float64)
    exp = nn_c.sqnorm(ft).astype(np.float64)[None, :] - 2 * ah @ bh.T
    assert np.abs(sc - exp).max() < 5e-6
    ri, rd = nn_c.nn_argmin(fs, ft)
    assert np.array_equal(idx[0].cpu().numpy(), ri) and np.array_equal(dist[0].cpu().numpy(), rd)
    big = (np.random.default_rng(3).standard_normal((200, 32)) * 3e4).astype(np.float32)     # overflows fp16 operands
    idx = cabi.nn_argmin(cu(big[None]), cu(big[None, ::-1].copy()), _jobs([[0, 0]]), algo=cabi.NN_TENSOR)
    assert np.array_equal(idx[0].cpu().numpy(), nn_c.nn_argmin(big, big[::-1].copy())[0])


def test_tensor_path_many_jobs_persistent_grid():
    """More work items than SMs (persistent CTAs loop; barrier phases carry across items) and n not a tile multiple."""
    feats, _, _ = O.synth_scene(6, 1100, seed=8)
    pairs = O.enumerate_pairs(6)
    jobs = np.concatenate([pairs, pairs[:, ::-1]], 0)
    f = cu(feats)
    idx, dist = cabi.nn_argmin(f, f, _jobs(jobs.tolist()), algo=cabi.NN_TENSOR, return_dist=True)
    ref = cabi.nn_argmin(f, f, _jobs(jobs.tolist()), algo=cabi.NN_EXACT_SIMT)
    assert torch.equal(idx, ref)
    for j in (0, 7, 29):
        ri, rd = nn_c.nn_argmin(feats[jobs[j][0]], feats[jobs[j][1]])
        assert np.array_equal(idx[j].cpu().numpy(), ri) and np.array_equal(dist[j].cpu().numpy(), rd)


@pytest.mark.parametrize("name", ["s300x700", "s1000"])
def test_soft_correspondences_vs_reference(golden_dir, name):
    """corr_type='soft', st=False (demo config): online-softmax kernel vs the reference's output and the fp64 oracle."""
    import importlib
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    n, m, seed = [int(v) for v in g[name + "_shape"]]
    feats, xyz, _ = O.synth_scene(2, max(n, m), seed=seed)
    fs, ft, xt = feats[0:1, :n], feats[1:2, :m], xyz[1:2, :m]
    T = float(g[name + "_soft_T"])
    out = cabi.nn_soft(cu(fs), cu(ft), cu(xt), _jobs([[0, 0]]), T)[0].cpu().numpy()
    assert np.abs(out - g[name + "_soft_st"]).max() < 2e-5
    assert np.abs(out - O.soft_correspondences(fs[0], ft[0], xt[0], T)).max() < 2e-5
    L = importlib.import_module("3d_multiview_reg_b200.lib.layers")
    soft = L.Soft_NN(corr_type="soft", st=False, device="cuda").cuda()
    assert np.abs(soft(cu(fs), cu(ft), cu(xt))[0].cpu().numpy() - g[name + "_soft_st"]).max() < 2e-5
    # sharp temperature -> converges to the hard match
    hard = cabi.gather_xyz(cu(xt), _jobs([[0, 0]]), cabi.nn_argmin(cu(fs), cu(ft), _jobs([[0, 0]])))[0].cpu().numpy()
    sharp = cabi.nn_soft(cu(fs), cu(ft), cu(xt), _jobs([[0, 0]]), 1e-4)[0].cpu().numpy()
    assert np.median(np.abs(sharp - hard).max(axis=1)) < 1e-6
