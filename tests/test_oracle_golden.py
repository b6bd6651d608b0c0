"""Pins the CPU oracle (oracle/) against outputs of the UNMODIFIED reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import lmpcr_oracle as O
from oracle import nn_c

NN_CASES = ["s300x700", "s1000", "s5000", "s2049x777"]


@pytest.fixture(scope="module")
def nn_g(golden_dir):
    return np.load(os.path.join(golden_dir, "nn_golden.npz"))


def _nn_inputs(g, name):
    n, m, seed = [int(v) for v in g[name + "_shape"]]
    feats, xyz, _ = O.synth_scene(2, max(n, m), seed=seed)
    return feats[0, :n], feats[1, :m], xyz[0, :n], xyz[1, :m]


@pytest.mark.parametrize("name", NN_CASES)
def test_nn_argmin_bit_exact(nn_g, name):
    """Bit-exact indices AND minimum distances versus lib/utils.py:968-992 + lib/layers.py:81."""
    fs, ft, _, _ = _nn_inputs(nn_g, name)
    idx_st, best = nn_c.nn_argmin(fs, ft)
    idx_ts, _ = nn_c.nn_argmin(ft, fs)
    assert np.array_equal(idx_st, nn_g[name + "_idx_st"])
    assert np.array_equal(idx_ts, nn_g[name + "_idx_ts"])
    assert np.array_equal(best, nn_g[name + "_min_st"])


def test_nn_numpy_matches_c(nn_g):
    fs, ft, _, _ = _nn_inputs(nn_g, "s300x700")
    idx, best = O.nn_argmin_f32(fs, ft)
    idx_c, best_c = nn_c.nn_argmin(fs, ft)
    assert np.array_equal(idx, idx_c) and np.array_equal(best, best_c)
    assert np.array_equal(O.pairwise_distance_f32(fs, ft), nn_c.pairwise_distance(fs, ft))
    assert np.array_equal(O.sqnorm_rows_f32(fs), nn_c.sqnorm(fs))


def test_nn_ties_first_minimum(nn_g):
    feats, _, _ = O.synth_scene(2, 256, seed=5)
    ft = np.concatenate([feats[1][:128], feats[1][:128]], axis=0)
    idx, _ = nn_c.nn_argmin(feats[0], ft)
    assert np.array_equal(idx, nn_g["ties_idx"])
    assert idx.max() < 128


def test_nn_unnormalised_features(nn_g):
    rng = np.random.default_rng(99)
    fa = (rng.standard_normal((400, 32)) * 2).astype(np.float32)
    fb = (rng.standard_normal((600, 32)) * 0.5 + 0.3).astype(np.float32)
    idx, _ = nn_c.nn_argmin(fa, fb)
    assert np.array_equal(idx, nn_g["unnorm_idx"])
    assert np.array_equal(nn_c.pairwise_distance(fa, fb)[:64, :64], nn_g["unnorm_dist"])


def test_mutuals_and_xs(nn_g):
    fs, ft, xs_, xt = _nn_inputs(nn_g, "s1000")
    idx_st, idx_ts, mutual_geo, xs = O.register_pair_stage1(fs, ft, xs_, xt, mutual_mode="geometric")
    # lib/utils.py:822-848 on hard matches == index chase + 5 cm test
    assert np.array_equal(mutual_geo, nn_g["s1000_mutual_geo"])
    c_st, c_ts = xt[idx_st], xs_[idx_ts]
    assert np.array_equal(O.extract_mutuals(xs_, xt, c_st, c_ts).astype(np.uint8), nn_g["s1000_mutual_geo"])
    assert xs.shape == (1, 1000, 6)
    assert abs(float(xs.astype(np.float64).sum()) - float(nn_g["s1000_xs_sum"])) < 1e-6
    mi = O.mutual_index(idx_st, idx_ts)
    assert mi.sum() > 0 and np.all(mutual_geo >= mi)     # index-mutual implies geometric-mutual (Q3)


@pytest.mark.parametrize("name", ["s300x700", "s1000"])
def test_soft_correspondences(nn_g, name):
    """Soft (non-ST) correspondences vs the reference's Soft_NN('soft', st=False): 2e-5 m (fp32 softmax over 700-1000 terms)."""
    fs, ft, _, xt = _nn_inputs(nn_g, name)
    T = float(nn_g[name + "_soft_T"])
    assert abs(T - 0.09) < 1e-6
    out = O.soft_correspondences(fs, ft, xt, T)
    assert np.abs(out - nn_g[name + "_soft_st"]).max() < 2e-5


def test_pair_enumeration():
    p = O.enumerate_pairs(5)
    assert p.shape == (10, 2) and p[0].tolist() == [0, 1] and p[-1].tolist() == [3, 4] and np.all(p[:, 0] < p[:, 1])


KB_CASES = ["n3", "n4", "n50", "n5000", "n777"]


@pytest.mark.parametrize("name", KB_CASES)
def test_kabsch_vs_reference(golden_dir, name):
    """Tolerance 1e-5 rad (chordal) / 1e-5 m, the north-star gate for stage 3 (same inputs)."""
    g = np.load(os.path.join(golden_dir, "kabsch_golden.npz"))
    P, N, seed = [int(v) for v in g[name + "_cfg"]]
    xs, _, _ = O.synth_xs(P, N, inlier_frac=0.5, seed=seed)
    for dtype in (np.float32, np.float64):
        R, t, res, flag = O.kabsch(xs[:, 0, :, :3], xs[:, 0, :, 3:], g[name + "_w"], dtype=dtype)
        assert not flag
        assert O.chordal_angle(R, g[name + "_R"]).max() < 1e-5
        assert np.abs(t - g[name + "_t"]).max() < 1e-5
        assert np.abs(res - g[name + "_res"]).max() < 1e-5
        assert np.allclose(np.linalg.det(R.astype(np.float64)), 1.0, atol=1e-5)


OA_CASES = ["full_p2_n2000", "full_p1_n5000", "small_p3_n64", "guard_p2_n500"]
# Gates for stage 2 (+3 on its outputs).  The reference's own fp32-vs-fp64 spread on these inputs is
# ~1e-4 in logits, 5e-5 rad, 1e-4 m (DESIGN.md "tolerances"); the gates are 5x that.
LOGIT_TOL, ROT_TOL, TRANS_TOL = 5e-4, 5e-4, 1e-3


@pytest.mark.parametrize("name", OA_CASES)
def test_oanet_vs_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "oanet_golden.npz"))
    P, N, seed, small, guard = [int(v) for v in g[name + "_cfg"]]
    kw = dict(net_channel=32, clusters=16) if small else {}
    sd = O.synth_state_dict(seed, **kw)
    if guard:
        sd["reg_init.output.bias"] = np.full((1,), -50.0, np.float32)
    xs, _, _ = O.synth_xs(P, N, seed=seed)
    out = O.oanet_forward(xs, sd, dtype=np.float32)
    for it in range(2):
        assert np.abs(out["logits"][it] - g["%s_logits%d" % (name, it)]).max() < LOGIT_TOL
        assert np.abs(out["scores"][it] - g["%s_scores%d" % (name, it)]).max() < LOGIT_TOL
        assert O.chordal_angle(out["rot_est"][it], g["%s_R%d" % (name, it)]).max() < ROT_TOL
        assert np.abs(out["trans_est"][it] - g["%s_t%d" % (name, it)]).max() < TRANS_TOL
    if guard:
        assert np.allclose(out["scores"][0], 1.0 / N)
    assert bool(out["gradient_flag"]) == bool(g[name + "_flag"])
    assert abs(float(np.abs(out["latent features"]).mean()) - float(g[name + "_latent_absmean"])) < 1e-3


def test_param_schema_counts():
    sch = O.oanet_param_schema()
    assert len(sch) == 334
    n_par = sum(int(np.prod(s)) for n, s in sch if not n.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_par == 2473050
