"""GPU parity, stage 3: warp-per-pair weighted Kabsch vs oracle / reference goldens.
Gate (north star): rotation < 1e-5 rad (chordal metric), translation and residuals < 1e-5 m."""
import os

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from util import cabi, cu

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["n3", "n4", "n50", "n5000", "n777"])
def test_kabsch_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "kabsch_golden.npz"))
    P, N, seed = [int(v) for v in g[name + "_cfg"]]
    xs, _, _ = O.synth_xs(P, N, inlier_frac=0.5, seed=seed)
    R, t, res, status = cabi.kabsch_xs(cu(xs), cu(g[name + "_w"]))
    assert int(status.sum().item()) == 0
    assert O.chordal_angle(R.cpu().numpy(), g[name + "_R"]).max() < TOL
    assert np.abs(t.cpu().numpy() - g[name + "_t"]).max() < TOL
    assert np.abs(res.cpu().numpy() - g[name + "_res"]).max() < TOL
    # separate [P,N,3] tensors (lib.utils.kabsch_transformation_estimation signature) give the same result
    R2, t2, res2, _ = cabi.kabsch_points(cu(xs[:, 0, :, :3]), cu(xs[:, 0, :, 3:]), cu(g[name + "_w"]))
    assert torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(res, res2)


def test_kabsch_recovers_planted_transform_and_properties():
    xs, Rs, ts = O.synth_xs(64, 2000, inlier_frac=1.0, seed=21, noise=0.0)
    w = np.random.default_rng(1).uniform(0.1, 1, (64, 2000)).astype(np.float32)
    R, t, res, status, conf = cabi.kabsch_xs(cu(xs), cu(w), want_conf=True)
    R, t = R.cpu().numpy(), t.cpu().numpy()
    assert O.chordal_angle(R, Rs).max() < 1e-5 and np.abs(t[:, :, 0] - ts).max() < 2e-5
    assert np.abs(np.linalg.det(R.astype(np.float64)) - 1).max() < 1e-5
    assert res.max().item() < 1e-4
    # invariance to weight scaling
    R3, t3, _, _ = cabi.kabsch_xs(cu(xs), cu(w * 7.0))
    assert O.chordal_angle(R3.cpu().numpy(), R).max() < 1e-5
    ref = O.pair_confidence(w, res.cpu().numpy())
    c = conf.cpu().numpy()
    assert np.array_equal(c[:, 0], ref[:, 0]) and np.allclose(c[:, 1], ref[:, 1], rtol=1e-5) and np.array_equal(c[:, 3], ref[:, 3])


def test_kabsch_reflection_and_degenerate():
    rng = np.random.default_rng(5)
    # planar, noise-free correspondences (sigma_3 = 0): the determinant fix must still give a proper rotation
    x1 = rng.uniform(0, 1, (1, 100, 3)).astype(np.float32)
    x1[..., 2] = 0
    Rg = O.random_rotation(rng)
    x2 = (x1 @ Rg.T + 0.3).astype(np.float32)
    R, t, res, status = cabi.kabsch_points(cu(x1), cu(x2), cu(np.ones((1, 100), np.float32)))
    assert abs(np.linalg.det(R[0].cpu().numpy().astype(np.float64)) - 1) < 1e-5 and res.max().item() < 1e-5
    # collinear points: rank 1 -> R = I, t = centroid offset (lib/utils.py:232 with R = I) + DEGENERATE status
    x1 = np.zeros((1, 50, 3), np.float32)
    x1[0, :, 0] = np.linspace(0, 1, 50)
    R, t, res, status = cabi.kabsch_points(cu(x1), cu(x1 + 1), cu(np.ones((1, 50), np.float32)))
    assert status[0].item() & cabi.STATUS_DEGENERATE
    assert np.array_equal(R[0].cpu().numpy(), np.eye(3, dtype=np.float32)) and np.abs(t.cpu().numpy() - 1.0).max() < 1e-6


def test_zero_weight_guard_modes():
    xs, _, _ = O.synth_xs(3, 200, seed=2)
    w = np.random.default_rng(0).uniform(0, 1, (3, 200)).astype(np.float32)
    w[1] = 0
    R, t, res, status = cabi.kabsch_xs(cu(xs), cu(w), guard_mode=cabi.GUARD_PAIR)
    assert status.cpu().tolist() == [0, cabi.STATUS_ZERO_WEIGHT, 0]
    Ro, to, _, _ = O.kabsch(xs[:, 0, :, :3], xs[:, 0, :, 3:], np.where(w.sum(1, keepdims=True) == 0, 1.0 / 200, w).astype(np.float32))
    assert O.chordal_angle(R.cpu().numpy(), Ro).max() < TOL and np.abs(t.cpu().numpy() - to).max() < TOL


def test_residuals_mirror():
    import importlib
    U = importlib.import_module("3d_multiview_reg_b200.lib.utils")
    xs, Rs, ts = O.synth_xs(4, 333, seed=8)
    R, t, res, flag = U.kabsch_transformation_estimation(cu(xs[:, 0, :, :3]), cu(xs[:, 0, :, 3:]), None)
    assert flag is False and tuple(R.shape) == (4, 3, 3) and tuple(t.shape) == (4, 3, 1) and tuple(res.shape) == (4, 333)
    r2 = U.transformation_residuals(cu(xs[:, 0, :, :3]), cu(xs[:, 0, :, 3:]), R, t)
    assert (r2 - res).abs().max().item() < 1e-6
    ro = O.transformation_residuals(xs[:, 0, :, :3], xs[:, 0, :, 3:], R.cpu().numpy(), t.cpu().numpy())
    assert np.abs(r2.cpu().numpy() - ro).max() < TOL
