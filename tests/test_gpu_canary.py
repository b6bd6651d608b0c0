"""Out-of-bounds guard for the caller-provided workspaces (compute-sanitizer is not available on this GPU pool): every workspace
is carved out of a larger buffer whose head and tail are filled with a byte pattern; after the call both guard bands must be
intact.  The kernels get exactly the number of bytes the size queries return."""
import importlib

import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from util import cabi, cu, load_oanet

GUARD = 1 << 16


class _Guarded:
    def __init__(self):
        self.bufs = []

    def __call__(self, nbytes, device):
        n = max(int(nbytes), 256)
        n_al = (n + 255) // 256 * 256
        base = torch.full((n_al + 2 * GUARD,), 0xAB, dtype=torch.uint8, device=device)
        self.bufs.append((base, n))
        return base[GUARD:GUARD + n]

    def check(self):
        assert self.bufs, "no workspace was requested"
        for base, n in self.bufs:
            assert bool((base[:GUARD] == 0xAB).all()), "write below the workspace"
            assert bool((base[GUARD + n:] == 0xAB).all()), "write past the workspace (%d bytes)" % n


@pytest.fixture
def guarded(monkeypatch):
    g = _Guarded()
    monkeypatch.setattr(cabi, "_ws", g)
    yield g
    g.check()


@pytest.mark.gpu
@pytest.mark.parametrize("n,algo", [(1000, 0), (1000, 1), (777, 1), (5000, 1)])
def test_nn_workspace_guard(guarded, n, algo):
    feats, _, _ = O.synth_scene(3, n, seed=71)
    f = cu(feats)
    jobs = torch.tensor([[0, 1], [1, 0], [2, 1]], dtype=torch.int32, device="cuda")
    idx = cabi.nn_argmin(f, f, jobs, algo=algo)
    ri, _ = O.nn_argmin_f32(feats[0], feats[1])
    assert np.array_equal(idx[0].cpu().numpy(), ri)
    cabi.nn_top2(f, f, jobs)


@pytest.mark.gpu
@pytest.mark.parametrize("P,N,algo,train", [(3, 1000, 1, False), (2, 777, 1, False), (5, 2000, 1, False), (3, 500, 0, False), (3, 1000, 1, True)])
def test_filter_workspace_guard(guarded, P, N, algo, train):
    sd = O.synth_state_dict(72)
    xs, _, _ = O.synth_xs(P, N, seed=72)
    net = load_oanet(sd, gemm_algo=algo)
    if train:
        net.train()
    out = net({"xs": torch.from_numpy(xs)})
    assert torch.isfinite(out["logits"][-1]).all() and torch.isfinite(out["rot_est"][-1]).all()


@pytest.mark.gpu
def test_pool_conv_overlap_workspace_guard(guarded):
    rng = np.random.default_rng(5)
    x = cu(rng.standard_normal((2, 128, 1000)).astype(np.float32))
    e = cu(rng.standard_normal((2, 500, 1000)).astype(np.float32))
    for mode in (0, 1):
        assert torch.isfinite(cabi.softmax_pool(x, e, mode)).all()
    w = cu(rng.standard_normal((128, 128)).astype(np.float32))
    assert torch.isfinite(cabi.conv1x1(x, w, gemm_algo=1)).all()
    pts = torch.from_numpy(rng.uniform(0, 2, (5000, 3))).cuda()
    assert cabi.overlap_count(pts, pts, None, 0.05) == 5000
    assert cabi.voxel_downsample(pts, 0.1).shape[0] > 0
