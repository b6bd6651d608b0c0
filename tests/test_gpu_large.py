"""GPU: BASELINE.json's larger configurations -- dense keypoints (50k x 50k, configs[2]) and the sweep sizes of configs[4] --
against the CPU oracle where it finishes in seconds, and through size-independent properties otherwise."""
import numpy as np
import pytest
import torch

from oracle import lmpcr_oracle as O
from oracle import nn_c
from util import cabi, cu, load_oanet

pytestmark = pytest.mark.gpu


def _jobs(pairs):
    return torch.tensor(pairs, dtype=torch.int32, device="cuda")


@pytest.mark.parametrize("n", [10000, 20000])
def test_nn_sweep_sizes_bit_exact(n):
    feats, _, _ = O.synth_scene(2, n, seed=n)
    f = cu(feats)
    idx, dist = cabi.nn_argmin(f, f, _jobs([[0, 1], [1, 0]]), algo=cabi.NN_TENSOR, return_dist=True)
    for j, (a, b) in enumerate([(0, 1), (1, 0)]):
        ri, rd = nn_c.nn_argmin(feats[a], feats[b])
        assert np.array_equal(idx[j].cpu().numpy(), ri) and np.array_equal(dist[j].cpu().numpy(), rd)


def test_dense_50k_pair():
    """configs[2]: one pair at 50,000 x 50,000 x 32 (the reference cannot run this: >= 10 GB per intermediate)."""
    n = 50000
    feats, xyz, poses = O.synth_scene(2, n, seed=50)
    f, x = cu(feats), cu(xyz)
    pairs = _jobs([[0, 1]])
    idx_st, d_st = cabi.nn_argmin(f, f, pairs, algo=cabi.NN_TENSOR, return_dist=True)
    idx_ts = cabi.nn_argmin(f, f, pairs.flip(1).contiguous(), algo=cabi.NN_TENSOR)
    ri, rd = nn_c.nn_argmin(feats[0], feats[1])                   # exact CPU oracle, ~3 s
    assert np.array_equal(idx_st[0].cpu().numpy(), ri) and np.array_equal(d_st[0].cpu().numpy(), rd)
    # properties: the exact SIMT path agrees bit-for-bit; mutual flags are symmetric under swapping the pair
    assert torch.equal(idx_ts, cabi.nn_argmin(f, f, pairs.flip(1).contiguous(), algo=cabi.NN_EXACT_SIMT))
    m_st, xs = cabi.mutual_xs(x, pairs, idx_st, idx_ts, cabi.MUTUAL_INDEX, 0.05)
    m_ts, _ = cabi.mutual_xs(x, pairs.flip(1).contiguous(), idx_ts, idx_st, cabi.MUTUAL_INDEX, 0.05)
    assert int(m_st.sum()) == int(m_ts.sum()) > 0
    # filtering network + Kabsch at N = 50,000 vs the fp64 oracle
    sd = O.synth_state_dict(50)
    net = load_oanet(sd, gemm_algo=1)
    with torch.no_grad():
        out = net({"xs": xs})
    o64 = O.oanet_forward(xs.cpu().numpy(), sd, dtype=np.float64)
    o32 = O.oanet_forward(xs.cpu().numpy(), sd, dtype=np.float32)
    # At N = 50,000 the fp32 evaluation itself (numpy restatement of the reference) is 3e-4 away from fp64 in the logits, 3x
    # more than at N = 2000.  The tensor path cuts the 50,000-term pooling reduction into segments that are summed with
    # round-to-nearest (tcgen05 accumulates with round-toward-zero: -1e-4 relative bias otherwise) and lands at 1.3e-4
    # (measured on B200), i.e. closer to fp64 than the fp32 restatement; the gate is 2x the restatement's own distance.
    e_l = np.abs(o32["logits"][-1] - o64["logits"][-1]).max()
    e_r = O.chordal_angle(o32["rot_est"][-1], o64["rot_est"][-1]).max()
    e_t = np.abs(o32["trans_est"][-1] - o64["trans_est"][-1]).max()
    assert np.abs(out["logits"][-1].cpu().numpy() - o64["logits"][-1]).max() < 2 * e_l + 1e-4
    assert O.chordal_angle(out["rot_est"][-1].cpu().numpy(), o64["rot_est"][-1]).max() < 2 * e_r + 2e-4
    assert np.abs(out["trans_est"][-1].cpu().numpy() - o64["trans_est"][-1]).max() < 2 * e_t + 5e-4
    w = out["scores"][-1].cpu().numpy()
    Ro, to, reso, _ = O.kabsch(xs[:, 0, :, :3].cpu().numpy(), xs[:, 0, :, 3:].cpu().numpy(), w, dtype=np.float64)
    assert O.chordal_angle(out["rot_est"][-1].cpu().numpy(), Ro).max() < 1e-5
    assert np.abs(out["trans_est"][-1].cpu().numpy() - to).max() < 1e-5


def test_full_scene_properties():
    """configs[1]-sized inputs (5000 keypoints) through size-independent properties: swapping a pair swaps the two index
    sets; registering a pair alone or inside a larger job list gives identical records; det(R) = +1."""
    import importlib
    scene = importlib.import_module("3d_multiview_reg_b200.scene")
    S, n = 6, 5000
    feats, xyz, _ = O.synth_scene(S, n, seed=61)
    f, x = cu(feats), cu(xyz)
    pairs = torch.from_numpy(O.enumerate_pairs(S)).cuda()
    a = cabi.nn_argmin(f, f, pairs, algo=cabi.NN_TENSOR)
    b = cabi.nn_argmin(f, f, pairs.flip(1).contiguous(), algo=cabi.NN_TENSOR)
    ab = cabi.nn_argmin(f, f, torch.cat([pairs.flip(1), pairs], 0).contiguous(), algo=cabi.NN_TENSOR)
    assert torch.equal(ab[:15], b) and torch.equal(ab[15:], a)
    net = load_oanet(O.synth_state_dict(61), gemm_algo=1)
    reg = scene.SceneRegistrar(net, nn_algo=cabi.NN_TENSOR)
    rec_all = reg.register_scene(f, x)
    rec_one, _ = reg.register_pairs(f, x, pairs[7:8].contiguous())
    assert torch.equal(rec_all[7:8], rec_one)
    R = rec_all[:, :9].reshape(-1, 3, 3).double().cpu().numpy()
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-5 and np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-5
