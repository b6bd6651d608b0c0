"""Seeded synthetic inputs for the pairwise-registration path: scenes of FCGF-like unit features with
planted overlaps, filter-only correspondence sets with a planted rigid motion, and random weights for the filtering
network in the reference's state_dict schema.

Neutral data generator: numpy only (PCG64, bit-identical on every machine); imports neither the CUDA package nor
oracle/.  bench.py, __graft_entry__.smoke(), tools/ and the tests all draw their inputs from here, so the measured
path and its checker see the same bytes.
"""
import math

import numpy as np


def random_rotation(rng):
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(0, math.pi)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def synth_scene(n_scans, n_pts, dim=32, overlap=0.3, sigma=0.3, seed=41):
    """Unit-norm `dim`-d features + 2.5 cm-grid coordinates; every scan shares `overlap`*n_pts world points
    (noise-perturbed features, 1 cm coordinate noise) with a common pool; per-scan random rigid pose."""
    rng = np.random.default_rng(seed)
    n_ov = int(round(overlap * n_pts))
    pool_xyz = np.round(rng.uniform(0, 3, size=(n_ov, 3)) / 0.025) * 0.025
    pool_f = rng.standard_normal((n_ov, dim))
    pool_f /= np.linalg.norm(pool_f, axis=1, keepdims=True)
    feats = np.empty((n_scans, n_pts, dim), np.float32)
    xyz = np.empty((n_scans, n_pts, 3), np.float32)
    poses = []
    for s in range(n_scans):
        R, t = random_rotation(rng), rng.standard_normal(3)
        poses.append((R, t))
        own_xyz = np.round(rng.uniform(0, 3, size=(n_pts - n_ov, 3)) / 0.025) * 0.025
        own_f = rng.standard_normal((n_pts - n_ov, dim))
        f = np.concatenate([pool_f + sigma * rng.standard_normal((n_ov, dim)) / math.sqrt(dim), own_f], axis=0)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        world = np.concatenate([pool_xyz + 0.01 * rng.standard_normal((n_ov, 3)), own_xyz], axis=0)
        local = (world - t) @ R                                      # world = R local + t
        perm = rng.permutation(n_pts)
        feats[s] = f[perm].astype(np.float32)
        xyz[s] = local[perm].astype(np.float32)
    return feats, xyz, poses


def synth_xs(n_pairs, n_pts, inlier_frac=0.3, seed=41, noise=0.01):
    """Filter-only inputs xs [P,1,N,6] with a planted rigid motion on `inlier_frac` of the correspondences."""
    rng = np.random.default_rng(seed)
    xs = np.empty((n_pairs, 1, n_pts, 6), np.float32)
    Rs = np.empty((n_pairs, 3, 3))
    ts = np.empty((n_pairs, 3))
    for p in range(n_pairs):
        R, t = random_rotation(rng), rng.standard_normal(3)
        x1 = rng.uniform(0, 3, size=(n_pts, 3))
        x2 = x1 @ R.T + t + noise * rng.standard_normal((n_pts, 3))
        n_out = n_pts - int(round(inlier_frac * n_pts))
        out_idx = rng.permutation(n_pts)[:n_out]
        x2[out_idx] = rng.uniform(-1, 4, size=(n_out, 3))
        xs[p, 0, :, :3] = x1
        xs[p, 0, :, 3:] = x2
        Rs[p], ts[p] = R, t
    return xs, Rs, ts


def synth_cloud_pair(seed, n_i, n_j, overlap):
    """Two partially overlapping clouds in their own frames + a slightly wrong 4x4 estimate of the pose that maps cloud j
    into the frame of cloud i (inputs of the overlap-ratio check, lib/utils.py:713).  fp64."""
    rng = np.random.default_rng(seed)
    n_ov = int(overlap * min(n_i, n_j))
    shared = rng.uniform(0, 2, (n_ov, 3))
    wi = np.concatenate([shared + 0.004 * rng.standard_normal((n_ov, 3)), rng.uniform(0, 2, (n_i - n_ov, 3)) + [2.2, 0, 0]])
    wj = np.concatenate([shared + 0.004 * rng.standard_normal((n_ov, 3)), rng.uniform(0, 2, (n_j - n_ov, 3)) - [2.2, 0, 0]])
    R, t = random_rotation(rng), rng.standard_normal(3)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t                                                     # x_i = R x_j + t
    pj = (wj - t) @ R                                                # j's own frame
    Tn = T.copy()
    Tn[:3, 3] += 0.01 * rng.standard_normal(3)
    return wi[rng.permutation(n_i)], pj[rng.permutation(n_j)], Tn


# --------------------------------------------------------------------------------------------------
# Parameter schema of the filtering network + seeded synthetic weights
# --------------------------------------------------------------------------------------------------


def oanet_param_schema(net_channel=128, clusters=500, net_depth=12, iter_num=1, side_channel=0):
    """[(name, shape)] in the order of OANet(cfg).state_dict() (oanet.py:133-163,199-215), including the
    BatchNorm buffers (`num_batches_tracked` has shape ())."""
    C, K = net_channel, clusters
    half = (net_depth // (iter_num + 1)) // 2
    out = []

    def conv(p, co, ci):
        out.append((p + ".weight", (co, ci, 1, 1)))
        out.append((p + ".bias", (co,)))

    def bn(p, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            out.append((p + "." + n, (c,)))
        out.append((p + ".num_batches_tracked", ()))

    def pointcn(p, ci, co):
        if ci != co:
            conv(p + ".shot_cut", co, ci)
        bn(p + ".conv.1", ci)
        conv(p + ".conv.3", co, ci)
        bn(p + ".conv.5", co)
        conv(p + ".conv.7", co, co)

    def block(p, cin):
        conv(p + ".conv1", C, cin)
        bn(p + ".down1.conv.1", C)
        conv(p + ".down1.conv.3", K, C)
        bn(p + ".up1.conv.1", C)
        conv(p + ".up1.conv.3", K, C)
        for i in range(half):
            pointcn(p + ".l1_1.%d" % i, C, C)
        pointcn(p + ".l1_2.0", 2 * C, C)
        for i in range(1, half):
            pointcn(p + ".l1_2.%d" % i, C, C)
        for i in range(half):
            q = p + ".l2.%d" % i
            bn(q + ".conv1.1", C)
            conv(q + ".conv1.3", C, C)
            bn(q + ".conv2.0", K)
            conv(q + ".conv2.2", K, K)
            bn(q + ".conv3.2", C)
            conv(q + ".conv3.4", C, C)
        conv(p + ".output", 1, C)

    block("reg_init", 6 + side_channel)
    for i in range(iter_num):
        block("reg_iter.%d" % i, 8 + side_channel)
    return out


def synth_state_dict(seed=41, **cfg):
    """Seeded weights for the filtering network (numpy PCG64: identical on every machine).  Conv weights
    ~ U(+-1/sqrt(fan_in)) like torch's default init; BatchNorm affine AND running statistics are
    randomised (default init leaves running_mean=0 / running_var=1, which would hide BN-folding bugs)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in oanet_param_schema(**cfg):
        leaf = name.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[name] = np.array(0, np.int64)
        elif len(shape) == 4:
            b = 1.0 / math.sqrt(shape[1])
            sd[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif leaf == "running_var":
            sd[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "running_mean":
            sd[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif leaf == "weight":                                   # BN gamma
            sd[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        else:                                                     # conv bias / BN beta
            sd[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
    return sd
