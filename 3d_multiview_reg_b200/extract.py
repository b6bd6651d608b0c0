"""Correspondence extraction in the on-disk format of the reference's training data (the caller on the
input side of the hot path):  scripts/extract_data.py:120-201 `run_correspondence_extraction` loops over the fragment
pairs of a scene on the CPU (sklearn brute-force 2-NN in both directions) and writes one `.npz` per pair with

    x       [n,6]  = [xyz_1[nn_21[:,0]], xyz_2]          row j belongs to point j of the SECOND fragment (:194)
    mutuals [n,1]  float64, 1 where nn_12[nn_21[j,0],0] == j                                             (:186-190)
    ratios  [n]    float64, d1/d2 of the 1->2 search (Lowe ratio; indexed by points of the FIRST fragment) (:191)

Here all pairs of a scene go through `lmpcr_nn_top2_algo` (tcgen05 screening + exact fp32 rescoring for 32-d features, bit-identical to the exact
fp32 CUDA-core kernel `lmpcr_nn_top2`) in batched launches and the three arrays
are assembled on the GPU.  File naming and keys follow the reference so that its PrecomputedPairwiseDataset-style readers
(lib/data.py) consume the files unchanged.  The random subsampling of :160-168 stays with the caller (pass the rows you
want); see `sample_indices` for the same with/without-replacement rule.
"""
import os

import numpy as np
import torch

from . import _cabi
from .lib.utils import pair_indices


def sample_indices(n_points, n_correspondences, rng=np.random):
    """scripts/extract_data.py:160-168: without replacement when the cloud is large enough, else with replacement."""
    return rng.choice(n_points, n_correspondences, replace=n_points < n_correspondences)


def scene_correspondences(feats, xyz, pairs=None, pair_chunk=512):
    """feats [S,n,32], xyz [S,n,3] CUDA fp32; pairs [P,2] int32 (default: all idx_1 < idx_2, :146-147).
    Returns dict(x [P,n,6] fp32, mutuals [P,n,1] fp64, ratios [P,n] fp64, nn_12 [P,n,2], nn_21 [P,n,2] int32)."""
    dev = feats.device
    if pairs is None:
        pairs = pair_indices(feats.shape[0], dev)
    pairs = pairs.to(device=dev, dtype=torch.int32).contiguous()
    P, n = pairs.shape[0], feats.shape[1]
    out = {"x": torch.empty((P, n, 6), dtype=torch.float32, device=dev), "mutuals": torch.empty((P, n, 1), dtype=torch.float64, device=dev),
           "ratios": torch.empty((P, n), dtype=torch.float64, device=dev), "nn_12": torch.empty((P, n, 2), dtype=torch.int32, device=dev),
           "nn_21": torch.empty((P, n, 2), dtype=torch.int32, device=dev)}
    ar = torch.arange(n, device=dev)
    for p0 in range(0, P, pair_chunk):
        pc = pairs[p0:p0 + pair_chunk]
        nc = pc.shape[0]
        idx, dist = _cabi.nn_top2(feats, feats, torch.cat([pc, pc.flip(1)], 0).contiguous())
        nn12, nn21, d12 = idx[:nc], idx[nc:], dist[:nc]
        first21 = nn21[:, :, 0].long()
        back = torch.gather(nn12[:, :, 0].long(), 1, first21)                      # nn_12[nn_21[j,0],0]
        out["mutuals"][p0:p0 + nc, :, 0] = (back == ar[None]).to(torch.float64)
        dd = d12.clamp_min(0).to(torch.float64).sqrt()
        out["ratios"][p0:p0 + nc] = dd[:, :, 0] / dd[:, :, 1]
        x1 = _cabi.gather_xyz(xyz, pc.flip(1).contiguous(), nn21[:, :, 0].contiguous())      # xyz_1[nn_21[:,0]]
        out["x"][p0:p0 + nc, :, :3] = x1
        out["x"][p0:p0 + nc, :, 3:] = xyz[pc[:, 1].long()]
        out["nn_12"][p0:p0 + nc] = nn12
        out["nn_21"][p0:p0 + nc] = nn21
    return out


def save_scene_correspondences(target_dir, scene_name, corr, pairs, skip_existing=True):
    """Writes `<scene>_<iii>_<jjj>.npz` with keys x / mutuals / ratios exactly as scripts/extract_data.py:196-201."""
    os.makedirs(target_dir, mode=0o755, exist_ok=True)
    x, mut, rat = corr["x"].cpu().numpy(), corr["mutuals"].cpu().numpy(), corr["ratios"].cpu().numpy()
    written = []
    for p, (a, b) in enumerate(pairs.cpu().tolist()):
        path = os.path.join(target_dir, "{}_{}_{}.npz".format(scene_name, str(a).zfill(3), str(b).zfill(3)))
        if skip_existing and os.path.exists(path):
            continue
        np.savez_compressed(path, x=x[p], mutuals=mut[p], ratios=rat[p])
        written.append(path)
    return written
