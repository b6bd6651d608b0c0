"""Mirror of the hot-path subset of the reference's lib/utils.py.  Same names, argument meaning, tensor layouts
and error behaviour; the arithmetic runs in liblmpcr_b200.so (sm_100a).  Inputs must live on a CUDA device --
there is no CPU fallback."""
from itertools import combinations

import torch
import yaml

from .. import _cabi


def load_config(path):
    """lib/utils.py:19-33."""
    with open(path, "r") as f:
        return yaml.safe_load(f)


def pairwise_distance(src, dst, normalized_feature=False):
    """lib/utils.py:968-992.  [b,n,c], [b,m,c] -> [b,n,m].  Provided for API completeness (bit-identical to the
    reference's CPU result); the registration path itself never materialises this matrix."""
    if normalized_feature:
        raise NotImplementedError("normalized_feature=True is never used by the reference's callers (lib/layers.py:57)")
    return _cabi.pairwise_distance(src, dst)


def knn_point(k, pos1, pos2):
    """lib/utils.py:274-299 for k == 1 (its only use, lib/utils.py:840).  Returns (sq-dist [b,m,1], idx [b,m,1] int64)."""
    if k != 1:
        raise NotImplementedError("knn_point is only used with k=1 on this path")
    sq, idx = _cabi.knn3d_1(pos1, pos2)
    return sq.unsqueeze(-1), idx.long().unsqueeze(-1)


def extract_mutuals(x1, x2, x1_soft_matches, x2_soft_matches, threshold=0.05):
    """lib/utils.py:822-848 (geometric mutual-NN flag) -> [b,n] float {0,1} on the inputs' device."""
    _, idx = _cabi.knn3d_1(x2, x1_soft_matches)
    back = torch.gather(x2_soft_matches, 1, idx.long().unsqueeze(-1).expand(-1, -1, x1.shape[2]))
    dist = torch.pow(x1 - back, 2).sum(dim=2)
    return (dist < threshold ** 2).to(x1.dtype)


def extract_overlaping_pairs(xyz, feat, conectivity_info=None):
    """lib/utils.py:850-885.  Kept for callers that want per-pair copies; the scene path (scene.py) passes scan
    indices to the kernels instead and never duplicates features."""
    if conectivity_info is None:
        conectivity_info = pair_indices(xyz.shape[0], xyz.device)
    ci = conectivity_info.long()
    return (torch.index_select(xyz, 0, ci[:, 0]), torch.index_select(xyz, 0, ci[:, 1]),
            torch.index_select(feat, 0, ci[:, 0]), torch.index_select(feat, 0, ci[:, 1]))


def pair_indices(n_scans, device="cpu"):
    """itertools.combinations(range(S), 2) in lexicographic order (lib/utils.py:873-876) as an int32 [P,2] tensor."""
    pairs = list(combinations(range(int(n_scans)), 2))
    return torch.tensor(pairs, dtype=torch.int32, device=device).reshape(-1, 2)


def construct_filtering_input_data(xyz_s, xyz_t, data, overlapped_pair_tensors, dist_th=0.05, mutuals_flag=None):
    """lib/utils.py:888-932.  GT-free branch (no 'T_global_0'): ys = 0, Rs = I, ts = 0."""
    if "T_global_0" in data:
        raise NotImplementedError("ground-truth labels are a training-time feature (out of scope)")
    b, n = xyz_s.shape[0], xyz_s.shape[1]
    xs = torch.cat((xyz_s, xyz_t), dim=-1)
    if mutuals_flag is not None:
        xs = torch.cat((xs, mutuals_flag.reshape(b, n, 1).to(xs.dtype)), dim=-1)
    return {"xs": xs.unsqueeze(1), "ys": torch.zeros(b, n, 1), "ts": torch.zeros(b, 3, 1),
            "Rs": torch.eye(3).unsqueeze(0).repeat(b, 1, 1)}


def kabsch_transformation_estimation(x1, x2, weights=None, normalize_w=True, eps=1e-7, best_k=0, w_threshold=0):
    """lib/utils.py:164-237.  x1,x2 [b,n,3], weights [b,n] -> (R [b,3,3], t [b,3,1], res [b,n], flag).
    `flag` is True when any pair's covariance has rank < 2 (identity pose returned for it), the analogue of the
    reference's swallowed SVD failure (lib/utils.py:214-223)."""
    if not normalize_w or eps != 1e-7 or best_k != 0 or w_threshold != 0:
        raise NotImplementedError("only the configuration used by the reference's callers is built "
                                  "(normalize_w=True, eps=1e-7, best_k=0, w_threshold=0)")
    if weights is None:
        weights = torch.ones(x1.shape[0], x1.shape[1], dtype=x1.dtype, device=x1.device)
    R, t, res, status = _cabi.kabsch_points(x1, x2, weights)
    return R, t, res, bool((status & _cabi.STATUS_DEGENERATE).any().item())


def transformation_residuals(x1, x2, R, t):
    """lib/utils.py:240-256."""
    return _cabi.residuals(x1, x2, R, t)


def compute_overlap_ratio(pc_i, pc_j, trans, method="3DMatch", voxel_size=0.025):
    """lib/utils.py:713-786: max over both directions of the fraction of points that have a point of the other cloud,
    moved by the estimated pose, within 5 cm ('3DMatch') or within 3 voxels after voxel down-sampling ('FCGF').
    pc_i, pc_j: [N,3] / [M,3] numpy arrays or tensors (any device; computed on the GPU in fp64 like the reference's numpy
    arrays), trans: [4,4].  The reference's KD-tree queries become a hash-grid kernel (lmpcr_overlap_count); the Open3D
    voxel_down_sample of the 'FCGF' method becomes lmpcr_voxel_downsample."""
    import numpy as np
    dev = torch.device("cuda", torch.cuda.current_device())
    as_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(device=dev, dtype=torch.float64)
    pi, pj, T = as_dev(pc_i).reshape(-1, 3), as_dev(pc_j).reshape(-1, 3), as_dev(trans).reshape(4, 4)
    T_inv = torch.linalg.inv(T)                      # 4x4 on the device: plumbing, like np.linalg.inv in the reference (:740)
    if method == "3DMatch":
        radius = 0.05
    elif method == "FCGF":
        pi, pj = _cabi.voxel_downsample(pi, voxel_size), _cabi.voxel_downsample(pj, voxel_size)
        radius = 3 * voxel_size
    else:
        raise ValueError("Wrong overlap computation method was selected.")     # the reference logs an error and then fails (:781)
    matching01 = _cabi.overlap_count(pi, pj, T, radius)          # pc_i against trans * pc_j          (:746-748)
    matching10 = _cabi.overlap_count(pj, pi, T_inv, radius)      # pc_j against trans^-1 * pc_i       (:750-752)
    return max(matching01 / pi.shape[0], matching10 / pj.shape[0])
