"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's pairwise-registration hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  The product (3d_multiview_reg_b200) never imports anything from oracle/.

Parity status: PINNED by execution of the unmodified reference.  The reference has no tests, golden
vectors or fixtures of its own, so this restatement is pinned against outputs of the
reference itself, generated in the build container by tests/golden/make_golden.py (which imports
/root/reference through oracle/refimport.py) and committed under tests/golden/*.npz;
tests/test_oracle_golden.py re-checks the restatement against those files on every run.

Every function cites the reference file:line (relative to the reference root) that it follows.
Arithmetic is carried out in `dtype` (float32 = the reference's precision; float64 = truth witness).
"""
import itertools
import math

import numpy as np

# --------------------------------------------------------------------------------------------------
# Stage 1: feature-space nearest neighbours  (lib/utils.py:968-992, lib/layers.py:57,81-86)
# --------------------------------------------------------------------------------------------------


def sqnorm_rows_f32(f):
    """torch.sum(f ** 2, dim=-1) in float32 as the reference evaluates it (lib/utils.py:989-990).

    Summation order probed against torch 2.11 CPU (see DESIGN.md "fp32 evaluation order"): eight lane
    accumulators t[l] = ((q[l] + q[8+l]) + q[16+l]) + ... over chunks of 8, then t[0] + t[1] + ... + t[7]
    sequentially, q = fl(f*f).  Bit-exact versus torch for D = 32 (the FCGF dimension).
    """
    f = np.ascontiguousarray(f, dtype=np.float32)
    n, d = f.shape
    assert d % 8 == 0
    q = (f * f).astype(np.float32)
    t = q[:, 0:8].copy()
    for c in range(1, d // 8):
        t = (t + q[:, 8 * c:8 * c + 8]).astype(np.float32)
    s = t[:, 0].copy()
    for l in range(1, 8):
        s = (s + t[:, l]).astype(np.float32)
    return s


def _dot_seq_fma_f32(a, b):
    """fp32 a @ b.T with a sequential fused-multiply-add chain over k (k = 0..D-1, accumulator starts at
    0), which is what torch.matmul (MKL sgemm, K = 32) produces bit-for-bit (lib/utils.py:984; probed).
    float64 emulation of fmaf: the product of two fp32 numbers is exact in fp64; the fp64 add followed
    by the fp32 rounding can differ from a true fma only on double-rounding ties (probability ~2^-29
    per op); oracle/nn_oracle.c holds the exact fmaf version used for large cases."""
    a64 = np.asarray(a, dtype=np.float64)
    b64 = np.asarray(b, dtype=np.float64)
    acc = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    for k in range(a.shape[1]):
        acc = (a64[:, k:k + 1] * b64[None, :, k] + acc.astype(np.float64)).astype(np.float32)
    return acc


def pairwise_distance_f32(src, dst):
    """lib/utils.py:968-992 with normalized_feature=False (the only mode its callers use,
    lib/layers.py:57):  dist = -(src @ dst^T); dist = 2*dist; dist += |src|^2[:,None]; dist += |dst|^2[None,:]
    each step rounded to fp32."""
    c = _dot_seq_fma_f32(src, dst)
    dist = (np.float32(2.0) * (-c)).astype(np.float32)
    dist = (dist + sqnorm_rows_f32(src)[:, None]).astype(np.float32)
    dist = (dist + sqnorm_rows_f32(dst)[None, :]).astype(np.float32)
    return dist


def nn_argmin_f32(src, dst):
    """Hard nearest neighbour of every src row among dst rows: lib/layers.py:81 `dist.min(dim=2)[1]`
    (first minimum wins ties, probed on torch CPU).  Returns (idx int32 [n], dist fp32 [n])."""
    dist = pairwise_distance_f32(src, dst)
    idx = np.argmin(dist, axis=1).astype(np.int32)
    return idx, dist[np.arange(dist.shape[0]), idx]


def nn_top2_f32(src, dst):
    """Two nearest neighbours per src row, scripts/extract_data.py:176-184 (`NearestNeighbors.kneighbors(n_neighbors=2)`,
    sklearn brute force, Euclidean).  sklearn 1.x evaluates float32 inputs with a chunked GEMM upcast to float64; the
    restatement keeps the torch fp32 distance used everywhere else on the path (first minimum, then the smallest of the
    rest) and is pinned against sklearn's own output to 1e-4 in the ratio (tests/golden/extract_golden.npz).
    Returns (idx int32 [n,2], squared fp32 distances [n,2])."""
    dist = pairwise_distance_f32(src, dst)
    r = np.arange(dist.shape[0])
    i1 = np.argmin(dist, axis=1)
    d1 = dist[r, i1].copy()
    dist[r, i1] = np.inf
    i2 = np.argmin(dist, axis=1)
    d2 = dist[r, i2]
    return np.stack([i1, i2], 1).astype(np.int32), np.stack([d1, d2], 1)


def extract_correspondences(f1, k1, f2, k2):
    """The arrays scripts/extract_data.py:176-201 stores per fragment pair (`x`, `mutuals`, `ratios`):
      nn  = 2-NN of scan-1 features among scan-2 features,  nn1 = 2-NN of scan-2 features among scan-1 features,
      x       = [k1[nn1[:,0]], k2]                               (row j belongs to point j of scan 2, :194)
      mutuals = 1 where nn[nn1[j,0],0] == j                     (:186-190, float64 [n,1])
      ratios  = nn_dists[:,0] / nn_dists[:,1]                   (:191 -- indexed by scan-1 points, as in the reference)"""
    nn, d = nn_top2_f32(f1, f2)
    nn1, _ = nn_top2_f32(f2, f1)
    n = f2.shape[0]
    mutuals = np.zeros((n, 1))
    mutuals[nn[nn1[:, 0], 0] == np.arange(n)] = 1
    dd = np.sqrt(np.maximum(d.astype(np.float64), 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        ratios = dd[:, 0] / dd[:, 1]
    x = np.concatenate([np.asarray(k1)[nn1[:, 0]], np.asarray(k2)], axis=1)
    return x, mutuals, ratios, nn, nn1


def hard_correspondences(x_f, y_f, y_c):
    """Soft_NN(corr_type='hard').forward (lib/layers.py:44-88): one-hot(argmin) @ y_c == y_c[argmin]."""
    idx, _ = nn_argmin_f32(x_f, y_f)
    return np.asarray(y_c, dtype=np.float32)[idx], idx


def soft_correspondences(x_f, y_f, y_c, temperature):
    """Soft_NN(corr_type='soft', st=False).forward (lib/layers.py:59-70,86): softmax(-dist/T, dim=2) @ y_c, with dist the
    reference's fp32 distance; the softmax and the blend are carried in fp64 here (truth witness)."""
    dist = pairwise_distance_f32(x_f, y_f).astype(np.float64)
    l = -dist / float(temperature)
    l -= l.max(axis=1, keepdims=True)
    w = np.exp(l)
    w /= w.sum(axis=1, keepdims=True)
    return w @ np.asarray(y_c, np.float64)


def mutual_index(idx_st, idx_ts):
    """Index definition of a mutual nearest neighbour, scripts/extract_data.py:186-190, expressed for
    source points:  i is mutual iff idx_ts[idx_st[i]] == i."""
    idx_st = np.asarray(idx_st)
    idx_ts = np.asarray(idx_ts)
    return (idx_ts[idx_st] == np.arange(idx_st.shape[0])).astype(np.uint8)


def knn_point_1(pos1, pos2):
    """lib/utils.py:274-299 with k=1: for every row of pos2 the index of the closest row of pos1 under
    sum(-(p1-p2)^2) (fp32, direct differences), torch.topk -> first maximum."""
    p1 = np.asarray(pos1, dtype=np.float32)
    p2 = np.asarray(pos2, dtype=np.float32)
    out = np.empty(p2.shape[0], dtype=np.int64)
    for s in range(0, p2.shape[0], 1024):
        d = p1[None, :, :] - p2[s:s + 1024, None, :]
        d = (d * d).astype(np.float32)
        dist = -((d[..., 0] + d[..., 1]).astype(np.float32) + d[..., 2]).astype(np.float32)
        out[s:s + 1024] = np.argmax(dist, axis=1)
    return out


def extract_mutuals(x1, x2, x1_soft_matches, x2_soft_matches, threshold=0.05):
    """Geometric mutual test, lib/utils.py:822-848 (single pair, no batch axis)."""
    idx = knn_point_1(x2, x1_soft_matches)
    delta = np.asarray(x1, np.float32) - np.asarray(x2_soft_matches, np.float32)[idx]
    d2 = (delta * delta).astype(np.float32)
    dist = ((d2[:, 0] + d2[:, 1]).astype(np.float32) + d2[:, 2]).astype(np.float32)
    return (dist < np.float32(threshold ** 2)).astype(np.float32)


def enumerate_pairs(n_scans):
    """lib/utils.py:873-876: itertools.combinations(range(S), 2), lexicographic."""
    return np.array(list(itertools.combinations(range(n_scans), 2)), dtype=np.int32).reshape(-1, 2)


def construct_xs(xyz_s, xyz_t_corr, mutuals=None):
    """lib/utils.py:915-926: xs = cat(xyz_s, xyz_t_corr [, mutuals]) -> [1, n, 6(7)] (batch axis added by caller)."""
    xs = np.concatenate([np.asarray(xyz_s, np.float32), np.asarray(xyz_t_corr, np.float32)], axis=-1)
    if mutuals is not None:
        xs = np.concatenate([xs, np.asarray(mutuals, np.float32).reshape(-1, 1)], axis=-1)
    return xs[None]


def register_pair_stage1(feat_s, feat_t, xyz_s, xyz_t, mutual_mode="index", mutual_thresh=0.05):
    """Stage 1 for one scan pair, hard NN both ways (lib/pairwise/__init__.py:110-120)."""
    idx_st, _ = nn_argmin_f32(feat_s, feat_t)
    idx_ts, _ = nn_argmin_f32(feat_t, feat_s)
    xyz_s = np.asarray(xyz_s, np.float32)
    xyz_t = np.asarray(xyz_t, np.float32)
    if mutual_mode == "index":
        mutual = mutual_index(idx_st, idx_ts)
    else:  # geometric definition evaluated through the index chase (hard NN: soft match == a target point)
        back = xyz_s[idx_ts[idx_st]]
        delta = xyz_s - back
        d2 = (delta * delta).astype(np.float32)
        dist = ((d2[:, 0] + d2[:, 1]).astype(np.float32) + d2[:, 2]).astype(np.float32)
        mutual = (dist < np.float32(mutual_thresh ** 2)).astype(np.uint8)
    xs = construct_xs(xyz_s, xyz_t[idx_st])
    return idx_st, idx_ts, mutual, xs


# --------------------------------------------------------------------------------------------------
# Stage 3: weighted Kabsch + residuals  (lib/utils.py:164-256)
# --------------------------------------------------------------------------------------------------


def transformation_residuals(x1, x2, R, t, dtype=np.float32):
    """lib/utils.py:240-256:  || (R x1^T + t)^T - x2 ||_2 per correspondence.  x1,x2 [P,n,3], R [P,3,3], t [P,3,1]."""
    x1 = np.asarray(x1, dtype)
    x2 = np.asarray(x2, dtype)
    rec = np.matmul(np.asarray(R, dtype), x1.transpose(0, 2, 1)) + np.asarray(t, dtype)
    return np.linalg.norm(rec.transpose(0, 2, 1) - x2, axis=2).astype(dtype)


def kabsch(x1, x2, weights, eps=1e-7, dtype=np.float32):
    """lib/utils.py:164-237 (normalize_w=True, best_k=0, w_threshold=0 -- the only mode its callers use).
    Returns R [P,3,3], t [P,3,1], res [P,n], flag."""
    x1 = np.asarray(x1, dtype)
    x2 = np.asarray(x2, dtype)
    w = np.asarray(weights, dtype)
    eps = dtype(eps)
    w = w / (w.sum(axis=1, keepdims=True) + eps)              # :187-189
    w = w[:, :, None]
    wsum = w.sum(axis=1)[:, None] + eps                       # :203-204
    x1_mean = np.matmul(w.transpose(0, 2, 1), x1) / wsum
    x2_mean = np.matmul(w.transpose(0, 2, 1), x2) / wsum
    x1c = x1 - x1_mean
    x2c = x2 - x2_mean
    cov = np.matmul(x1c.transpose(0, 2, 1), w * x2c)          # :209-212 (diag_embed matmul)
    try:
        u, s, vh = np.linalg.svd(cov)                         # :215
    except np.linalg.LinAlgError:                             # :216-223
        P = x1.shape[0]
        R = np.tile(np.eye(3, dtype=dtype), (P, 1, 1))
        t = np.zeros((P, 3, 1), dtype)
        return R, t, transformation_residuals(x1, x2, R, t, dtype), True
    v = vh.transpose(0, 2, 1)
    det = np.linalg.det(np.matmul(vh, u.transpose(0, 2, 1)))  # :225  det(V^T U^T)
    D = np.tile(np.eye(3, dtype=dtype), (x1.shape[0], 1, 1))
    D[:, 2, 2] = det.astype(dtype)
    R = np.matmul(v, np.matmul(D, u.transpose(0, 2, 1))).astype(dtype)       # :229
    t = (x2_mean.transpose(0, 2, 1) - np.matmul(R, x1_mean.transpose(0, 2, 1))).astype(dtype)  # :232
    return R, t, transformation_residuals(x1, x2, R, t, dtype), False


def pair_confidence(weights, residuals, eps=1e-7, inlier_w=0.5, inlier_res=0.05):
    """Per-pair confidence record.  The released reference code has no per-pair confidence (SURVEY.md
    8a row a15); the build defines it from the quantities its callers threshold:
    [ #(w > 0.5)  (scripts/benchmark_pairwise_registration.py:211),  sum(w),
      sqrt(sum(w_norm * res^2))  with w_norm as in lib/utils.py:187-189,
      #(res < 5 cm)  (lib/utils.py:888 dist_th) ]."""
    w = np.asarray(weights, np.float64)
    r = np.asarray(residuals, np.float64)
    wn = w / (w.sum(axis=1, keepdims=True) + eps)
    return np.stack([(w > inlier_w).sum(axis=1).astype(np.float64), w.sum(axis=1),
                     np.sqrt((wn * r * r).sum(axis=1)), (r < inlier_res).sum(axis=1).astype(np.float64)], axis=1)


def chordal_angle(Ra, Rb):
    """2*asin(|Ra-Rb|_F / (2*sqrt(2))) in fp64 -- the rotation metric used for every pose gate (the
    reference's acos((tr-1)/2), lib/utils.py:135-143, has a 5e-4 rad noise floor in fp32)."""
    d = np.linalg.norm(np.asarray(Ra, np.float64) - np.asarray(Rb, np.float64), axis=(-2, -1))
    return 2.0 * np.arcsin(np.clip(d / (2.0 * math.sqrt(2.0)), 0.0, 1.0))


# --------------------------------------------------------------------------------------------------
# Stage 2: correspondence-weighting network  (lib/filtering/oanet.py)
# --------------------------------------------------------------------------------------------------


class _SD:
    """state_dict view with a key prefix; values as numpy arrays of the working dtype."""

    def __init__(self, sd, prefix, dtype, train=False, updates=None):
        self.sd, self.prefix, self.dtype = sd, prefix, dtype
        self.train, self.updates = train, updates          # train: BatchNorm uses batch statistics and records its buffer updates

    def sub(self, name):
        return _SD(self.sd, self.prefix + name + ".", self.dtype, self.train, self.updates)

    def __getitem__(self, name):
        v = self.sd[self.prefix + name]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.asarray(v, self.dtype)

    def has(self, name):
        return (self.prefix + name) in self.sd


def _inorm(x, eps):
    """nn.InstanceNorm2d(affine=False, track_running_stats=False): per (pair, channel) over the point axis,
    biased variance (oanet.py:27,31,65,79,101,118).  x [P,C,L]."""
    mean = x.mean(axis=2, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=2, keepdims=True)
    return (x - mean) / np.sqrt(var + x.dtype.type(eps))


def _bn_eval(x, sd, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d, channel axis = 1 (oanet.py:28,32,66,73,80,102,119).  Eval mode: running statistics.  Train mode
    (sd.train; the state scripts/benchmark_pairwise_registration.py leaves the model in): statistics of the
    batch -- biased variance over (pairs, points) for the normalisation, unbiased for the running_var update, momentum 0.1
    (oanet.py:15), num_batches_tracked + 1; the updated buffers are recorded in sd.updates."""
    if sd.train:
        mean = x.mean(axis=(0, 2))
        var = ((x - mean[None, :, None]) ** 2).mean(axis=(0, 2))
        n = x.shape[0] * x.shape[2]
        m = x.dtype.type(momentum)
        sd.updates[sd.prefix + "running_mean"] = (1 - m) * sd["running_mean"] + m * mean
        sd.updates[sd.prefix + "running_var"] = (1 - m) * sd["running_var"] + m * var * x.dtype.type(n / max(n - 1, 1))
        sd.updates[sd.prefix + "num_batches_tracked"] = np.asarray(sd.sd[sd.prefix + "num_batches_tracked"]).astype(np.int64) + 1
        s = sd["weight"] / np.sqrt(var + x.dtype.type(eps))
        return (x - mean[None, :, None]) * s[None, :, None] + sd["bias"][None, :, None]
    s = sd["weight"] / np.sqrt(sd["running_var"] + x.dtype.type(eps))
    return (x - sd["running_mean"][None, :, None]) * s[None, :, None] + sd["bias"][None, :, None]


def _conv1x1(x, sd):
    """nn.Conv2d(kernel_size=1): W [Co,Ci,1,1] applied per point."""
    W = sd["weight"].reshape(sd["weight"].shape[0], -1)
    return np.einsum("oc,pcl->pol", W, x, optimize=True).astype(x.dtype) + sd["bias"][None, :, None]


def _relu(x):
    return np.maximum(x, 0)


def _pointcn(x, sd):
    """PointCN, oanet.py:18-43."""
    h = _relu(_bn_eval(_inorm(x, 1e-5), sd.sub("conv.1")))
    h = _conv1x1(h, sd.sub("conv.3"))
    h = _relu(_bn_eval(_inorm(h, 1e-5), sd.sub("conv.5")))
    h = _conv1x1(h, sd.sub("conv.7"))
    if sd.has("shot_cut.weight"):
        return h + _conv1x1(x, sd.sub("shot_cut"))
    return h + x


def _oafilter(x, sd):
    """OAFilter, oanet.py:56-93 (x [P,C,K])."""
    out = _conv1x1(_relu(_bn_eval(_inorm(x, 1e-3), sd.sub("conv1.1"))), sd.sub("conv1.3"))
    out = out.transpose(0, 2, 1)                                   # trans(1,2): [P,K,C]
    out = out + _conv1x1(_relu(_bn_eval(out, sd.sub("conv2.0"))), sd.sub("conv2.2"))
    out = out.transpose(0, 2, 1)
    out = _conv1x1(_relu(_bn_eval(_inorm(out, 1e-3), sd.sub("conv3.2"))), sd.sub("conv3.4"))
    return out + x


def _softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def _diff_pool(x, sd):
    """diff_pool, oanet.py:96-110: softmax over the POINT axis."""
    embed = _conv1x1(_relu(_bn_eval(_inorm(x, 1e-3), sd.sub("conv.1"))), sd.sub("conv.3"))   # [P,K,N]
    S = _softmax(embed, axis=2)
    return np.matmul(x, S.transpose(0, 2, 1))                       # [P,C,K]


def _diff_unpool(x_up, x_down, sd):
    """diff_unpool, oanet.py:113-129: softmax over the CLUSTER axis."""
    embed = _conv1x1(_relu(_bn_eval(_inorm(x_up, 1e-3), sd.sub("conv.1"))), sd.sub("conv.3"))  # [P,K,N]
    S = _softmax(embed, axis=1)
    return np.matmul(x_down, S)                                     # [P,C,N]


def oan_block(data, xs, sd, depth, dtype, guard="batch"):
    """OANBlock.forward, oanet.py:165-185.  data [P,Cin,N]; xs [P,1,N,>=6]."""
    half = depth // 2
    x = _conv1x1(data, sd.sub("conv1"))
    for i in range(half):
        x = _pointcn(x, sd.sub("l1_1.%d" % i))
    x1_1 = x
    x_down = _diff_pool(x1_1, sd.sub("down1"))
    x2 = x_down
    for i in range(half):
        x2 = _oafilter(x2, sd.sub("l2.%d" % i))
    x_up = _diff_unpool(x1_1, x2, sd.sub("up1"))
    out = np.concatenate([x1_1, x_up], axis=1)
    for i in range(half):
        out = _pointcn(out, sd.sub("l1_2.%d" % i))
    logits = _conv1x1(out, sd.sub("output"))[:, 0, :]
    weights = _relu(np.tanh(logits))
    zero = weights.sum(axis=1) == 0.0
    if guard == "batch":                                            # oanet.py:177-178 (batch-coupled, Q6)
        if np.any(zero):
            weights = weights + dtype(1.0 / weights.shape[1])
    else:                                                           # per-pair guard (scene / multi-GPU mode)
        weights = weights + zero[:, None].astype(dtype) * dtype(1.0 / weights.shape[1])
    x1, x2c = xs[:, 0, :, :3], xs[:, 0, :, 3:6]
    R, t, res, flag = kabsch(x1, x2c, weights, dtype=dtype)
    return logits, weights, R, t, res, out, flag


def oanet_forward(xs, state_dict, net_depth=12, iter_num=1, prefix="", dtype=np.float32, guard="batch", train=False):
    """OANet.forward, oanet.py:218-265.  xs [P,1,N,6(+1)].  state_dict: reference key names (optionally with `prefix`, e.g.
    'filtering_module.').  train=True: BatchNorm in training mode (batch statistics); out["bn_updates"] then maps every
    BatchNorm buffer name to its value after the forward pass (the input state_dict is not modified)."""
    xs = np.asarray(xs, dtype)
    assert xs.ndim == 4 and xs.shape[1] == 1
    depth = net_depth // (iter_num + 1)
    updates = {}
    sd = _SD(state_dict, prefix, dtype, train, updates)
    inp = xs.transpose(0, 3, 2, 1)[:, :, :, 0]                       # [P,Cx,N]
    out = {"logits": [], "scores": [], "rot_est": [], "trans_est": [], "residuals": []}
    logits, w, R, t, res, lat, flag = oan_block(inp, xs, sd.sub("reg_init"), depth, dtype, guard)
    for k, v in zip(("logits", "scores", "rot_est", "trans_est", "residuals"), (logits, w, R, t, res)):
        out[k].append(v)
    for i in range(iter_num):
        data = np.concatenate([inp, res[:, None, :], w[:, None, :]], axis=1)
        logits, w, R, t, res, lat, f2 = oan_block(data, xs, sd.sub("reg_iter.%d" % i), depth, dtype, guard)
        flag = flag or f2
        for k, v in zip(("logits", "scores", "rot_est", "trans_est", "residuals"), (logits, w, R, t, res)):
            out[k].append(v)
    out["latent features"] = lat[:, :, :, None]
    out["gradient_flag"] = flag
    if train:
        out["bn_updates"] = updates
    return out


# --------------------------------------------------------------------------------------------------
# Overlap ratio under an estimated pose  (lib/utils.py:713-786)
# --------------------------------------------------------------------------------------------------


def voxel_down_sample(pts, voxel):
    """Restates Open3D 0.9 `PointCloud::VoxelDownSample` (called at lib/utils.py:760-761; Open3D itself is absent here, so
    this part is UNPINNED): voxel index = floor((p - (min_bound - voxel/2)) / voxel), output = mean of the points of each
    occupied voxel.  Rows are returned sorted by (ix, iy, iz); Open3D's order is an unordered_map's."""
    pts = np.asarray(pts, np.float64)
    origin = pts.min(axis=0) - 0.5 * voxel
    ijk = np.floor((pts - origin) / voxel).astype(np.int64)
    order = np.lexsort((ijk[:, 2], ijk[:, 1], ijk[:, 0]))
    ijk, p = ijk[order], pts[order]
    head = np.ones(len(p), bool)
    head[1:] = np.any(ijk[1:] != ijk[:-1], axis=1)
    starts = np.flatnonzero(head)
    sums = np.add.reduceat(p, starts, axis=0)
    counts = np.diff(np.append(starts, len(p)))[:, None]
    return sums / counts


def _count_within(query, target, radius):
    """Number of query points whose nearest target point is closer than radius (brute force in fp64, blocked)."""
    n = 0
    for s in range(0, len(query), 2048):
        d2 = ((query[s:s + 2048, None, :] - target[None, :, :]) ** 2).sum(-1)
        n += int((np.sqrt(d2.min(axis=1)) < radius).sum())
    return n


def compute_overlap_ratio(pc_i, pc_j, trans, method="3DMatch", voxel_size=0.025):
    """lib/utils.py:713-786 with the KD-tree replaced by brute force (same Euclidean metric, same strict `<`)."""
    pc_i, pc_j, trans = np.asarray(pc_i, np.float64), np.asarray(pc_j, np.float64), np.asarray(trans, np.float64)
    trans_inv = np.linalg.inv(trans)
    if method == "FCGF":
        pc_i, pc_j = voxel_down_sample(pc_i, voxel_size), voxel_down_sample(pc_j, voxel_size)
        radius = 3 * voxel_size
    else:
        radius = 0.05
    pc_i_t = (trans_inv[0:3, 0:3] @ pc_i.T + trans_inv[0:3, 3].reshape(-1, 1)).T
    pc_j_t = (trans[0:3, 0:3] @ pc_j.T + trans[0:3, 3].reshape(-1, 1)).T
    m01 = _count_within(pc_i, pc_j_t, radius)
    m10 = _count_within(pc_j, pc_i_t, radius)
    return max(m01 / pc_i.shape[0], m10 / pc_j.shape[0]), m01, m10


# --------------------------------------------------------------------------------------------------
# Synthetic inputs: shared, neutral generators (synthdata.py at the repo root) re-exported for the tests
# --------------------------------------------------------------------------------------------------
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthdata import oanet_param_schema, random_rotation, synth_cloud_pair, synth_scene, synth_state_dict, synth_xs  # noqa: E402,F401


# ----------------------------------------------------------------------------- keypoint sampler (lib/layers.py:90-154)
def sampler_rand(coords, feats, pts, m, rng=np.random):
    """Restatement of Sampler('rand').forward: per cloud `choice(range, m, replace=False)` when EVERY cloud of the batch has at
    least m points, else `choice(range, m, replace=True)` (lib/layers.py:126,141-145); then the two index_selects.  `rng` is
    numpy's global stream by default, exactly what the reference consumes.  Returns (idx [b,m], coords [b,m,3], feats [b,m,c])."""
    pts = [int(v) for v in pts]
    replace = not (min(m, min(pts)) >= m)
    idx, start = [], 0
    for n in pts:
        idx.append(rng.choice(np.arange(start, start + n), m, replace=replace))
        start += n
    idx = np.stack(idx, 0)
    return idx, coords[idx], feats[idx]


def sampler_contract(idx, pts, m):
    """What any 'rand' sample must satisfy, whatever the random stream: shape [b,m], every index inside its own cloud, and no
    repeats unless the batch took the with-replacement branch.  Returns the list of violations (empty = valid)."""
    pts = [int(v) for v in pts]
    idx = np.asarray(idx)
    bad = []
    if idx.shape != (len(pts), m):
        return ["shape %s != %s" % (idx.shape, (len(pts), m))]
    replace = not (min(m, min(pts)) >= m)
    start = 0
    for s, n in enumerate(pts):
        row = idx[s]
        if row.min() < start or row.max() >= start + n:
            bad.append("cloud %d: index outside [%d, %d)" % (s, start, start + n))
        if not replace and len(np.unique(row)) != m:
            bad.append("cloud %d: repeated index without replacement" % s)
        start += n
    return bad
