"""ctypes binding of liblmpcr_b200.so (include/lmpcr_b200.h).

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only: every function here takes
CUDA tensors, hands their raw device pointers and the current CUDA stream to the C ABI, and returns fresh
tensors.  There is NO CPU fallback: if the shared library is missing, or the tensors are not on a CUDA (sm_100)
device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblmpcr_b200.so")

NN_EXACT_SIMT, NN_TENSOR = 0, 1
MUTUAL_INDEX, MUTUAL_GEOMETRIC = 0, 1
BN_EVAL, BN_BATCH = 0, 1
GUARD_BATCH, GUARD_PAIR = 0, 1
STATUS_ZERO_WEIGHT, STATUS_DEGENERATE = 1, 2

EXPORTS = [
    "lmpcr_abi_version", "lmpcr_launch_count", "lmpcr_launch_count_named", "lmpcr_last_error", "lmpcr_nn_tensor_debug", "lmpcr_debug_tc_profile", "lmpcr_debug_pcn_profile", "lmpcr_debug_oaf_profile", "lmpcr_debug_pool_profile", "lmpcr_debug_ktime_enable", "lmpcr_debug_ktime_read", "lmpcr_conv1x1", "lmpcr_conv1x1_workspace_bytes", "lmpcr_nn_soft", "lmpcr_nn_top2", "lmpcr_nn_top2_algo", "lmpcr_softmax_pool", "lmpcr_softmax_pool_workspace_bytes", "lmpcr_softmax_unpool", "lmpcr_softmax_unpool_workspace_bytes", "lmpcr_overlap_workspace_bytes", "lmpcr_overlap_count", "lmpcr_voxel_downsample", "lmpcr_device_info", "lmpcr_nn_workspace_bytes", "lmpcr_nn_argmin",
    "lmpcr_pairwise_distance", "lmpcr_gather_xyz", "lmpcr_mutual_xs", "lmpcr_knn3d_1", "lmpcr_kabsch", "lmpcr_residuals",
    "lmpcr_filter_num_params", "lmpcr_filter_workspace_bytes", "lmpcr_filter_forward", "lmpcr_pack_pose_records",
    "lmpcr_filter_pack_bytes", "lmpcr_filter_pack_weights", "lmpcr_filter_forward_packed",
    "lmpcr_pointcn_stack", "lmpcr_pointcn_stack_workspace_bytes",
    "lmpcr_sample_workspace_bytes", "lmpcr_sample_keypoints",
    "lmpcr_diff_pool_fused_workspace_bytes", "lmpcr_diff_pool_fused",
    "lmpcr_embed_fused_workspace_bytes", "lmpcr_embed_fused",
    "lmpcr_conv_wide_workspace_bytes", "lmpcr_conv_wide",
    "lmpcr_oafilter_stack_workspace_bytes", "lmpcr_oafilter_stack",
]


class LmpcrError(RuntimeError):
    pass


class FilterCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("net_channel", "clusters", "net_depth", "iter_num", "side_channel", "guard_mode", "gemm_algo", "bn_mode")]


_lib = None
_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


def load():
    """Loads the C-ABI library (once).  Raises LmpcrError when it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LMPCR_B200_LIB", LIB_PATH)      # override: an experimental build of the same ABI
    if not os.path.exists(path):
        raise LmpcrError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU / PyTorch fallback for this path)" % path)
    lib = ctypes.CDLL(path)
    lib.lmpcr_abi_version.restype = _i
    lib.lmpcr_last_error.restype = ctypes.c_char_p
    lib.lmpcr_device_info.argtypes = [ctypes.POINTER(_i)] * 4
    lib.lmpcr_nn_workspace_bytes.restype = _sz
    lib.lmpcr_nn_workspace_bytes.argtypes = [_i] * 7
    lib.lmpcr_nn_argmin.argtypes = [_vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp]
    lib.lmpcr_sample_workspace_bytes.restype = _sz
    lib.lmpcr_sample_workspace_bytes.argtypes = [_i, _i]
    lib.lmpcr_sample_keypoints.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_uint64, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_nn_top2.argtypes = [_vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_nn_top2_algo.argtypes = [_vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp]
    lib.lmpcr_overlap_workspace_bytes.restype = _sz
    lib.lmpcr_overlap_workspace_bytes.argtypes = [_i]
    lib.lmpcr_overlap_count.argtypes = [_vp, _i, _vp, _i, _vp, ctypes.c_double, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_voxel_downsample.argtypes = [_vp, _i, ctypes.c_double, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_softmax_pool_workspace_bytes.restype = _sz
    lib.lmpcr_softmax_pool_workspace_bytes.argtypes = [_i, _i, _i, _i]
    lib.lmpcr_softmax_pool.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]
    lib.lmpcr_softmax_unpool_workspace_bytes.restype = _sz
    lib.lmpcr_softmax_unpool_workspace_bytes.argtypes = [_i, _i, _i, _i]
    lib.lmpcr_softmax_unpool.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]
    lib.lmpcr_nn_soft.argtypes = [_vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _f, _vp, _vp, _sz, _vp]
    lib.lmpcr_launch_count.restype = ctypes.c_longlong
    lib.lmpcr_launch_count_named.restype = ctypes.c_longlong
    lib.lmpcr_launch_count_named.argtypes = [ctypes.c_char_p]
    lib.lmpcr_nn_tensor_debug.argtypes = [_vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_pairwise_distance.argtypes = [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]
    lib.lmpcr_gather_xyz.argtypes = [_vp, _i, _vp, _i, _vp, _i, _vp, _vp]
    lib.lmpcr_mutual_xs.argtypes = [_vp, _i, _vp, _i, _vp, _vp, _i, _f, _vp, _vp, _i, _vp]
    lib.lmpcr_knn3d_1.argtypes = [_vp, _i, _vp, _i, _i, _vp, _vp, _vp]
    lib.lmpcr_kabsch.argtypes = [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lmpcr_residuals.argtypes = [_vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp]
    lib.lmpcr_conv1x1_workspace_bytes.restype = _sz
    lib.lmpcr_conv1x1_workspace_bytes.argtypes = [_i, _i]
    lib.lmpcr_conv1x1.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _sz, _vp]
    lib.lmpcr_filter_num_params.argtypes = [ctypes.POINTER(FilterCfg)]
    lib.lmpcr_filter_workspace_bytes.restype = _sz
    lib.lmpcr_filter_workspace_bytes.argtypes = [ctypes.POINTER(FilterCfg), _i, _i]
    lib.lmpcr_filter_forward.argtypes = [_vp, _i, _i, ctypes.POINTER(_vp), _i, ctypes.POINTER(FilterCfg),
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_pack_pose_records.argtypes = [_vp, _vp, _vp, _vp, _i, _vp, _vp]
    lib.lmpcr_pointcn_stack_workspace_bytes.restype = _sz
    lib.lmpcr_pointcn_stack_workspace_bytes.argtypes = [_i, _i]
    lib.lmpcr_pointcn_stack.argtypes = [_vp, _i, _i, ctypes.POINTER(_vp), _i, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_oafilter_stack_workspace_bytes.restype = _sz
    lib.lmpcr_oafilter_stack_workspace_bytes.argtypes = [_i, _i, _i]
    lib.lmpcr_oafilter_stack.argtypes = [_vp, _i, _i, ctypes.POINTER(_vp), _i, _vp, _vp, _sz, _vp]
    lib.lmpcr_conv_wide_workspace_bytes.restype = _sz
    lib.lmpcr_conv_wide_workspace_bytes.argtypes = []
    lib.lmpcr_conv_wide.argtypes = [_vp, _i, _i] + [_vp] * 12 + [_vp, _sz, _vp]
    lib.lmpcr_embed_fused_workspace_bytes.restype = _sz
    lib.lmpcr_embed_fused_workspace_bytes.argtypes = [_i, _i, _i]
    lib.lmpcr_embed_fused.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]
    lib.lmpcr_diff_pool_fused_workspace_bytes.restype = _sz
    lib.lmpcr_diff_pool_fused_workspace_bytes.argtypes = [_i, _i]
    lib.lmpcr_diff_pool_fused.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]
    lib.lmpcr_filter_pack_bytes.restype = _sz
    lib.lmpcr_filter_pack_bytes.argtypes = [ctypes.POINTER(FilterCfg)]
    lib.lmpcr_filter_pack_weights.argtypes = [ctypes.POINTER(_vp), _i, ctypes.POINTER(FilterCfg), _vp, _sz, _vp]
    lib.lmpcr_filter_forward_packed.argtypes = [_vp, _i, _i, ctypes.POINTER(_vp), _i, ctypes.POINTER(FilterCfg), _vp, _sz,
                                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
    if lib.lmpcr_abi_version() != 1:
        raise LmpcrError("liblmpcr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise LmpcrError("lmpcr error %d: %s" % (rc, load().lmpcr_last_error().decode()))


def _dev(t, dtype=torch.float32, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise LmpcrError("%s must be a CUDA tensor (the B200 path has no CPU fallback)" % name)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def device_info():
    v = [ctypes.c_int(0) for _ in range(4)]
    _check(load().lmpcr_device_info(*[ctypes.byref(x) for x in v]))
    return {"sm_count": v[0].value, "l2_bytes": v[1].value, "cc": (v[2].value, v[3].value)}


# ---------------------------------------------------------------------------------------------- stage 1
def nn_argmin(q_feat, b_feat, jobs, algo=NN_EXACT_SIMT, return_dist=False):
    """q_feat [Sq,n,D], b_feat [Sb,m,D] fp32 CUDA; jobs [J,2] int32 (query set, target set) -> idx [J,n] int32."""
    lib = load()
    q = _dev(q_feat, name="q_feat")
    b = q if b_feat is q_feat else _dev(b_feat, name="b_feat")
    jobs = _dev(jobs, torch.int32, "jobs")
    assert q.dim() == 3 and b.dim() == 3 and q.shape[2] == b.shape[2] and jobs.dim() == 2 and jobs.shape[1] == 2
    J, n = jobs.shape[0], q.shape[1]
    with torch.cuda.device(q.device):
        idx = torch.empty((J, n), dtype=torch.int32, device=q.device)
        dist = torch.empty((J, n), dtype=torch.float32, device=q.device) if return_dist else None
        nbytes = lib.lmpcr_nn_workspace_bytes(q.shape[0], n, b.shape[0], b.shape[1], q.shape[2], J, algo)
        ws = _ws(nbytes, q.device)
        _check(lib.lmpcr_nn_argmin(_p(q), q.shape[0], n, _p(b), b.shape[0], b.shape[1], q.shape[2], _p(jobs), J, _p(idx), _p(dist),
                                   algo, _p(ws), ws.numel(), _stream(q)))
    return (idx, dist) if return_dist else idx


def nn_top2(q_feat, b_feat, jobs, algo=None):
    """Two nearest neighbours: (idx [J,n,2] int32, squared fp32 distances [J,n,2])  (scripts/extract_data.py:178-184).
    algo None: the tcgen05 path for 32-d features (bit-identical to the exact CUDA-core kernel), else NN_EXACT_SIMT."""
    lib = load()
    q = _dev(q_feat, name="q_feat")
    b = q if b_feat is q_feat else _dev(b_feat, name="b_feat")
    jobs = _dev(jobs, torch.int32, "jobs")
    J, n = jobs.shape[0], q.shape[1]
    if algo is None:
        algo = NN_TENSOR if (q.shape[2] == 32 and b.shape[1] >= 2) else NN_EXACT_SIMT
    with torch.cuda.device(q.device):
        idx = torch.empty((J, n, 2), dtype=torch.int32, device=q.device)
        dist = torch.empty((J, n, 2), dtype=torch.float32, device=q.device)
        ws = _ws(lib.lmpcr_nn_workspace_bytes(q.shape[0], n, b.shape[0], b.shape[1], q.shape[2], J, int(algo)), q.device)
        _check(lib.lmpcr_nn_top2_algo(_p(q), q.shape[0], n, _p(b), b.shape[0], b.shape[1], q.shape[2], _p(jobs), J, _p(idx), _p(dist), int(algo),
                                      _p(ws), ws.numel(), _stream(q)))
    return idx, dist


def overlap_count(query, target, T, radius):
    """Number of `query` points [n,3] (fp64, CUDA) with a point of `target` [m,3], moved by the 4x4 pose T (or None), closer than
    `radius` (lib/utils.py:743-751).  Returns a python int (one device->host read)."""
    lib = load()
    q = _dev(query, torch.float64, "query")
    b = _dev(target, torch.float64, "target")
    Tm = _dev(T, torch.float64, "T") if T is not None else None
    with torch.cuda.device(q.device):
        out = torch.zeros(2, dtype=torch.int32, device=q.device)
        ws = _ws(lib.lmpcr_overlap_workspace_bytes(b.shape[0]), q.device)
        _check(lib.lmpcr_overlap_count(_p(q), q.shape[0], _p(b), b.shape[0], _p(Tm), float(radius), _p(out), _p(out[1:]), _p(ws), ws.numel(),
                                       _stream(q)))
        cnt, flag = out.tolist()
    if flag:
        raise LmpcrError("lmpcr_overlap_count: coordinates outside the hash grid (|x| >= 2^20 * radius)")
    return cnt


def voxel_downsample(points, voxel_size):
    """Open3D-style voxel_down_sample of fp64 points [n,3] (CUDA) -> [n_voxels,3] (lib/utils.py:754-762)."""
    lib = load()
    x = _dev(points, torch.float64, "points")
    with torch.cuda.device(x.device):
        out = torch.empty((max(x.shape[0], 1), 3), dtype=torch.float64, device=x.device)
        cnt = torch.zeros(2, dtype=torch.int32, device=x.device)
        ws = _ws(lib.lmpcr_overlap_workspace_bytes(x.shape[0]), x.device)
        _check(lib.lmpcr_voxel_downsample(_p(x), x.shape[0], float(voxel_size), _p(out), _p(cnt), _p(cnt[1:]), _p(ws), ws.numel(), _stream(x)))
        n, flag = cnt.tolist()
    if flag:
        raise LmpcrError("lmpcr_voxel_downsample: extent exceeds 2^20 voxels per axis")
    return out[:n]


def sample_keypoints(coords, feats, pts_list, n_samples, replace, seed):
    """Device keypoint sampler (lib/layers.py:90-154, 'rand'): coords [total,3], feats [total,dim] fp32 CUDA = the concatenated
    clouds of a batch, pts_list = points per cloud.  Returns (idx [b,m] int32 global rows, coords [b,m,3], feats [b,m,dim])."""
    lib = load()
    c = _dev(coords, torch.float32, "coords")
    f = _dev(feats, torch.float32, "feats")
    pts = [int(v) for v in pts_list]
    b, total = len(pts), int(sum(pts))
    if c.shape[0] != total or f.shape[0] != total:
        raise LmpcrError("sample_keypoints: pts_list sums to %d but coords / feats have %d / %d rows" % (total, c.shape[0], f.shape[0]))
    off = [0]
    for v in pts:
        off.append(off[-1] + v)
    off_h = (ctypes.c_int32 * (b + 1))(*off)
    with torch.cuda.device(c.device):
        off_d = torch.tensor(off, dtype=torch.int32, device=c.device)
        idx = torch.empty((b, n_samples), dtype=torch.int32, device=c.device)
        co = torch.empty((b, n_samples, 3), dtype=torch.float32, device=c.device)
        fo = torch.empty((b, n_samples, f.shape[1]), dtype=torch.float32, device=c.device)
        ws = _ws(lib.lmpcr_sample_workspace_bytes(total, b), c.device)
        _check(lib.lmpcr_sample_keypoints(_p(c), _p(f), _p(off_d), ctypes.cast(off_h, _vp), b, f.shape[1], int(n_samples), 1 if replace else 0,
                                          ctypes.c_uint64(int(seed) & (2**64 - 1)), _p(idx), _p(co), _p(fo), _p(ws), ws.numel(), _stream(c)))
    return idx, co, fo


def nn_soft(q_feat, b_feat, b_xyz, jobs, temperature):
    """Soft correspondences: out[j,i,:] = softmax_k(-dist/T) @ b_xyz  (lib/layers.py:59-70,86, st=False)."""
    lib = load()
    q = _dev(q_feat, name="q_feat")
    b = q if b_feat is q_feat else _dev(b_feat, name="b_feat")
    x = _dev(b_xyz, name="b_xyz")
    jobs = _dev(jobs, torch.int32, "jobs")
    J, n = jobs.shape[0], q.shape[1]
    with torch.cuda.device(q.device):
        out = torch.empty((J, n, 3), dtype=torch.float32, device=q.device)
        ws = _ws(lib.lmpcr_nn_workspace_bytes(q.shape[0], n, b.shape[0], b.shape[1], q.shape[2], J, NN_EXACT_SIMT), q.device)
        _check(lib.lmpcr_nn_soft(_p(q), q.shape[0], n, _p(b), _p(x), b.shape[0], b.shape[1], q.shape[2], _p(jobs), J, float(temperature),
                                 _p(out), _p(ws), ws.numel(), _stream(q)))
    return out


def launch_count():
    return int(load().lmpcr_launch_count())


def ktime_enable(on=True):
    """Start (and reset) / stop the per-kernel device timers (lmpcr_debug_ktime_enable)."""
    load().lmpcr_debug_ktime_enable(1 if on else 0)


def ktime_read(kernel_name):
    """(launches, total milliseconds) of `kernel_name` since ktime_enable(True); synchronises on the recorded events."""
    n, ms = ctypes.c_int(0), ctypes.c_float(0.0)
    lib = load()
    lib.lmpcr_debug_ktime_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)]
    _check(lib.lmpcr_debug_ktime_read(kernel_name.encode(), ctypes.byref(n), ctypes.byref(ms)))
    return n.value, ms.value


def launch_count_named(kernel_name):
    return int(load().lmpcr_launch_count_named(kernel_name.encode()))


def nn_tensor_debug(q_feat, b_feat, jobs):
    """Tensor-path NN with the raw screening scores: returns (idx [J,n], dist [J,n], scores [J,n,m_pad], approx_min [J,n])."""
    lib = load()
    q = _dev(q_feat, name="q_feat")
    b = q if b_feat is q_feat else _dev(b_feat, name="b_feat")
    jobs = _dev(jobs, torch.int32, "jobs")
    J, n, m = jobs.shape[0], q.shape[1], b.shape[1]
    m_pad = (m + 255) // 256 * 256
    with torch.cuda.device(q.device):
        idx = torch.empty((J, n), dtype=torch.int32, device=q.device)
        dist = torch.empty((J, n), dtype=torch.float32, device=q.device)
        scores = torch.zeros((J, n, m_pad), dtype=torch.float32, device=q.device)
        amin = torch.empty((J, n), dtype=torch.float32, device=q.device)
        ws = _ws(lib.lmpcr_nn_workspace_bytes(q.shape[0], n, b.shape[0], m, q.shape[2], J, NN_TENSOR), q.device)
        _check(lib.lmpcr_nn_tensor_debug(_p(q), q.shape[0], n, _p(b), b.shape[0], m, q.shape[2], _p(jobs), J, _p(idx), _p(dist),
                                         _p(scores), _p(amin), _p(ws), ws.numel(), _stream(q)))
    return idx, dist, scores, amin


def pairwise_distance(src, dst):
    lib = load()
    s, d = _dev(src, name="src"), _dev(dst, name="dst")
    B, n, dim = s.shape
    m = d.shape[1]
    with torch.cuda.device(s.device):
        out = torch.empty((B, n, m), dtype=torch.float32, device=s.device)
        ws = _ws(4 * (B * n + B * m) + 1024, s.device)
        _check(lib.lmpcr_pairwise_distance(_p(s), n, _p(d), m, dim, B, _p(out), _p(ws), ws.numel(), _stream(s)))
    return out


def gather_xyz(b_xyz, jobs, idx):
    lib = load()
    x = _dev(b_xyz, name="b_xyz")
    jobs = _dev(jobs, torch.int32, "jobs")
    idx = _dev(idx, torch.int32, "idx")
    with torch.cuda.device(x.device):
        out = torch.empty((idx.shape[0], idx.shape[1], 3), dtype=torch.float32, device=x.device)
        _check(lib.lmpcr_gather_xyz(_p(x), x.shape[1], _p(jobs), jobs.shape[0], _p(idx), idx.shape[1], _p(out), _stream(x)))
    return out


def mutual_xs(xyz, pairs, idx_st, idx_ts, mutual_mode=MUTUAL_INDEX, mutual_thresh=0.05, xs_channels=6, want_mutual=True):
    """xyz [S,n,3]; pairs [P,2]; idx_st/idx_ts [P,n] -> (mutual uint8 [P,n] | None, xs [P,1,n,C])."""
    lib = load()
    x = _dev(xyz, name="xyz")
    pairs = _dev(pairs, torch.int32, "pairs")
    idx_st = _dev(idx_st, torch.int32, "idx_st")
    idx_ts = _dev(idx_ts, torch.int32, "idx_ts") if idx_ts is not None else None
    P, n = idx_st.shape
    with torch.cuda.device(x.device):
        mutual = torch.empty((P, n), dtype=torch.uint8, device=x.device) if want_mutual else None
        xs = torch.empty((P, 1, n, xs_channels), dtype=torch.float32, device=x.device)
        _check(lib.lmpcr_mutual_xs(_p(x), n, _p(pairs), P, _p(idx_st), _p(idx_ts), mutual_mode, float(mutual_thresh), _p(mutual), _p(xs),
                                   xs_channels, _stream(x)))
    return mutual, xs


def knn3d_1(pos1, pos2):
    lib = load()
    a, b = _dev(pos1, name="pos1"), _dev(pos2, name="pos2")
    B, n, _ = a.shape
    m = b.shape[1]
    with torch.cuda.device(a.device):
        idx = torch.empty((B, m), dtype=torch.int32, device=a.device)
        sq = torch.empty((B, m), dtype=torch.float32, device=a.device)
        _check(lib.lmpcr_knn3d_1(_p(a), n, _p(b), m, B, _p(idx), _p(sq), _stream(a)))
    return sq, idx


# ---------------------------------------------------------------------------------------------- stage 3
def kabsch_xs(xs, w, guard_mode=GUARD_PAIR, want_conf=False):
    """xs [P,1,N,C>=6] or [P,N,C]; w [P,N] -> R [P,3,3], t [P,3,1], res [P,N], status [P] (, conf [P,4])."""
    lib = load()
    x = _dev(xs, name="xs")
    w = _dev(w, name="weights")
    C = x.shape[-1]
    P, N = w.shape
    assert x.numel() == P * N * C and C >= 6
    with torch.cuda.device(x.device):
        R = torch.empty((P, 3, 3), dtype=torch.float32, device=x.device)
        t = torch.empty((P, 3, 1), dtype=torch.float32, device=x.device)
        res = torch.empty((P, N), dtype=torch.float32, device=x.device)
        conf = torch.empty((P, 4), dtype=torch.float32, device=x.device) if want_conf else None
        status = torch.zeros((P,), dtype=torch.int32, device=x.device)
        wout = torch.empty_like(w)
        wout.copy_(w)
        _check(lib.lmpcr_kabsch(_p(x), ctypes.c_void_p(x.data_ptr() + 12), C, _p(w), P, N, guard_mode, None, _p(wout), _p(R), _p(t),
                                _p(res), _p(conf), _p(status), _stream(x)))
    return (R, t, res, status, conf) if want_conf else (R, t, res, status)


def kabsch_points(x1, x2, w):
    """x1, x2 [P,N,3]; w [P,N].  No zero-weight guard (as lib/utils.py:164 itself has none)."""
    lib = load()
    a, b, w = _dev(x1, name="x1"), _dev(x2, name="x2"), _dev(w, name="weights")
    P, N = w.shape
    with torch.cuda.device(a.device):
        R = torch.empty((P, 3, 3), dtype=torch.float32, device=a.device)
        t = torch.empty((P, 3, 1), dtype=torch.float32, device=a.device)
        res = torch.empty((P, N), dtype=torch.float32, device=a.device)
        status = torch.zeros((P,), dtype=torch.int32, device=a.device)
        _check(lib.lmpcr_kabsch(_p(a), _p(b), 3, _p(w), P, N, GUARD_BATCH, None, None, _p(R), _p(t), _p(res), None, _p(status), _stream(a)))
    return R, t, res, status


def residuals(x1, x2, R, t):
    lib = load()
    a, b = _dev(x1, name="x1"), _dev(x2, name="x2")
    R, t = _dev(R, name="R"), _dev(t, name="t")
    P, N, _ = a.shape
    with torch.cuda.device(a.device):
        res = torch.empty((P, N), dtype=torch.float32, device=a.device)
        _check(lib.lmpcr_residuals(_p(a), _p(b), 3, _p(R), _p(t), P, N, _p(res), _stream(a)))
    return res


# ---------------------------------------------------------------------------------------------- stage 2
def _param_ptrs(params, cfg, strict=False):
    """Device-pointer table of the state_dict tensors.  strict (training-mode BatchNorm: the kernels update running_mean /
    running_var in place through these pointers): every tensor must already be contiguous fp32 CUDA -- a silent copy would
    lose the update."""
    lib = load()
    n_par = lib.lmpcr_filter_num_params(ctypes.byref(cfg))
    if len(params) != n_par:
        raise LmpcrError("expected %d parameter tensors, got %d" % (n_par, len(params)))
    if strict:
        for p in params:
            if not (isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise LmpcrError("training-mode BatchNorm updates the running statistics in place: parameters and buffers must be "
                                 "contiguous float32 CUDA tensors (got %s %s)" % (getattr(p, "dtype", type(p)), getattr(p, "device", "")))
    keep = [_dev(p, name="parameter") for p in params]
    return keep, (ctypes.c_void_p * n_par)(*[p.data_ptr() for p in keep]), n_par


def filter_pack_weights(params, cfg):
    """The load_state_dict-time step: all GEMM weights -> bf16 hi/lo operand tiles, once (lmpcr_filter_pack_weights).
    Returns an opaque uint8 CUDA tensor for filter_forward(packed=...)."""
    lib = load()
    keep, table, n_par = _param_ptrs(params, cfg)
    dev = keep[0].device
    with torch.cuda.device(dev):
        packed = _ws(lib.lmpcr_filter_pack_bytes(ctypes.byref(cfg)), dev)
        _check(lib.lmpcr_filter_pack_weights(table, n_par, ctypes.byref(cfg), _p(packed), packed.numel(), _stream(packed)))
    return packed


def filter_forward(xs, params, cfg, want_latent=True, want_conf=True, workspace=None, packed=None):
    """xs [P,1,N,6+side] CUDA fp32; params: list of CUDA fp32 tensors in state_dict order (no num_batches_tracked);
    cfg: FilterCfg; packed: filter_pack_weights(params, cfg) or None (weights are then split inside the call).
    Returns dict(logits [I,P,N], scores [I,P,N], R [I,P,3,3], t [I,P,3,1], residuals [P,N],
    latent [P,C,N] | None, conf [P,4] | None, status [P])."""
    lib = load()
    x = _dev(xs, name="xs")
    assert x.dim() == 4 and x.shape[1] == 1 and x.shape[3] == 6 + cfg.side_channel
    P, N = x.shape[0], x.shape[2]
    keep, table, n_par = _param_ptrs(params, cfg, strict=(cfg.bn_mode == BN_BATCH))
    I = cfg.iter_num + 1
    dev = x.device
    with torch.cuda.device(dev):
        out = {
            "logits": torch.empty((I, P, N), dtype=torch.float32, device=dev),
            "scores": torch.empty((I, P, N), dtype=torch.float32, device=dev),
            "R": torch.empty((I, P, 3, 3), dtype=torch.float32, device=dev),
            "t": torch.empty((I, P, 3, 1), dtype=torch.float32, device=dev),
            "residuals": torch.empty((P, N), dtype=torch.float32, device=dev),
            "latent": torch.empty((P, cfg.net_channel, N), dtype=torch.float32, device=dev) if want_latent else None,
            "conf": torch.empty((P, 4), dtype=torch.float32, device=dev) if want_conf else None,
            "status": torch.empty((P,), dtype=torch.int32, device=dev),
        }
        if P == 0:
            return out
        if workspace is None:
            workspace = _ws(lib.lmpcr_filter_workspace_bytes(ctypes.byref(cfg), P, N), dev)
        if packed is not None:
            _check(lib.lmpcr_filter_forward_packed(_p(x), P, N, table, n_par, ctypes.byref(cfg), _p(packed), packed.numel(),
                                                   _p(out["logits"]), _p(out["scores"]), _p(out["R"]), _p(out["t"]), _p(out["residuals"]),
                                                   _p(out["latent"]), _p(out["conf"]), _p(out["status"]), _p(workspace), workspace.numel(),
                                                   _stream(x)))
        else:
            _check(lib.lmpcr_filter_forward(_p(x), P, N, table, n_par, ctypes.byref(cfg), _p(out["logits"]), _p(out["scores"]), _p(out["R"]),
                                            _p(out["t"]), _p(out["residuals"]), _p(out["latent"]), _p(out["conf"]), _p(out["status"]),
                                            _p(workspace), workspace.numel(), _stream(x)))
    out["_workspace"] = workspace      # handed back so that callers can keep it for the next call
    return out


def conv1x1(x, weight, bias=None, scale=None, shift=None, residual=None, gemm_algo=1, out=None, workspace=None):
    """Fused (scale/shift + ReLU) -> 1x1 conv -> +bias (+residual).  x [P,cin,N], weight [cout,cin] -> out [P,cout,N]."""
    lib = load()
    x = _dev(x, name="x")
    w = _dev(weight, name="weight").reshape(weight.shape[0], -1)
    P, cin, N = x.shape
    cout = w.shape[0]
    opt = lambda t, n: _dev(t, name=n) if t is not None else None
    bias, scale, shift, residual = opt(bias, "bias"), opt(scale, "scale"), opt(shift, "shift"), opt(residual, "residual")
    with torch.cuda.device(x.device):
        if out is None:
            out = torch.empty((P, cout, N), dtype=torch.float32, device=x.device)
        if workspace is None:
            workspace = _ws(lib.lmpcr_conv1x1_workspace_bytes(cout, cin), x.device)
        _check(lib.lmpcr_conv1x1(_p(x), P, cin, N, _p(w), _p(bias), _p(scale), _p(shift), _p(residual), cout, _p(out), gemm_algo,
                                 _p(workspace), workspace.numel(), _stream(x)))
    return out


def pointcn_stack(x, layer_params, out=None, want_stats=False):
    """A stack of plain PointCN layers in one pair-resident launch (lmpcr_pointcn_stack).  x [P,128,N]; layer_params: list (one entry
    per layer) of 12 tensors in state_dict order: BN conv.1 (w, b, rm, rv), conv.3 (w, b), BN conv.5 (x4), conv.7 (w, b)."""
    lib = load()
    x = _dev(x, name="x")
    P, C, N = x.shape
    flat = [_dev(t, name="parameter").reshape(-1) if t.dim() != 4 else _dev(t, name="parameter") for lp in layer_params for t in lp]
    table = (ctypes.c_void_p * len(flat))(*[t.data_ptr() for t in flat])
    with torch.cuda.device(x.device):
        if out is None:
            out = torch.empty_like(x)
        stats = torch.empty((P, C, 2), dtype=torch.float32, device=x.device) if want_stats else None
        ws = _ws(lib.lmpcr_pointcn_stack_workspace_bytes(P, len(layer_params)), x.device)
        _check(lib.lmpcr_pointcn_stack(_p(x), P, N, table, len(layer_params), _p(out), _p(stats), _p(ws), ws.numel(), _stream(x)))
    return (out, stats) if want_stats else out


def oafilter_stack(x, layer_params):
    """The OAFilter stack of an OANBlock in one pair-resident launch (lmpcr_oafilter_stack).  x [P,128,K]; layer_params: list (one entry per
    layer) of 18 tensors in state_dict order: BN conv1.1 (w, b, rm, rv), conv1.3 (w, b), BN conv2.0 (x4, K channels), conv2.2 (w [K,K], b),
    BN conv3.2 (x4), conv3.4 (w, b).  Returns [P,128,K]."""
    lib = load()
    x = _dev(x, name="x")
    P, C, K = x.shape
    if C != 128:
        raise LmpcrError("oafilter_stack: x must be [P,128,K]")
    flat = [_dev(t, name="parameter") for lp in layer_params for t in lp]
    if len(flat) != 18 * len(layer_params):
        raise LmpcrError("oafilter_stack: 18 tensors per layer expected")
    table = (ctypes.c_void_p * len(flat))(*[t.data_ptr() for t in flat])
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        ws = _ws(lib.lmpcr_oafilter_stack_workspace_bytes(P, K, len(layer_params)), x.device)
        _check(lib.lmpcr_oafilter_stack(_p(x), P, K, table, len(layer_params), _p(out), _p(ws), ws.numel(), _stream(x)))
    return out


def diff_pool_fused(x, scale, shift, weight, mode=0):
    """diff_pool in one launch (lmpcr_diff_pool_fused): x [P,128,N], scale / shift [P,128], weight [K,128] -> [P,128,K].
    mode 0: single pass + fallback launch (production); 1: two passes for every item."""
    lib = load()
    x, sc, sh, w = _dev(x, name="x"), _dev(scale, name="scale"), _dev(shift, name="shift"), _dev(weight, name="weight")
    P, C, N = x.shape
    K = w.shape[0]
    with torch.cuda.device(x.device):
        out = torch.empty((P, C, K), dtype=torch.float32, device=x.device)
        ws = _ws(lib.lmpcr_diff_pool_fused_workspace_bytes(P, K), x.device)
        _check(lib.lmpcr_diff_pool_fused(_p(x), P, N, _p(sc), _p(sh), _p(w), K, int(mode), _p(out), _p(ws), ws.numel(), _stream(x)))
    return out


def conv_wide(x, convs, want_stats=False):
    """One or two 256 -> 128 convolutions over the same input in one launch (lmpcr_conv_wide).  x [P,256,N]; convs: list of 1-2 dicts with
    weight [128,256] and optional bias [128], scale / shift [P,256].  Returns a list of outputs [P,128,N] (and of [P,128,2] statistics)."""
    lib = load()
    x = _dev(x, name="x")
    P, C, N = x.shape
    opt = lambda t, n: _dev(t, name=n) if t is not None else None
    with torch.cuda.device(x.device):
        args, outs, stats, keep = [], [], [], []
        for i in range(2):
            if i < len(convs):
                c = convs[i]
                w, b, sc, sh = _dev(c["weight"], name="weight"), opt(c.get("bias"), "bias"), opt(c.get("scale"), "scale"), opt(c.get("shift"), "shift")
                o = torch.empty((P, 128, N), dtype=torch.float32, device=x.device)
                s = torch.empty((P, 128, 2), dtype=torch.float32, device=x.device) if want_stats else None
                keep += [w, b, sc, sh]
                outs.append(o); stats.append(s)
                args += [_p(w), _p(b), _p(sc), _p(sh), _p(o), _p(s)]
            else:
                args += [_p(None)] * 6
        ws = _ws(lib.lmpcr_conv_wide_workspace_bytes(), x.device)
        _check(lib.lmpcr_conv_wide(_p(x), P, N, *args, _p(ws), ws.numel(), _stream(x)))
    return (outs, stats) if want_stats else outs


def embed_fused(x, scale, shift, weight, bias=None, want_colmax=False):
    """The embedding conv of diff_unpool on the pair-resident kernel (lmpcr_embed_fused): x [P,128,N], scale / shift [P,128], weight [K,128],
    bias [K] -> embed [P,K,N] (and colmax [P,N] = log2(e) * max over the clusters)."""
    lib = load()
    x, sc, sh, w = _dev(x, name="x"), _dev(scale, name="scale"), _dev(shift, name="shift"), _dev(weight, name="weight")
    b = _dev(bias, name="bias") if bias is not None else None
    P, C, N = x.shape
    K = w.shape[0]
    with torch.cuda.device(x.device):
        E = torch.empty((P, K, N), dtype=torch.float32, device=x.device)
        cm = torch.empty((P, N), dtype=torch.float32, device=x.device) if want_colmax else None
        ws = _ws(lib.lmpcr_embed_fused_workspace_bytes(P, N, K), x.device)
        _check(lib.lmpcr_embed_fused(_p(x), P, N, _p(sc), _p(sh), _p(w), _p(b), K, _p(E), _p(cm), _p(ws), ws.numel(), _stream(x)))
    return (E, cm) if want_colmax else E


def softmax_pool(x, embed, mode=1):
    """diff_pool's weighted sum (oanet.py:107-109): x [P,C,N], embed [P,K,N] -> [P,C,K]; mode 0 separate statistics, 1 deferred."""
    lib = load()
    x, e = _dev(x, name="x"), _dev(embed, name="embed")
    P, C, N = x.shape
    K = e.shape[1]
    with torch.cuda.device(x.device):
        out = torch.empty((P, C, K), dtype=torch.float32, device=x.device)
        ws = _ws(lib.lmpcr_softmax_pool_workspace_bytes(P, C, K, N), x.device)
        _check(lib.lmpcr_softmax_pool(_p(x), _p(e), P, C, K, N, int(mode), _p(out), _p(ws), ws.numel(), _stream(x)))
    return out


def softmax_unpool(x_down, embed, mode=1):
    """diff_unpool's weighted sum (oanet.py:126-128): x_down [P,C,K], embed [P,K,N] -> [P,C,N]; softmax over the cluster axis."""
    lib = load()
    x, e = _dev(x_down, name="x_down"), _dev(embed, name="embed")
    P, C, K = x.shape
    N = e.shape[2]
    with torch.cuda.device(x.device):
        out = torch.empty((P, C, N), dtype=torch.float32, device=x.device)
        ws = _ws(lib.lmpcr_softmax_unpool_workspace_bytes(P, C, K, N), x.device)
        _check(lib.lmpcr_softmax_unpool(_p(x), _p(e), P, C, K, N, int(mode), _p(out), _p(ws), ws.numel(), _stream(x)))
    return out


def filter_workspace_bytes(cfg, P, N):
    return int(load().lmpcr_filter_workspace_bytes(ctypes.byref(cfg), P, N))


def reusable_workspace(cached, cfg, P, N, device):
    """`cached` if it is a scratch tensor on `device` large enough for (cfg, P, N), else None (filter_forward then allocates)."""
    if cached is None or cached.device != device or P <= 0:
        return None
    return cached if cached.numel() >= filter_workspace_bytes(cfg, P, N) else None


def pack_pose_records(R, t, conf, status):
    lib = load()
    R, t = _dev(R, name="R"), _dev(t, name="t")
    P = R.shape[0]
    with torch.cuda.device(R.device):
        rec = torch.empty((P, 16), dtype=torch.float32, device=R.device)
        _check(lib.lmpcr_pack_pose_records(_p(R), _p(t), _p(conf), _p(status), P, _p(rec), _stream(R)))
    return rec
