"""All-pairs registration of a multiview scene and its data-parallel sharding over GPUs.

The reference has no scene-level driver for this path: scripts/extract_data.py:146-147 loops `idx_1 < idx_2`
over all fragments on the CPU, and lib/pairwise/__init__.py:107 builds the same n-choose-2 list with
itertools.combinations.  Pairs are independent units, so the multi-GPU strategy is plain partitioning:

  * every rank holds all scans of the scene (features [S,n,32] + coordinates [S,n,3]; 42 MB for 60 scans);
  * the lexicographic pair list is cut into contiguous, equally sized (padded) ranges, one per rank, so the
    source scan stays fixed over long runs (L2 reuse) and the partition never changes per-pair arithmetic;
  * each rank runs stage 1 -> 2 -> 3 on its range and writes 16-float pose records
    (R 9, t 3, #inliers, sum w, rms residual, status);
  * ONE NCCL all_gather of the [shard,16] record blocks (159 KB per rank at 19,900 pairs) ends the step.

The zero-weight guard runs per pair here (LMPCR_GUARD_PAIR) so that results do not depend on which pairs share
a rank (the reference's guard is batch-coupled, oanet.py:177-178).
"""
import torch

from . import _cabi
from .lib.utils import pair_indices

RECORD_FLOATS = 16


def partition_pairs(n_pairs, world_size):
    """Contiguous ranges [start, stop) of ceil(P/W) pairs per rank (the last ranks may be short or empty)."""
    shard = (n_pairs + world_size - 1) // world_size if n_pairs > 0 else 0
    return shard, [(min(r * shard, n_pairs), min((r + 1) * shard, n_pairs)) for r in range(world_size)]


def unpack_records(rec):
    """[P,16] records -> dict(R [P,3,3], t [P,3,1], n_inliers, sum_w, rms_residual, status)."""
    return {"R": rec[:, 0:9].reshape(-1, 3, 3), "t": rec[:, 9:12].reshape(-1, 3, 1), "n_inliers": rec[:, 12],
            "sum_w": rec[:, 13], "rms_residual": rec[:, 14], "status": rec[:, 15].to(torch.int32)}


class SceneRegistrar:
    """Registers scan pairs of one scene with the sm_100a kernels.

    filtering_module: a `lib.filtering.oanet.OANet` (this package's mirror) already on the GPU.
    nn_algo: _cabi.NN_EXACT_SIMT | _cabi.NN_TENSOR.   nn_chunk: pairs per stage-1 call; pair_chunk: pairs per stage-2/3 call.
    """

    def __init__(self, filtering_module, nn_algo=_cabi.NN_EXACT_SIMT, pair_chunk=296, mutual_mode=_cabi.MUTUAL_INDEX,
                 mutual_thresh=0.05, nn_chunk=2048):
        self.net = filtering_module
        self.nn_algo = nn_algo
        self.pair_chunk = int(pair_chunk)
        self.nn_chunk = max(int(nn_chunk), self.pair_chunk)
        self.mutual_mode = mutual_mode
        self.mutual_thresh = mutual_thresh
        self._workspace = None      # filter scratch, kept across chunks and calls (sized by the largest chunk seen)
        self.timers = None          # optional dict of stage -> list[(start_event, stop_event)]

    def _param_table(self):
        # rebuilt per call (334 attribute look-ups): a cached table would go stale after net.to() / .half() / load_state_dict
        return self.net.param_table()

    def _tic(self, name):
        if self.timers is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return (name, ev)

    def _toc(self, tok):
        if tok is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.timers.setdefault(tok[0], []).append((tok[1], ev))

    def register_pairs(self, feats, xyz, pairs, keep_correspondences=False):
        """feats [S,n,D], xyz [S,n,3] (CUDA fp32), pairs [P,2] int32 (CUDA) -> records [P,16] (+ extras)."""
        dev = feats.device
        P = pairs.shape[0]
        rec = torch.empty((P, RECORD_FLOATS), dtype=torch.float32, device=dev)
        extras = {"idx_st": [], "idx_ts": [], "mutual": [], "scores": []} if keep_correspondences else None
        cfg = self.net.cabi_cfg()
        cfg.guard_mode = _cabi.GUARD_PAIR
        # eval-mode BatchNorm always: batch statistics would couple the pairs of a chunk (results would depend on pair_chunk, the
        # rank partition and the call order) and the kernels would overwrite running_mean / running_var -- a forgotten .eval()
        # must not do that to a scene
        cfg.bn_mode = _cabi.BN_EVAL
        params = self._param_table()
        packed = self.net.packed_weights(params)
        # stage 1 for large slabs of pairs, both directions in ONE call: the per-scan operand preparation of the tensor
        # path is then paid once per slab instead of twice per filter chunk
        for s0 in range(0, P, self.nn_chunk):
            ps = pairs[s0:s0 + self.nn_chunk].contiguous()
            ns = ps.shape[0]
            t = self._tic("nn")
            idx_both = _cabi.nn_argmin(feats, feats, torch.cat([ps, ps.flip(1)], 0).contiguous(), algo=self.nn_algo)
            self._toc(t)
            for q0 in range(0, ns, self.pair_chunk):
                pc = ps[q0:q0 + self.pair_chunk].contiguous()
                nc = pc.shape[0]
                idx_st, idx_ts = idx_both[q0:q0 + nc], idx_both[ns + q0:ns + q0 + nc]
                t = self._tic("mutual_xs")
                mutual, xs = _cabi.mutual_xs(xyz, pc, idx_st, idx_ts, self.mutual_mode, self.mutual_thresh,
                                             xs_channels=6 + cfg.side_channel, want_mutual=keep_correspondences or cfg.side_channel == 1)
                self._toc(t)
                t = self._tic("filter")
                ws = _cabi.reusable_workspace(self._workspace, cfg, nc, xs.shape[2], dev)
                out = _cabi.filter_forward(xs, params, cfg, want_latent=False, want_conf=True, workspace=ws, packed=packed)
                self._workspace = out.pop("_workspace", None)
                self._toc(t)
                t = self._tic("records")
                p0 = s0 + q0
                rec[p0:p0 + nc] = _cabi.pack_pose_records(out["R"][-1], out["t"][-1], out["conf"], out["status"])
                self._toc(t)
                if extras is not None:
                    extras["idx_st"].append(idx_st)
                    extras["idx_ts"].append(idx_ts)
                    extras["mutual"].append(mutual)
                    extras["scores"].append(out["scores"][-1])
        if extras is not None:
            extras = {k: torch.cat(v, 0) for k, v in extras.items()}
        return rec, extras

    def register_scene(self, feats, xyz, pairs=None, rank=0, world_size=1, group=None, gather=True):
        """All pairs of the scene (or the given list), sharded over `world_size` ranks.
        Returns records [P,16] for ALL pairs on every rank when gather=True (NCCL/gloo all_gather), else this
        rank's [shard_len,16] block and its (start, stop) range."""
        dev = feats.device
        if pairs is None:
            pairs = pair_indices(feats.shape[0], dev)
        pairs = pairs.to(device=dev, dtype=torch.int32)
        P = pairs.shape[0]
        shard, ranges = partition_pairs(P, world_size)
        start, stop = ranges[rank]
        mine, _ = self.register_pairs(feats, xyz, pairs[start:stop].contiguous())
        if world_size == 1 or not gather:
            return (mine if world_size == 1 else (mine, (start, stop)))
        return all_gather_records(mine, shard, P, world_size, group)


def all_gather_records(mine, shard, n_pairs, world_size, group=None):
    """Pads this rank's block to `shard` rows and all-gathers: the only collective of the path."""
    import torch.distributed as dist
    send = torch.zeros((shard, RECORD_FLOATS), dtype=torch.float32, device=mine.device)
    send[:mine.shape[0]] = mine
    recv = torch.empty((world_size * shard, RECORD_FLOATS), dtype=torch.float32, device=mine.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n_pairs]
