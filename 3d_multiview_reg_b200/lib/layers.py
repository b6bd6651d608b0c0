"""Mirror of the reference's lib/layers.py for the hot path: Soft_NN (feature matching) and Sampler."""
import numpy as np
import torch

from .. import _cabi


class Soft_NN(torch.nn.Module):
    """lib/layers.py:10-88.  corr_type 'hard', and 'soft' with st=True (whose forward value is the hard match:
    y_hard - y_soft.detach() + y_soft, lib/layers.py:63-67), run on the sm_100a NN kernels; 'soft' with st=False (the demo
    configuration: softmax-blended target coordinates) runs the online-softmax kernel lmpcr_nn_soft.  The stochastic
    'soft_gumbel' mode is not built and raises.
    `_temperature` is kept so that checkpoints load (state_dict key feature_matching._temperature)."""

    def __init__(self, corr_type="soft", st=True, temp=0.3, min_temp=1e-4, device="cuda", algo=None):
        super().__init__()
        assert corr_type in ["soft", "hard", "soft_gumbel"], \
            "Wrong correspondence type selected. Must be one of [soft, soft_gumbel, hard]"
        self.device = device
        self.corr_type = corr_type
        self.st = st
        self.algo = algo
        self.register_buffer("min_temp", torch.tensor([min_temp]), persistent=False)
        self._temperature = torch.nn.Parameter(torch.tensor(temp, dtype=torch.float32))

    def get_temp(self):
        return torch.max(self._temperature ** 2, self.min_temp.to(self._temperature.device))

    def nn_indices(self, x_f, y_f):
        """argmin_j dist(x_f[b,i], y_f[b,j]) -> [b,n] int32 (bit-identical to lib/layers.py:81)."""
        b = x_f.shape[0]
        jobs = torch.arange(b, dtype=torch.int32, device=x_f.device).unsqueeze(1).repeat(1, 2)
        algo = self.algo if self.algo is not None else default_nn_algo(x_f.shape[2])
        return _cabi.nn_argmin(x_f, y_f, jobs, algo=algo), jobs

    def forward(self, x_f, y_f, y_c):
        if self.corr_type == "soft_gumbel":
            raise NotImplementedError("corr_type='soft_gumbel' (stochastic) is not built on the B200 path")
        if self.corr_type == "soft" and not self.st:
            b = x_f.shape[0]
            jobs = torch.arange(b, dtype=torch.int32, device=x_f.device).unsqueeze(1).repeat(1, 2)
            return _cabi.nn_soft(x_f, y_f, y_c, jobs, float(self.get_temp().item()))
        idx, jobs = self.nn_indices(x_f, y_f)
        return _cabi.gather_xyz(y_c, jobs, idx)


_NN_ALGO = {"value": _cabi.NN_TENSOR}     # tcgen05 screening + exact rescoring for 32-d features; exact SIMT otherwise


def default_nn_algo(dim):
    if _NN_ALGO["value"] is not None and dim == 32:
        return _NN_ALGO["value"]
    return _cabi.NN_EXACT_SIMT


def set_default_nn_algo(algo):
    _NN_ALGO["value"] = algo


class Sampler(torch.nn.Module):
    """lib/layers.py:90-154, samp_type 'rand'.  Same constructor, arguments and outputs ([b,m,3], [b,m,c]) as the reference.

    rng='device' (default): one `lmpcr_sample_keypoints` call for the whole batch -- Philox keys + a segmented sort on the GPU, no
    host round trip; the seed is drawn from torch's CPU generator, so `torch.manual_seed` makes the sample reproducible.  The draw
    is a uniformly random ordered subset like `np.random.choice(..., replace=False)`, but not numpy's number sequence.
    rng='numpy': the reference's own host stream (`np.random.choice` per cloud, lib/layers.py:141-145), for runs that must
    reproduce a numpy-seeded reference experiment index for index; the gather still runs on the device.
    'fps' depends on pointnet2_ops, whose import is commented out in the reference (lib/layers.py:7): it raises there, too."""

    def __init__(self, samp_type="fps", targeted_num_points=2000, rng="device"):
        super().__init__()
        assert samp_type in ["fps", "rand"], "Wrong sampling type selected. Must be one of [fps, rand]"
        assert rng in ["device", "numpy"]
        self.samp_type = samp_type
        self.targeted_num_points = targeted_num_points
        self.rng = rng

    def forward(self, input_C, input_F, pts_list):
        if self.samp_type != "rand":
            raise NotImplementedError("fps sampling needs pointnet2_ops (dead code in the reference)")
        pts = [int(v) for v in pts_list]
        # lib/layers.py:126,142-145: without replacement only if EVERY cloud of the batch has the targeted number of points
        replace = not (min(self.targeted_num_points, min(pts)) >= self.targeted_num_points)
        if self.rng == "device":
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            _, sampled_C, sampled_F = _cabi.sample_keypoints(input_C, input_F, pts, self.targeted_num_points, replace, seed)
            return sampled_C.to(input_C.dtype), sampled_F.to(input_F.dtype)
        sampled_C, sampled_F, start = [], [], 0
        for n in pts:
            rng = np.arange(start, start + n)
            idxs = np.random.choice(rng, self.targeted_num_points, replace=replace)
            idxs = torch.from_numpy(idxs).to(input_C.device).long()
            sampled_F.append(torch.index_select(input_F, 0, idxs))
            sampled_C.append(torch.index_select(input_C, 0, idxs))
            start += n
        return torch.stack(sampled_C, 0), torch.stack(sampled_F, 0)
