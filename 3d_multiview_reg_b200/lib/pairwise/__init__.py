"""Mirror of lib/pairwise/__init__.py:15-142 (`PairwiseReg`)."""
import torch
import torch.nn as nn

from ... import _cabi
from ..layers import Sampler, Soft_NN
from ..utils import construct_filtering_input_data, extract_mutuals, pair_indices
from . import config  # noqa: F401  (lib.pairwise.config.get_model, as in the reference)

__all__ = ["config", "PairwiseReg"]


class PairwiseReg(nn.Module):
    """Drop-in for lib/pairwise/__init__.py:15 with the same constructor signature, attributes, methods, dict keys
    and tensor layouts.

    descriptor_module: the reference wires FCGF (MinkowskiEngine) here; that network is outside this build.  Any
    callable mapping `input_dict` to per-point features [sum(pts_list), 32] may be supplied (objects with an `.F`
    attribute are unwrapped); alternatively pass the features directly as input_dict['features'].  With a falsy
    descriptor_module the model runs in precomputed mode exactly like the reference (:122-125).
    """

    def __init__(self, descriptor_module, filtering_module, device, samp_type="fps", corr_type="soft",
                 mutuals_flag=False, connectivity_info=None, tgt_num_points=2000,
                 straight_through_gradient=True, train_descriptor=False, sampler_rng="device"):
        super().__init__()
        self.device = device
        self.samp_type = samp_type
        self.corr_type = corr_type
        self.mutuals = mutuals_flag
        self.connectivity_info = connectivity_info
        self.train_descriptor = train_descriptor
        self.descriptor_module = descriptor_module
        if self.descriptor_module:
            # sampler_rng (not in the reference): 'device' = keypoints drawn on the GPU, 'numpy' = the reference's host stream
            self.sampler = Sampler(samp_type=self.samp_type, targeted_num_points=tgt_num_points, rng=sampler_rng)
            self.feature_matching = Soft_NN(corr_type=self.corr_type, st=straight_through_gradient, device=device)
            self.precomputed_desc = False
        else:
            self.precomputed_desc = True
        self.filtering_module = filtering_module

    def forward(self, data):
        filtering_input, f_0, f_1 = self.compute_descriptors(input_dict=data)
        registration_outputs = self.filter_correspondences(filtering_input)
        return filtering_input, f_0, f_1, registration_outputs

    def compute_descriptors(self, input_dict):
        """lib/pairwise/__init__.py:73-127.  Scan features/coords stay resident once ([S,n,.]); the NN kernel is
        given (source scan, target scan) index pairs, so no per-pair feature copies are made (cf. lib/utils.py:879)."""
        if self.precomputed_desc:
            return input_dict, None, None
        xyz_down = input_dict["pcd0"].to(self.device)
        if "features" in input_dict:
            F0 = input_dict["features"].to(self.device)
        else:
            F0 = self.descriptor_module(input_dict)
            F0 = getattr(F0, "F", F0).to(self.device)
        F1 = torch.empty(F0.shape[0], 0, device=self.device)
        xyz_batch, f_batch = self.sampler(xyz_down, F0, input_dict["pts_list"])
        pairs = self.connectivity_info if self.connectivity_info is not None else pair_indices(xyz_batch.shape[0])
        pairs = torch.as_tensor(pairs, dtype=torch.int32, device=xyz_batch.device).reshape(-1, 2)
        xyz_batch = xyz_batch.float().contiguous()
        f_batch = f_batch.float().contiguous()
        fm = self.feature_matching
        if fm.corr_type == "soft_gumbel":
            raise NotImplementedError("corr_type='soft_gumbel' (stochastic) is not built on the B200 path")
        pairs_r = pairs.flip(1).contiguous()
        if fm.corr_type == "soft" and not fm.st:
            # demo configuration: softmax-blended target coordinates (lib/layers.py:59-70,86)
            T = float(fm.get_temp().item())
            nn_C_s_t = _cabi.nn_soft(f_batch, f_batch, xyz_batch, pairs, T)                      # :110
            xyz_s = torch.index_select(xyz_batch, 0, pairs[:, 0].long())
            if self.mutuals:
                nn_C_t_s = _cabi.nn_soft(f_batch, f_batch, xyz_batch, pairs_r, T)                # :111
                xyz_t = torch.index_select(xyz_batch, 0, pairs[:, 1].long())
                self.last_mutuals = extract_mutuals(xyz_s, xyz_t, nn_C_s_t, nn_C_t_s)            # computed, not forwarded (Q2)
            xs = torch.cat((xyz_s, nn_C_s_t), dim=-1).unsqueeze(1)                               # lib/utils.py:915-926
            n_pairs, n = xs.shape[0], xs.shape[2]
        else:
            from ..layers import default_nn_algo
            algo = default_nn_algo(f_batch.shape[2])
            idx_st = _cabi.nn_argmin(f_batch, f_batch, pairs, algo=algo)                         # :110
            idx_ts = _cabi.nn_argmin(f_batch, f_batch, pairs_r, algo=algo)                       # :111
            if self.mutuals:
                # computed like the reference (:114-115) -- and, like there, not forwarded to the filter
                nn_C_s_t = _cabi.gather_xyz(xyz_batch, pairs, idx_st)
                nn_C_t_s = _cabi.gather_xyz(xyz_batch, pairs_r, idx_ts)
                xyz_s = torch.index_select(xyz_batch, 0, pairs[:, 0].long())
                xyz_t = torch.index_select(xyz_batch, 0, pairs[:, 1].long())
                self.last_mutuals = extract_mutuals(xyz_s, xyz_t, nn_C_s_t, nn_C_t_s)
            _, xs = _cabi.mutual_xs(xyz_batch, pairs, idx_st, idx_ts, want_mutual=False)
            n_pairs, n = idx_st.shape
        filtering_input = {"xs": xs, "ys": torch.zeros(n_pairs, n, 1), "ts": torch.zeros(n_pairs, 3, 1),
                           "Rs": torch.eye(3).unsqueeze(0).repeat(n_pairs, 1, 1)}               # lib/utils.py:911-913
        return filtering_input, F0, F1

    def filter_correspondences(self, input_dict):
        """lib/pairwise/__init__.py:131-142."""
        return self.filtering_module(input_dict)
