import importlib

import numpy as np
import torch

pkg = importlib.import_module("3d_multiview_reg_b200")
cabi = pkg._cabi


def cu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def load_oanet(sd, **cfg_over):
    """This package's OANet mirror on the GPU, loaded from a reference-named numpy state dict."""
    oanet = importlib.import_module("3d_multiview_reg_b200.lib.filtering.oanet")
    misc = dict(iter_num=1, net_depth=12, net_channel=128, clusters=500, normalize_weights=True, use_gpu=True)
    misc.update(cfg_over)
    net = oanet.OANet({"misc": misc, "data": {"use_mutuals": 0}}).eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return net.cuda()
