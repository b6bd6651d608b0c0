"""Mirror of lib/config.py:10-26."""
import torch

from . import pairwise

method_dict = {"pairwise": pairwise}


def get_model(cfg):
    method = cfg["method"]["task"]
    device = torch.device("cuda" if (torch.cuda.is_available() and cfg["misc"]["use_gpu"]) else "cpu")
    return method_dict[method].config.get_model(cfg, device=device)
