"""Mirror of lib/pairwise/config.py:7-59 (model factory).  get_trainer is a training-time feature (out of scope)."""
from ..filtering import filtering_dict


def get_model(cfg, device):
    from . import PairwiseReg
    filtering_module = get_filter(cfg, device)
    descriptor_module = get_descriptor(cfg, device)
    return PairwiseReg(descriptor_module=descriptor_module, filtering_module=filtering_module, device=device,
                       samp_type=cfg["train"]["samp_type"], corr_type=cfg["train"]["corr_type"], connectivity_info=None,
                       tgt_num_points=cfg["data"]["max_num_points"], straight_through_gradient=cfg["train"]["st_grad_flag"])


def get_descriptor(cfg, device):
    name = cfg["method"]["descriptor_module"]
    if not name:
        return None
    raise NotImplementedError("descriptor_module=%r: the FCGF network (MinkowskiEngine) is outside this build; construct "
                              "PairwiseReg with your own descriptor callable or pass input_dict['features']" % name)


def get_filter(cfg, device):
    name = cfg["method"]["filter_module"]
    if not name:
        return None
    return filtering_dict[name](cfg).to(device)


def get_trainer(cfg, model, optimizer, logger, device):
    raise NotImplementedError("training is outside the scope of the B200 inference path")
