"""Mirror of lib/filtering/oanet.py:188-265 (`OANet`).  The module holds parameters and buffers under exactly the
reference's state_dict names and runs the whole forward pass -- both OANBlocks and their
weighted-Kabsch calls -- through one C-ABI call (lmpcr_filter_forward).  Eval-mode numerics only."""
import logging
import math
import warnings

import torch
import torch.nn as nn

from ... import _cabi


def param_schema(net_channel, clusters, net_depth, iter_num, side_channel):
    """[(dotted name, shape, kind)] in the order of the reference's OANet.state_dict()."""
    C, K = net_channel, clusters
    half = (net_depth // (iter_num + 1)) // 2
    out = []

    def conv(p, co, ci):
        out.append((p + ".weight", (co, ci, 1, 1), "conv_w"))
        out.append((p + ".bias", (co,), "conv_b"))

    def bn(p, c):
        out.extend([(p + ".weight", (c,), "bn_w"), (p + ".bias", (c,), "bn_b"), (p + ".running_mean", (c,), "bn_rm"),
                    (p + ".running_var", (c,), "bn_rv"), (p + ".num_batches_tracked", (), "bn_nbt")])

    def pointcn(p, ci, co):
        if ci != co:
            conv(p + ".shot_cut", co, ci)
        bn(p + ".conv.1", ci), conv(p + ".conv.3", co, ci), bn(p + ".conv.5", co), conv(p + ".conv.7", co, co)

    for bname, cin in [("reg_init", 6 + side_channel)] + [("reg_iter.%d" % i, 8 + side_channel) for i in range(iter_num)]:
        conv(bname + ".conv1", C, cin)
        bn(bname + ".down1.conv.1", C), conv(bname + ".down1.conv.3", K, C)
        bn(bname + ".up1.conv.1", C), conv(bname + ".up1.conv.3", K, C)
        for i in range(half):
            pointcn("%s.l1_1.%d" % (bname, i), C, C)
        pointcn(bname + ".l1_2.0", 2 * C, C)
        for i in range(1, half):
            pointcn("%s.l1_2.%d" % (bname, i), C, C)
        for i in range(half):
            q = "%s.l2.%d" % (bname, i)
            bn(q + ".conv1.1", C), conv(q + ".conv1.3", C, C)
            bn(q + ".conv2.0", K), conv(q + ".conv2.2", K, K)
            bn(q + ".conv3.2", C), conv(q + ".conv3.4", C, C)
        conv(bname + ".output", 1, C)
    return out


class _Node(nn.Module):
    """Anonymous container; children / parameters are attached by dotted name."""


def _attach(root, dotted, tensor, is_param):
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    if is_param:
        node.register_parameter(parts[-1], nn.Parameter(tensor))
    else:
        node.register_buffer(parts[-1], tensor)


class OANet(nn.Module):
    """Drop-in for lib/filtering/oanet.py:188 `OANet(cfg)`.

    forward(data): data['xs'] [B,1,N,6(+1)] float32 on any device (moved with .to(self.device) exactly like
    oanet.py:234,240) -> dict with keys 'logits', 'scores', 'rot_est', 'trans_est' (python lists of length
    iter_num+1), 'latent features' [B,C,N,1], 'gradient_flag' (oanet.py:257-263).  Extra keys (not in the
    reference): 'residuals' [B,N], 'confidence' [B,4], 'status' [B]."""

    def __init__(self, cfg):
        super().__init__()
        m = cfg["misc"]
        self.iter_num = m["iter_num"]
        self.net_depth = m["net_depth"]
        self.net_channel = m["net_channel"]
        self.clusters = m["clusters"]
        self.side_channel = (cfg["data"]["use_mutuals"] == 2)
        self.guard_mode = _cabi.GUARD_BATCH        # reference semantics (oanet.py:177-178); scene.py uses GUARD_PAIR
        # 1 = tcgen05 split-bf16 GEMMs (the measured path; the reference's YAMLs carry no such key, so they get it); 0 = fp32 CUDA cores
        self.gemm_algo = int(m.get("gemm_algo", 1))
        self._packed = None           # (key, packed weight blobs): rebuilt when a conv weight changes (load_state_dict, .to(), optimiser step)
        self._workspace = None        # scratch of the last forward, reused while it is large enough
        self.device = torch.device("cuda" if (torch.cuda.is_available() and m["use_gpu"]) else "cpu")
        self._schema = param_schema(self.net_channel, self.clusters, self.net_depth, self.iter_num, int(self.side_channel))
        for name, shape, kind in self._schema:
            if kind == "conv_w":
                w = torch.empty(shape)
                nn.init.kaiming_uniform_(w, a=math.sqrt(5))
                _attach(self, name, w, True)
            elif kind == "conv_b":
                bound = 1.0 / math.sqrt(self._fan_in(name))
                _attach(self, name, torch.empty(shape).uniform_(-bound, bound), True)
            elif kind == "bn_w":
                _attach(self, name, torch.ones(shape), True)
            elif kind == "bn_b":
                _attach(self, name, torch.zeros(shape), True)
            elif kind == "bn_rm":
                _attach(self, name, torch.zeros(shape), False)
            elif kind == "bn_rv":
                _attach(self, name, torch.ones(shape), False)
            else:
                _attach(self, name, torch.tensor(0, dtype=torch.long), False)
        self._warned_train = False
        logging.info("OANET(B200): channels:%d, clusters:%d, blocks:%d", self.net_channel, self.clusters, self.iter_num + 1)

    def _fan_in(self, bias_name):
        wname = bias_name[:-len("bias")] + "weight"
        for n, s, _ in self._schema:
            if n == wname:
                return s[1]
        return 1

    def cabi_cfg(self):
        return _cabi.FilterCfg(self.net_channel, self.clusters, self.net_depth, self.iter_num, int(self.side_channel),
                               self.guard_mode, self.gemm_algo, _cabi.BN_BATCH if self.training else _cabi.BN_EVAL)

    def param_table(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return [sd[n].detach() for n, _, kind in self._schema if kind != "bn_nbt"]

    def packed_weights(self, params=None):
        """bf16 hi/lo operand tiles of every GEMM weight (SURVEY.md 8b "packed once at load_state_dict"): cached per
        (data_ptr, _version) of the conv weights, so a load_state_dict / .to() / in-place update repacks and nothing else does."""
        if self.gemm_algo != 1:
            return None
        params = self.param_table() if params is None else params
        key = tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p, (_, _, kind) in
                    zip(params, [e for e in self._schema if e[2] != "bn_nbt"]) if kind == "conv_w")
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, _cabi.filter_pack_weights(params, self.cabi_cfg()))
        return self._packed[1]

    def forward(self, data):
        assert data["xs"].dim() == 4 and data["xs"].shape[1] == 1
        # self.training (scripts/benchmark_pairwise_registration.py never calls .eval()): BatchNorm uses the statistics of this
        # batch and the kernels update running_mean / running_var in place, like nn.BatchNorm2d; forward only -- no autograd graph
        if self.device.type != "cuda":
            raise _cabi.LmpcrError("OANet(B200) needs a CUDA device (cfg['misc']['use_gpu'] and an sm_100 GPU); no CPU fallback")
        xs = data["xs"].to(self.device, dtype=torch.float32)
        with torch.no_grad():
            params = self.param_table()
            cfg = self.cabi_cfg()
            ws = _cabi.reusable_workspace(self._workspace, cfg, xs.shape[0], xs.shape[2], xs.device)
            out = _cabi.filter_forward(xs, params, cfg, workspace=ws, packed=self.packed_weights(params))
            self._workspace = out.pop("_workspace", None)
            if self.training:
                for name, buf in self.named_buffers():
                    if name.endswith("num_batches_tracked"):
                        buf += 1
        n_it = self.iter_num + 1
        flag = bool((out["status"] & _cabi.STATUS_DEGENERATE).any().item())
        return {
            "logits": [out["logits"][i] for i in range(n_it)],
            "scores": [out["scores"][i] for i in range(n_it)],
            "rot_est": [out["R"][i] for i in range(n_it)],
            "trans_est": [out["t"][i] for i in range(n_it)],
            "latent features": out["latent"].unsqueeze(3),
            "gradient_flag": flag,
            "residuals": out["residuals"],
            "confidence": out["conf"],
            "status": out["status"],
        }
