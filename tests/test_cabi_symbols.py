"""CPU: the C-ABI library loads, exports every symbol include/lmpcr_b200.h declares, and refuses to compute
without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from util import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "lmpcr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lmpcr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(cabi.EXPORTS) == syms
    assert lib.lmpcr_abi_version() == 1


def test_param_count_and_workspace_queries():
    lib = cabi.load()
    cfg = cabi.FilterCfg(128, 500, 12, 1, 0, 0, 0, 0)
    assert lib.lmpcr_filter_num_params(ctypes.byref(cfg)) == 334 - 46      # 46 num_batches_tracked buffers
    assert cabi.filter_workspace_bytes(cfg, 4, 5000) > 4 * 5000 * 4
    assert lib.lmpcr_nn_workspace_bytes(2, 5000, 2, 5000, 32, 2, 0) >= 2 * 5000 * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = cabi.load()
    buf = (ctypes.c_float * 64)()
    rc = lib.lmpcr_kabsch(buf, buf, 3, buf, 1, 4, 0, None, None, buf, buf, None, None, None, None)
    assert rc == -4 and b"no CPU fallback" in lib.lmpcr_last_error() or b"device" in lib.lmpcr_last_error().lower()
    with pytest.raises(cabi.LmpcrError):
        cabi.nn_argmin(torch.zeros(1, 8, 32), torch.zeros(1, 8, 32), torch.zeros(1, 2, dtype=torch.int32))
