"""3d_multiview_reg_b200 -- B200-native (sm_100a) pairwise-registration hot path of LMPCR
(zgojcic/3D_multiview_reg): feature-space mutual-NN -> correspondence-weighting network -> weighted Kabsch.

Layout
  csrc/            hand-written CUDA kernels + the extern "C" surface (include/lmpcr_b200.h) -> liblmpcr_b200.so
  _cabi.py         ctypes binding (raw device pointers + CUDA stream; torch only owns the memory)
  lib/             host-side mirror of the reference's module surface for this path:
                   lib.pairwise.PairwiseReg, lib.filtering.filtering_dict['oanet'], lib.layers.Soft_NN,
                   lib.utils.{kabsch_transformation_estimation, transformation_residuals, ...}, lib.config.get_model
  scene.py         all-pairs scene registration, pair partitioning over GPUs, NCCL pose all-gather

The directory name is not a Python identifier; import it with
    importlib.import_module("3d_multiview_reg_b200")
or call `install_as_lib()` to alias the mirror as the top-level package `lib`, which is what the reference's
scripts import (scripts/pairwise_demo.py:20-26, scripts/benchmark_pairwise_registration.py:27-36).
"""
import importlib
import sys

from . import _cabi  # noqa: F401
from ._cabi import LmpcrError  # noqa: F401

__all__ = ["install_as_lib", "build", "LmpcrError"]


def build(force=False):
    from . import _build as _b
    return _b.build(force=force)


def install_as_lib():
    """Register this package's `lib` mirror as the top-level module `lib` (and its sub-modules), so that
    `from lib.pairwise import ...` / `from lib import config` in the reference's scripts resolve here."""
    pkg = __name__
    names = ["lib", "lib.utils", "lib.layers", "lib.config", "lib.filtering", "lib.filtering.oanet", "lib.pairwise",
             "lib.pairwise.config"]
    for n in names:
        sys.modules[n] = importlib.import_module(pkg + "." + n)
    return sys.modules["lib"]
