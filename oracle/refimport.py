"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (zgojcic/3D_multiview_reg) from
/root/reference on CPU so that (a) the numpy/C restatement in this directory can be validated against
it and (b) golden vectors can be generated (tests/golden/make_golden.py).

/root/reference does not exist on the GPU box: nothing that runs there may import this module.
The reference imports open3d / nibabel / MinkowskiEngine / coloredlogs at module top
(lib/utils.py:8-9, lib/pairwise/__init__.py:3, lib/logger.py) -- none is installed here and none is
needed by the hot path, so empty stub modules are registered before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LMPCR_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference's top-level `lib` package (lib.utils, lib.layers, lib.filtering.oanet,
    lib.pairwise, lib.config are imported as a side effect)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch.nn as nn

    _stub("open3d")
    nib = _stub("nibabel")
    nq = _stub("nibabel.quaternions")
    nib.quaternions = nq
    me = _stub("MinkowskiEngine", MinkowskiNetwork=nn.Module)
    mf = _stub("MinkowskiEngine.MinkowskiFunctional")
    me.MinkowskiFunctional = mf
    _stub("coloredlogs", install=lambda *a, **k: None)
    _stub("tensorboardX", SummaryWriter=object)
    _stub("easydict", EasyDict=dict)
    # `lib` must resolve to the reference, not to the drop-in mirror of this repo.
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    lib = importlib.import_module("lib")
    for sub in ("lib.utils", "lib.layers", "lib.filtering", "lib.filtering.oanet", "lib.pairwise", "lib.config"):
        importlib.import_module(sub)
    return lib


def release_reference():
    """Remove the reference's `lib` from sys.modules/sys.path again."""
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
