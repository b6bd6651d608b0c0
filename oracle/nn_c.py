"""ctypes loader for oracle/_build/libnn_oracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnn_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "nn_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "_build/libnn_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.lmpcr_oracle_threads.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nn_argmin(src, dst):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.ascontiguousarray(dst, np.float32)
    n, d = src.shape
    m = dst.shape[0]
    idx = np.empty(n, np.int32)
    best = np.empty(n, np.float32)
    lib().lmpcr_oracle_nn_argmin(_fp(src), ctypes.c_int(n), _fp(dst), ctypes.c_int(m), ctypes.c_int(d), _fp(idx), _fp(best))
    return idx, best


def pairwise_distance(src, dst):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.ascontiguousarray(dst, np.float32)
    out = np.empty((src.shape[0], dst.shape[0]), np.float32)
    lib().lmpcr_oracle_pairwise_distance(_fp(src), ctypes.c_int(src.shape[0]), _fp(dst), ctypes.c_int(dst.shape[0]),
                                         ctypes.c_int(src.shape[1]), _fp(out))
    return out


def sqnorm(f):
    f = np.ascontiguousarray(f, np.float32)
    out = np.empty(f.shape[0], np.float32)
    lib().lmpcr_oracle_sqnorm(_fp(f), ctypes.c_int(f.shape[0]), ctypes.c_int(f.shape[1]), _fp(out))
    return out


def threads():
    return int(lib().lmpcr_oracle_threads())
