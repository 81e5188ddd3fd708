"""ctypes loader for oracle/c/liboracle_c.so (plain-C DTW / median).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_DIR, "liboracle_c.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_DIR, "dtw_median.c")):
        subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle_c.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_dtw.restype = ctypes.c_int
        _LIB.oracle_dtw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]
        _LIB.oracle_median_filter.restype = None
        _LIB.oracle_median_filter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
                                              ctypes.c_int]
    return _LIB


def dtw(x: np.ndarray):
    """x float32 [N, M] -> (path int32 [2, n], jumps int32 [N])."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, M = x.shape
    ti = np.empty(N + M, np.int32)
    tj = np.empty(N + M, np.int32)
    jumps = np.zeros(N, np.int32)
    n = lib().oracle_dtw(x.ctypes.data, N, M, ti.ctypes.data, tj.ctypes.data, jumps.ctypes.data)
    assert n > 0
    return np.stack([ti[:n], tj[:n]]), jumps


def median_filter(x: np.ndarray, w: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().oracle_median_filter(x.ctypes.data, out.ctypes.data, int(np.prod(x.shape[:-1])), x.shape[-1], w)
    return out
