"""ctypes access to oracle/_build/libbanded_ref.so (the C port of the reference's CPU path).
TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libbanded_ref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle`")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ref_chol_solve_batch.restype = ctypes.c_int64
        _lib.ref_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def chol_solve_batch(diag, sub, rhs=None, nthreads=0):
    """Reference path (block->band, banded Cholesky, band->block, re-band, banded solve) for a
    float64 batch ``diag [B,T,D,D]``, ``sub [B,T-1,D,D]`` (or None), ``rhs [B,T,D]`` (or None)."""
    diag = np.ascontiguousarray(diag, dtype=np.float64)
    b, t, d, _ = diag.shape
    sub = None if sub is None else np.ascontiguousarray(sub, dtype=np.float64)
    rhs = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
    ld = np.empty_like(diag)
    ls = None if sub is None else np.empty_like(sub)
    x = None if rhs is None else np.empty_like(rhs)
    info = np.zeros(b, dtype=np.int32)
    lib().ref_chol_solve_batch(
        _p(diag), _p(sub), _p(rhs), _p(ld), _p(ls), _p(x), _p(info),
        ctypes.c_int64(b), ctypes.c_int64(t), ctypes.c_int(d), ctypes.c_int(nthreads),
    )
    return ld, ls, x, info


def max_threads() -> int:
    return int(lib().ref_max_threads())
