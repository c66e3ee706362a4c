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
        for name in ("ref_kalman_loglik_batch", "ref_nat_to_ssm_batch", "ref_ssm_to_expectations_batch"):
            getattr(_lib, name).restype = ctypes.c_int64
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


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def kalman_loglik_batch(mu0, chol_p0, a_s, b_s, chol_q, h, y, chol_r, nthreads=0):
    """``KalmanFilter.log_likelihood`` per chain through the reference's SpInGP route
    (``kalman_filter.py:184-255``).  ``a_s [B,T-1,D,D]``, ``h [T,m,D]`` or ``[B,T,m,D]``, ``y [B,T,m]``,
    ``chol_r [m,m]``.  Returns ``[B]``."""
    mu0, chol_p0, a_s, b_s, chol_q, h, y, chol_r = map(_c, (mu0, chol_p0, a_s, b_s, chol_q, h, y, chol_r))
    b, n, d, _ = a_s.shape
    m = h.shape[-2]
    hb = 1 if h.ndim == 3 else b
    out = np.empty(b)
    nfail = lib().ref_kalman_loglik_batch(
        _p(mu0), _p(chol_p0), _p(a_s), _p(b_s), _p(chol_q), _p(h), _p(y), _p(chol_r), _p(out),
        ctypes.c_int64(b), ctypes.c_int64(n + 1), ctypes.c_int(d), ctypes.c_int(m), ctypes.c_int64(hb),
        ctypes.c_int(nthreads))
    if nfail:
        raise ArithmeticError(f"Banded Cholesky decomposition failure in {nfail} chains")
    return out


def nat_to_ssm_batch(theta_lin, theta_diag, theta_sub, nthreads=0):
    """``naturals_to_ssm_params`` (``ssm_gaussian_transformations.py:332-511``) for ``[B,T,...]`` inputs.
    Returns ``(As, offsets, chol_P0, chol_Qs, mu0)``."""
    theta_lin, theta_diag, theta_sub = map(_c, (theta_lin, theta_diag, theta_sub))
    b, t, d = theta_lin.shape
    a_s, offs = np.empty((b, t - 1, d, d)), np.empty((b, t - 1, d))
    l0, lq, mu0 = np.empty((b, d, d)), np.empty((b, t - 1, d, d)), np.empty((b, d))
    nfail = lib().ref_nat_to_ssm_batch(
        _p(theta_lin), _p(theta_diag), _p(theta_sub), _p(a_s), _p(offs), _p(l0), _p(lq), _p(mu0),
        ctypes.c_int64(b), ctypes.c_int64(t), ctypes.c_int(d), ctypes.c_int(nthreads))
    if nfail:
        raise ArithmeticError(f"Cholesky failure in {nfail} chains")
    return a_s, offs, l0, lq, mu0


def ssm_to_expectations_batch(mu0, chol_p0, a_s, b_s, chol_q, nthreads=0):
    """``ssm_to_expectations`` (``ssm_gaussian_transformations.py:31-89``) for ``[B,...]`` parameters."""
    mu0, chol_p0, a_s, b_s, chol_q = map(_c, (mu0, chol_p0, a_s, b_s, chol_q))
    b, n, d, _ = a_s.shape
    el, ed, es = np.empty((b, n + 1, d)), np.empty((b, n + 1, d, d)), np.empty((b, n, d, d))
    nfail = lib().ref_ssm_to_expectations_batch(
        _p(mu0), _p(chol_p0), _p(a_s), _p(b_s), _p(chol_q), _p(el), _p(ed), _p(es),
        ctypes.c_int64(b), ctypes.c_int64(n + 1), ctypes.c_int(d), ctypes.c_int(nthreads))
    if nfail:
        raise ArithmeticError(f"Cholesky failure in {nfail} chains")
    return el, ed, es
