"""Host-buffer entry points: the operators for data that lives in HOST memory.

This is what a host-resident caller (the reference runs TensorFlow on CPU tensors) binds: thin
ctypes wrappers of the C ABI's ``mf_host_*`` functions (``include/markovflow_b200.h``,
``csrc/capi_host.cu``).  The chunking, the three-stream copy/compute pipeline and the device staging
slots live in the library; nothing here touches the data.  Inputs / outputs should be pinned
(``torch.empty(..., pin_memory=True)`` or ``mf_host_pin``) for the copies to be asynchronous.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, dtype_code, i64
from .block_tri_diag import _raise_if_failed


def _host(t: Optional[torch.Tensor], name: str, dtype=None) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if t.is_cuda:
        raise ValueError(f"{name}: the host entry points take CPU tensors")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must have dtype {dtype}")
    return t.contiguous()


def _hp(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _device_index(device) -> int:
    if device is None:
        return -1
    device = torch.device(device)
    return -1 if device.index is None else int(device.index)


def cholesky_solve_host(
    diag: torch.Tensor,
    sub: Optional[torch.Tensor],
    rhs: Optional[torch.Tensor] = None,
    out: Optional[Tuple[torch.Tensor, ...]] = None,
    chunk: int = 128,
    device: Optional[torch.device] = None,
):
    """``SymmetricBlockTriDiagonal(diag, sub).cholesky`` (+ ``.solve(rhs)``) on host tensors
    (``mf_host_btd_cholesky``).

    ``diag [B,T,D,D]``, ``sub [B,T-1,D,D]``, ``rhs [B,T,D]`` are CPU tensors; returns CPU tensors
    ``(Ld, Ls, x_or_None, info)`` (written into ``out`` when given).  Returns after all copies have
    landed.  ``cholesky_solve_host.last_bytes`` = ``(h2d_bytes, d2h_bytes)`` of the last call.
    """
    diag = _host(diag, "diag")
    dtype = diag.dtype
    sub, rhs = _host(sub, "sub", dtype), _host(rhs, "rhs", dtype)
    b, t, d, _ = diag.shape
    if out is None:
        pin = diag.is_pinned()
        ld_h = torch.empty_like(diag, pin_memory=pin)
        ls_h = torch.empty_like(sub, pin_memory=pin) if sub is not None else None
        x_h = torch.empty_like(rhs, pin_memory=pin) if rhs is not None else None
        info_h = torch.empty(b, dtype=torch.int32, pin_memory=pin)
    else:
        ld_h, ls_h, x_h, info_h = out
        for o in (ld_h, ls_h, x_h, info_h):
            if o is not None and (o.is_cuda or not o.is_contiguous()):
                raise ValueError("out tensors must be contiguous CPU tensors")
    moved = (ctypes.c_int64 * 2)()
    check(
        _lib.lib().mf_host_btd_cholesky(
            dtype_code(dtype), _hp(diag), _hp(sub), _hp(rhs), _hp(ld_h), _hp(ls_h), _hp(x_h), None,
            _hp(info_h), i64(b), i64(t), i64(d), i64(chunk), ctypes.c_int(_device_index(device)), moved),
        "mf_host_btd_cholesky",
    )
    cholesky_solve_host.last_bytes = (int(moved[0]), int(moved[1]))
    cholesky_solve_host.last_launches = (b + chunk - 1) // max(1, min(chunk, b))
    _raise_if_failed(info_h, "cholesky_solve_host")
    return ld_h, ls_h, x_h, info_h


def kalman_log_likelihood_host(
    initial_mean, chol_initial_covariance, state_transitions, state_offsets, chol_process_covariances,
    emission_matrix, observations, chol_obs_covariance, chunk_steps: int = 0,
    device: Optional[torch.device] = None, out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """``KalmanFilter.log_likelihood`` per series (``kalman_filter.py:184-255``) for state-space-model
    parameters held in HOST memory (``mf_host_kalman_log_likelihood``): ``state_transitions
    [B,T-1,D,D]`` ... ``observations [B,T,m]``, ``emission_matrix [T,m,D]`` or ``[B,T,m,D]``,
    ``chol_obs_covariance [m,m]`` or ``[T,m,m]``.  The series are streamed to the device in chunks of
    time and reduced there; only ``[B]`` values come back.  ``.last_bytes`` as above."""
    a = _host(state_transitions, "state_transitions")
    dtype = a.dtype
    mu0, l0 = _host(initial_mean, "initial_mean", dtype), _host(chol_initial_covariance, "chol_P0", dtype)
    bb, lq = _host(state_offsets, "state_offsets", dtype), _host(chol_process_covariances, "chol_Q", dtype)
    h, y = _host(emission_matrix, "emission_matrix", dtype), _host(observations, "observations", dtype)
    lr = _host(chol_obs_covariance, "chol_obs_covariance", dtype)
    bsz, n, d, _ = a.shape
    t = n + 1
    m = int(h.shape[-2])
    hb = 1 if h.dim() == 3 else bsz
    rs = 1 if lr.dim() == 2 else t
    if tuple(y.shape) != (bsz, t, m) or tuple(h.shape[-3:]) != (t, m, d):
        raise ValueError("observations / emission matrix do not match the state-space model")
    if out is None:
        out = torch.empty(bsz, dtype=dtype, pin_memory=a.is_pinned())
    moved = (ctypes.c_int64 * 2)()
    check(
        _lib.lib().mf_host_kalman_log_likelihood(
            dtype_code(dtype), _hp(mu0), _hp(l0), _hp(a), _hp(bb), _hp(lq), _hp(h), _hp(y), _hp(lr),
            _hp(out), i64(bsz), i64(t), i64(d), i64(m), i64(hb), i64(rs), i64(chunk_steps),
            ctypes.c_int(_device_index(device)), moved),
        "mf_host_kalman_log_likelihood",
    )
    kalman_log_likelihood_host.last_bytes = (int(moved[0]), int(moved[1]))
    return out
