"""Stationary Matern priors whose Kalman log-likelihood is evaluated WITHOUT materialising the SSM
(SURVEY.md §8f-2).

Reference path (``GaussianProcessRegression.log_likelihood``, ``models/gaussian_process_regression.py``):
``kernel.state_space_model(time_points)`` (``kernels/sde_kernel.py:153-171``) writes ``A_k``,
``chol Q_k``, ``b_k`` for every step, ``kernel.generate_emission_model`` (``:173-211``) writes ``H``,
and ``KalmanFilter.log_likelihood`` (``kalman_filter.py:184-255``) reads them back.  Here
``mf_kalman_matern_log_likelihood`` builds ``A_k, Q_k`` in registers from ``Δt_k`` (closed forms of
``kernels/matern.py:80-86,299-324,434-460`` and ``Q_k = P∞ − A_k P∞ A_kᵀ + jitter·I``,
``sde_kernel.py:421-446``): a step reads two values instead of ``2D²+2D+1``.

Only what that path needs is mirrored: ``Matern12/32/52(lengthscale, variance, jitter)`` with
``state_dim`` and ``kalman_log_likelihood``; the kernels' other methods stay with the caller.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr
from .interop import as_torch, boundary, require_cuda
from .kalman_filter import _workspace


@boundary
def matern_kalman_log_likelihood(state_dim: int, lengthscale, variance, observations,
                                 chol_obs_covariance, time_points=None, time_deltas=None,
                                 jitter: float = 0.0, first_is_initial: bool = True,
                                 return_element: bool = False):
    """Per-series marginal log-likelihood ``batch_shape`` of ``y_k = x_k[0] + N(0, R)`` under a
    Matern-(``state_dim`` − ½) prior.

    ``lengthscale``, ``variance``: scalars or ``batch_shape`` tensors; ``observations``:
    ``batch_shape + [T]`` (or ``[..., T, 1]``); ``chol_obs_covariance``: scalar / ``[1,1]``;
    exactly one of ``time_points`` ``batch_shape + [T]`` and ``time_deltas`` ``batch_shape + [T-1]``
    (``[T]`` with ``first_is_initial=False``: a later time segment of a longer series, in which case
    only the scan element is meaningful).  ``return_element=True`` also returns the
    ``[..., 3D²+2D+1]`` scan element of the whole segment (time-sharded evaluation).
    """
    if state_dim not in (1, 2, 3):
        raise ValueError("state_dim must be 1 (Matern12), 2 (Matern32) or 3 (Matern52)")
    y = as_torch(observations)
    require_cuda(y, "observations")
    if (time_points is None) == (time_deltas is None):
        raise ValueError("pass exactly one of time_points and time_deltas")
    times = as_torch(time_points if time_points is not None else time_deltas, y.device)
    if y.dim() == times.dim() + 1 and y.shape[-1] == 1:  # [..., T, 1] as the reference stores it
        y = y[..., 0]
    batch = tuple(y.shape[:-1])
    t = int(y.shape[-1])
    if t < 1:
        raise ValueError("need at least one observation")
    dtype, dev = y.dtype, y.device
    if time_points is not None:
        tp = as_torch(time_points, dev).to(dtype)
        if tuple(tp.shape) != batch + (t,):
            raise ValueError(f"time_points must be {batch + (t,)}, got {tuple(tp.shape)}")
        if not first_is_initial:
            raise ValueError("a later time segment needs time_deltas (the delta INTO its first step)")
        dts = tp[..., 1:] - tp[..., :-1]
    else:
        dts = as_torch(time_deltas, dev).to(dtype)
    nt = t - (1 if first_is_initial else 0)
    if tuple(dts.shape) != batch + (nt,):
        raise ValueError(f"time_deltas must be {batch + (nt,)}, got {tuple(dts.shape)}")
    bsz = 1
    for s in batch:
        bsz *= int(s)

    def per_series(v, name):
        v = torch.as_tensor(v, dtype=dtype, device=dev)
        if v.dim() == 0:
            v = v.expand(batch) if batch else v
        if tuple(v.shape) != batch:
            raise ValueError(f"{name} must be a scalar or {batch}, got {tuple(v.shape)}")
        return v.reshape(bsz).contiguous()

    ls = per_series(lengthscale, "lengthscale")
    var = per_series(variance, "variance")
    lr = torch.as_tensor(chol_obs_covariance, dtype=dtype, device=dev).reshape(-1)
    if lr.numel() != 1:
        raise ValueError("chol_obs_covariance must hold one value (output_dim is 1)")
    lr = lr.contiguous()
    y2 = y.reshape(bsz, t).contiguous()
    dt2 = dts.reshape(bsz, nt).contiguous()
    d = state_dim
    n = 3 * d * d + 2 * d + 1
    out = torch.empty(bsz, dtype=dtype, device=dev)
    elem = torch.empty(bsz, n, dtype=dtype, device=dev) if (return_element or not first_is_initial) else None
    lib = _lib.lib()
    lib.mf_kalman_matern_workspace_bytes.restype = _lib.ctypes.c_size_t
    nbytes = int(lib.mf_kalman_matern_workspace_bytes(dtype_code(dtype), i64(bsz), i64(t), i64(d)))
    ws = _workspace(nbytes, dev)
    check(
        lib.mf_kalman_matern_log_likelihood(
            dtype_code(dtype), ptr(ls), ptr(var), _lib.ctypes.c_double(float(jitter)), ptr(dt2),
            ptr(y2), ptr(lr), ptr(out), ptr(elem), i64(bsz), i64(t), i64(d),
            _lib.ctypes.c_int(1 if first_is_initial else 0), ptr(ws),
            _lib.ctypes.c_size_t(nbytes), current_stream()),
        "mf_kalman_matern_log_likelihood",
    )
    ll = out.reshape(batch)
    if return_element or not first_is_initial:
        return ll, elem.reshape(batch + (n,))
    return ll


class _Matern:
    """Common part of the mirrored Matern kernels (reference ``kernels/matern.py``)."""

    state_dim: int = 0

    def __init__(self, lengthscale, variance, output_dim: int = 1, jitter: float = 0.0) -> None:
        if output_dim != 1:
            raise ValueError("the fused path covers output_dim = 1")
        self.lengthscale = lengthscale
        self.variance = variance
        self.jitter = float(jitter)
        self.output_dim = output_dim

    @boundary
    def kalman_log_likelihood_per_chain(self, time_points, observations, chol_obs_covariance):
        return matern_kalman_log_likelihood(self.state_dim, self.lengthscale, self.variance,
                                            observations, chol_obs_covariance,
                                            time_points=time_points, jitter=self.jitter)

    @boundary
    def kalman_log_likelihood(self, time_points, observations, chol_obs_covariance) -> torch.Tensor:
        """``KalmanFilter(kernel.state_space_model(t), kernel.generate_emission_model(t), R, y)
        .log_likelihood()`` (reference ``kalman_filter.py:184-255``: summed over the batch)."""
        return torch.sum(self.kalman_log_likelihood_per_chain(time_points, observations,
                                                              chol_obs_covariance))


class Matern12(_Matern):
    """``kernels/matern.py:27-143``."""
    state_dim = 1


class Matern32(_Matern):
    """``kernels/matern.py:237-372``."""
    state_dim = 2


class Matern52(_Matern):
    """``kernels/matern.py:376-501``."""
    state_dim = 3
