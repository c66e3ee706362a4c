"""Natural / expectation parameterisations of a :class:`StateSpaceModel`, with the function names,
argument order and return order of ``markovflow/ssm_gaussian_transformations.py:31-593``.

Each transform is one CUDA kernel (``mf_nat_to_ssm``, ``mf_ssm_to_naturals``,
``mf_ssm_to_expectations``, ``mf_expectations_to_ssm``).  ``naturals_to_ssm_params`` -- six
sweeps in the reference (banded Cholesky, sparse inverse subset, general solve, banded triangular
solve, two batched Choleskys) -- is a single backward ``U D Uᵀ`` sweep here.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr
from .block_tri_diag import _prod, _raise_if_failed
from .autograd import needs_grad
from .interop import as_torch, boundary, require_cuda
from .state_space_model import StateSpaceModel

Tensor = torch.Tensor


def _flat_params(lin, diag, sub):
    lin = as_torch(lin)
    diag = as_torch(diag, lin.device)
    sub = as_torch(sub, lin.device)
    require_cuda(lin, "parameters")
    if lin.dim() < 2 or diag.dim() != lin.dim() + 1 or sub.dim() != lin.dim() + 1:
        raise ValueError("expected [...,T,D], [...,T,D,D], [...,T-1,D,D]")
    batch = tuple(lin.shape[:-2])
    t, d = int(lin.shape[-2]), int(lin.shape[-1])
    if tuple(diag.shape) != batch + (t, d, d) or tuple(sub.shape) != batch + (t - 1, d, d):
        raise ValueError(
            f"inconsistent parameter shapes {tuple(lin.shape)}, {tuple(diag.shape)}, {tuple(sub.shape)}")
    if not (lin.dtype == diag.dtype == sub.dtype):
        raise ValueError("parameters must share a dtype")
    bsz = _prod(batch)
    return (lin.reshape(bsz, t, d).contiguous(), diag.reshape(bsz, t, d, d).contiguous(),
            sub.reshape(bsz, t - 1, d, d).contiguous(), batch, bsz, t, d)


def _ssm_outputs(entry: str, lin, diag, sub, *extra) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    lin, diag, sub, batch, bsz, t, d = _flat_params(lin, diag, sub)
    if needs_grad(lin, diag, sub):
        return _ssm_outputs_diff(entry, lin, diag, sub, batch, t, d, *extra)
    a = torch.empty_like(sub)
    off = torch.empty_like(lin)
    chol = torch.empty_like(diag)
    info = torch.empty(bsz, dtype=torch.int32, device=lin.device)
    fn = getattr(_lib.lib(), entry)
    check(
        fn(dtype_code(lin.dtype), ptr(lin), ptr(diag), ptr(sub), ptr(a), ptr(off), ptr(chol),
           ptr(info), i64(bsz), i64(t), i64(d), *extra, current_stream()),
        entry,
    )
    _raise_if_failed(info, entry)
    a = a.reshape(batch + (t - 1, d, d))
    off = off.reshape(batch + (t, d))
    chol = chol.reshape(batch + (t, d, d))
    # As, offsets, chol_initial_covariance, chol_process_covariances, initial_mean
    return a, off[..., 1:, :], chol[..., 0, :, :], chol[..., 1:, :, :], off[..., 0, :]


@boundary
def _ssm_outputs_diff(entry: str, lin, diag, sub, batch, t, d, *extra):
    """The same outputs with reverse mode (autograd.py).  ``expectations_to_ssm_params`` (a per-step map): the CUDA
    kernel runs forward, the torch restatement of the map is differentiated backward.  ``naturals_to_ssm_params``
    (a backward recursion): the block Cholesky + forward substitution of the time-reversed precision on the CUDA
    sweeps and their adjoint sweeps (``autograd.naturals_to_ssm_diff``)."""
    from .autograd import _RecomputeFn, expectations_to_ssm_torch, naturals_to_ssm_diff

    if entry == "mf_nat_to_ssm":
        a, off, chol = naturals_to_ssm_diff(lin, diag, sub, smoothing=bool(extra[0]))
        a = a.reshape(batch + (t - 1, d, d))
        off = off.reshape(batch + (t, d))
        chol = chol.reshape(batch + (t, d, d))
        return a, off[..., 1:, :], chol[..., 0, :, :], chol[..., 1:, :, :], off[..., 0, :]
    if entry != "mf_expectations_to_ssm":
        raise NotImplementedError(f"{entry} has no reverse mode")

    def cuda_fn(lin_, diag_, sub_):
        a_ = torch.empty_like(sub_)
        off_ = torch.empty_like(lin_)
        chol_ = torch.empty_like(diag_)
        info = torch.empty(lin_.shape[0], dtype=torch.int32, device=lin_.device)
        check(getattr(_lib.lib(), entry)(dtype_code(lin_.dtype), ptr(lin_), ptr(diag_), ptr(sub_), ptr(a_),
                                         ptr(off_), ptr(chol_), ptr(info), i64(lin_.shape[0]), i64(t), i64(d),
                                         *extra, current_stream()), entry)
        _raise_if_failed(info, entry)
        return a_, off_, chol_

    a, off, chol = _RecomputeFn.apply(cuda_fn, expectations_to_ssm_torch, 3, lin, diag, sub)
    a = a.reshape(batch + (t - 1, d, d))
    off = off.reshape(batch + (t, d))
    chol = chol.reshape(batch + (t, d, d))
    return a, off[..., 1:, :], chol[..., 0, :, :], chol[..., 1:, :, :], off[..., 0, :]


def ssm_to_expectations(ssm: StateSpaceModel) -> Tuple[Tensor, Tensor, Tensor]:
    """``(E[x_k], E[x_k x_kᵀ], E[x_{k+1} x_kᵀ])`` (reference :31-89)."""
    mu0, l0, a, b, lq, bsz, t, d = ssm._flat()
    if needs_grad(mu0, l0, a, b, lq):
        # the differentiable moment sweep (adjoint: mf_ssm_marginals_bwd) + the outer products of :79-87
        mean, cov, sub = ssm._marginals(True, True, True)
        mu = mean[..., None]
        return (mean, cov + mu @ mu.transpose(-1, -2),
                sub + mu[..., 1:, :, :] @ mu[..., :-1, :, :].transpose(-1, -2))
    lin = torch.empty(bsz, t, d, dtype=a.dtype, device=a.device)
    diag = torch.empty(bsz, t, d, d, dtype=a.dtype, device=a.device)
    sub = torch.empty(bsz, t - 1, d, d, dtype=a.dtype, device=a.device)
    check(
        _lib.lib().mf_ssm_to_expectations(
            dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(lin), ptr(diag),
            ptr(sub), i64(bsz), i64(t), i64(d), current_stream()),
        "mf_ssm_to_expectations",
    )
    bs = tuple(ssm.batch_shape)
    return lin.reshape(bs + (t, d)), diag.reshape(bs + (t, d, d)), sub.reshape(bs + (t - 1, d, d))


@boundary
def expectations_to_ssm_params(eta_linear, eta_diag, eta_subdiag):
    """Returns ``(As, offsets, chol_P0, chol_Qs, mu0)`` (reference :92-178)."""
    return _ssm_outputs("mf_expectations_to_ssm", eta_linear, eta_diag, eta_subdiag)


def _to_naturals_torch(ssm: StateSpaceModel, smoothing: bool):
    """``ssm_to_naturals`` / ``ssm_to_naturals_no_smoothing`` in differentiable torch ops (per-step maps,
    reference :181-329), used when a parameter requires a gradient."""
    a = ssm._A_s
    d = ssm.state_dim
    offsets = torch.cat([ssm._mu_0[..., None, :], ssm._b_s], dim=-2)[..., None]
    chols = torch.tril(torch.cat([ssm._chol_P_0[..., None, :, :], ssm._chol_Q_s], dim=-3))
    eye = torch.eye(d, dtype=a.dtype, device=a.device).expand(chols.shape)
    prec = torch.cholesky_solve(eye, chols)
    tmp = torch.cholesky_solve(offsets, chols)
    inv_q_a = torch.cholesky_solve(a, chols[..., 1:, :, :])
    if not smoothing:
        return tmp[..., 0], -0.5 * prec, inv_q_a
    lin = torch.cat([tmp[..., :-1, :, :] - a.transpose(-1, -2) @ tmp[..., 1:, :, :], tmp[..., -1:, :, :]], dim=-3)
    aqa = a.transpose(-1, -2) @ inv_q_a
    aqa = torch.cat([aqa, torch.zeros_like(aqa[..., :1, :, :])], dim=-3)
    return lin[..., 0], -0.5 * (prec + aqa), inv_q_a


def _to_naturals(ssm: StateSpaceModel, smoothing: bool):
    mu0, l0, a, b, lq, bsz, t, d = ssm._flat()
    if needs_grad(mu0, l0, a, b, lq):
        return _to_naturals_torch(ssm, smoothing)
    lin = torch.empty(bsz, t, d, dtype=a.dtype, device=a.device)
    diag = torch.empty(bsz, t, d, d, dtype=a.dtype, device=a.device)
    sub = torch.empty(bsz, t - 1, d, d, dtype=a.dtype, device=a.device)
    check(
        _lib.lib().mf_ssm_to_naturals(
            dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(lin), ptr(diag),
            ptr(sub), i64(bsz), i64(t), i64(d), int(smoothing), current_stream()),
        "mf_ssm_to_naturals",
    )
    bs = tuple(ssm.batch_shape)
    return lin.reshape(bs + (t, d)), diag.reshape(bs + (t, d, d)), sub.reshape(bs + (t - 1, d, d))


@boundary
def ssm_to_naturals(ssm: StateSpaceModel) -> Tuple[Tensor, Tensor, Tensor]:
    """``(θ_lin, θ_diag, θ_sub)`` (reference :181-253)."""
    return _to_naturals(ssm, True)


@boundary
def ssm_to_naturals_no_smoothing(ssm: StateSpaceModel) -> Tuple[Tensor, Tensor, Tensor]:
    """Reference :256-329."""
    return _to_naturals(ssm, False)


@boundary
def naturals_to_ssm_params(theta_linear, theta_diag, theta_subdiag):
    """Returns ``(As, offsets, chol_P0, chol_Qs, mu0)`` (reference :332-511)."""
    return _ssm_outputs("mf_nat_to_ssm", theta_linear, theta_diag, theta_subdiag, 1)


@boundary
def naturals_to_ssm_params_no_smoothing(theta_linear, theta_diag, theta_subdiag):
    """Reference :514-593."""
    return _ssm_outputs("mf_nat_to_ssm", theta_linear, theta_diag, theta_subdiag, 0)
