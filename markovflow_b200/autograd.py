"""Reverse-mode differentiation of the operators.

The reference obtains gradients from TensorFlow's tape: the banded ops register gradients in
``banded-matrices`` and everything else is TF ops (callers: ``ssm_natgrad.py:142-172``,
``tests/integration/models/test_gaussian_process_regression.py:117-130``).  Here the sequential
recurrences are ``torch.autograd.Function`` s whose backward passes are adjoint CUDA sweeps behind the C
ABI (``mf_btd_cholesky_bwd``, ``mf_ssm_marginals_bwd``; triangular solves are their own adjoints with
the transpose flag flipped), and the per-step maps (``_build_precision``, ``expectations_to_ssm_params``)
run their CUDA kernel forward and differentiate a torch restatement of the same map backward.
Composite quantities -- ``KalmanFilter.log_likelihood``, ``kl_divergence`` -- are assembled from these
primitives with the reference's own formulas (``kalman_filter.py:184-255``,
``state_space_model.py:528-593``) whenever an input requires a gradient; otherwise the fused
forward-only kernels run.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr


def needs_grad(*tensors) -> bool:
    """True when autograd is recording and any operand requires a gradient."""
    return torch.is_grad_enabled() and any(
        isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


# ---------------------------------------------------------------------------------------------------
# SymmetricBlockTriDiagonal.cholesky  (block_tri_diag.py:423-436)
# ---------------------------------------------------------------------------------------------------
class CholeskyFn(torch.autograd.Function):
    """``(diag [B,T,D,D], sub [B,T-1,D,D] | None) -> (Ld, Ls | None, info)``."""

    @staticmethod
    def forward(ctx, diag, sub):
        diag, sub = _c(diag.detach()), _c(None if sub is None else sub.detach())
        b, t, d, _ = diag.shape
        ld = torch.empty_like(diag)
        ls = torch.empty_like(sub) if sub is not None else None
        info = torch.empty(b, dtype=torch.int32, device=diag.device)
        check(_lib.lib().mf_btd_cholesky(dtype_code(diag.dtype), ptr(diag), ptr(sub), None, ptr(ld), ptr(ls),
                                         None, None, ptr(info), i64(b), i64(t), i64(d), current_stream()),
              "mf_btd_cholesky")
        ctx.save_for_backward(ld, ls)
        ctx.has_sub = sub is not None
        ctx.mark_non_differentiable(info)
        if ls is None:
            return ld, None, info
        return ld, ls, info

    @staticmethod
    def backward(ctx, g_ld, g_ls, _g_info):
        ld, ls = ctx.saved_tensors
        b, t, d, _ = ld.shape
        g_ld, g_ls = _c(g_ld), _c(g_ls) if ctx.has_sub else None
        g_diag = torch.empty_like(ld)
        g_sub = torch.empty_like(ls) if ctx.has_sub else None
        check(_lib.lib().mf_btd_cholesky_bwd(dtype_code(ld.dtype), ptr(ld), ptr(ls), ptr(g_ld), ptr(g_ls),
                                             ptr(g_diag), ptr(g_sub), i64(b), i64(t), i64(d), current_stream()),
              "mf_btd_cholesky_bwd")
        return g_diag, g_sub


# ---------------------------------------------------------------------------------------------------
# LowerTriangularBlockTriDiagonal.solve  (block_tri_diag.py:339-351)
# ---------------------------------------------------------------------------------------------------
def _solve_raw(ld, ls, rhs, transpose: bool) -> torch.Tensor:
    n, t, d = rhs.shape
    bm = (ld if ld is not None else ls).shape[0] if (ld is not None or ls is not None) else 1
    out = torch.empty_like(rhs)
    check(_lib.lib().mf_btd_solve(dtype_code(rhs.dtype), ptr(ld), ptr(ls), ptr(rhs), ptr(out), i64(n), i64(bm),
                                  i64(t), i64(d), int(bool(transpose)), current_stream()), "mf_btd_solve")
    return out


class SolveFn(torch.autograd.Function):
    """``x = L^-1 rhs`` (or ``L^-T rhs``): ``ld [Bm,T,D,D] | None`` (None = identity diagonal blocks),
    ``ls [Bm,T-1,D,D] | None``, ``rhs [n,T,D]`` with ``n`` a multiple of ``Bm`` (row c uses matrix c % Bm)."""

    @staticmethod
    def forward(ctx, ld, ls, rhs, transpose: bool):
        ld, ls, rhs = (_c(None if x is None else x.detach()) for x in (ld, ls, rhs))
        x = _solve_raw(ld, ls, rhs, transpose)
        ctx.save_for_backward(*(v for v in (ld, ls, x) if v is not None))
        ctx.flags = (ld is not None, ls is not None, bool(transpose))
        return x

    @staticmethod
    def backward(ctx, g_x):
        has_ld, has_ls, transpose = ctx.flags
        saved = list(ctx.saved_tensors)
        ld = saved.pop(0) if has_ld else None
        ls = saved.pop(0) if has_ls else None
        x = saved.pop(0)
        g = _solve_raw(ld, ls, _c(g_x), not transpose)  # adjoint of the right-hand side
        n, t, d = x.shape
        bm = (ld if has_ld else ls).shape[0] if (has_ld or has_ls) else 1
        u, v = (x, g) if transpose else (g, x)  # L_bar = -(u v^T) on the block pattern
        u4, v4 = u.reshape(n // bm, bm, t, d), v.reshape(n // bm, bm, t, d)
        g_ld = g_ls = None
        if has_ld and ctx.needs_input_grad[0]:
            g_ld = -torch.tril(torch.einsum("sbti,sbtj->btij", u4, v4))
        if has_ls and ctx.needs_input_grad[1]:
            g_ls = -torch.einsum("sbti,sbtj->btij", u4[:, :, 1:], v4[:, :, :-1])
        return g_ld, g_ls, g, None


# ---------------------------------------------------------------------------------------------------
# marginal means / covariances / lag-one covariances  (state_space_model.py:231-275, 326-341)
# ---------------------------------------------------------------------------------------------------
class MarginalsFn(torch.autograd.Function):
    """``(mu0 [B,D], chol_p0 [B,D,D], a [B,T-1,D,D], b [B,T-1,D], chol_q [B,T-1,D,D]) ->
    (mean [B,T,D], cov [B,T,D,D], sub [B,T-1,D,D] = a_k cov_k)``."""

    @staticmethod
    def forward(ctx, mu0, l0, a, b, lq):
        mu0, l0, a, b, lq = (_c(x.detach()) for x in (mu0, l0, a, b, lq))
        bsz, n, d, _ = a.shape
        t = n + 1
        mean = torch.empty(bsz, t, d, dtype=a.dtype, device=a.device)
        cov = torch.empty(bsz, t, d, d, dtype=a.dtype, device=a.device)
        sub = torch.empty(bsz, n, d, d, dtype=a.dtype, device=a.device)
        check(_lib.lib().mf_ssm_marginals(dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(mean),
                                          ptr(cov), ptr(sub), i64(bsz), i64(t), i64(d), current_stream()),
              "mf_ssm_marginals")
        ctx.save_for_backward(l0, a, lq, mean, cov)
        return mean, cov, sub

    @staticmethod
    def backward(ctx, g_mean, g_cov, g_sub):
        l0, a, lq, mean, cov = ctx.saved_tensors
        bsz, n, d, _ = a.shape
        g_mean, g_cov, g_sub = _c(g_mean), _c(g_cov), _c(g_sub)
        g_mu0 = torch.empty(bsz, d, dtype=a.dtype, device=a.device)
        g_l0, g_a, g_lq = torch.empty_like(l0), torch.empty_like(a), torch.empty_like(lq)
        g_b = torch.empty(bsz, n, d, dtype=a.dtype, device=a.device)
        check(_lib.lib().mf_ssm_marginals_bwd(dtype_code(a.dtype), ptr(l0), ptr(a), ptr(lq), ptr(mean), ptr(cov),
                                              ptr(g_mean), ptr(g_cov), ptr(g_sub), ptr(g_mu0), ptr(g_l0), ptr(g_a),
                                              ptr(g_b), ptr(g_lq), i64(bsz), i64(n + 1), i64(d), current_stream()),
              "mf_ssm_marginals_bwd")
        return g_mu0, g_l0, g_a, g_b, g_lq


# ---------------------------------------------------------------------------------------------------
# per-step maps: CUDA kernel forward, torch restatement differentiated backward
# ---------------------------------------------------------------------------------------------------
def precision_blocks_torch(l0, a, lq, h=None, r_inv=None):
    """``_build_precision`` (+ ``H^T R^-1 H``) in torch ops (``state_space_model.py:431-483``,
    ``kalman_filter.py:85-101``): the differentiable restatement of ``mf_ssm_build_precision``."""
    bsz, n, d, _ = a.shape
    eye = torch.eye(d, dtype=a.dtype, device=a.device)
    l0, lq = torch.tril(l0), torch.tril(lq)  # only the lower triangles are parameters (and are read)
    inv_q_a = torch.cholesky_solve(a, lq)
    aqa = a.transpose(-1, -2) @ inv_q_a
    chols = torch.cat([l0[:, None], lq], dim=1)
    diag = torch.cholesky_solve(eye.expand(bsz, n + 1, d, d), chols)
    diag = torch.cat([diag[:, :-1] + aqa, diag[:, -1:]], dim=1)
    if h is not None:
        hrh = h.transpose(-1, -2) @ (r_inv @ h)  # [hb,T,D,D] or [T,D,D] broadcasts over the batch
        diag = diag + hrh
    return diag, -inv_q_a


class _RecomputeFn(torch.autograd.Function):
    """Runs ``cuda_fn`` forward; backward differentiates ``torch_fn`` (the same map in torch ops)."""

    @staticmethod
    def forward(ctx, cuda_fn, torch_fn, n_out, *inputs):
        ctx.torch_fn = torch_fn
        ctx.save_for_backward(*inputs)
        with torch.no_grad():
            outs = cuda_fn(*(x.detach() for x in inputs))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        inputs = [x.detach().requires_grad_(True) for x in ctx.saved_tensors]
        with torch.enable_grad():
            outs = ctx.torch_fn(*inputs)
        pairs = [(o, g) for o, g in zip(outs, grads) if g is not None]
        gin = torch.autograd.grad([o for o, _ in pairs], inputs, [g for _, g in pairs], allow_unused=True)
        return (None, None, None) + tuple(gin)


def expectations_to_ssm_torch(eta_lin, eta_diag, eta_sub):
    """``expectations_to_ssm_params`` in torch ops (``ssm_gaussian_transformations.py:92-178``); returns the
    concatenated layout of ``mf_expectations_to_ssm``: ``(a [B,T-1,D,D], offsets [B,T,D], chols [B,T,D,D])``."""
    eta = eta_lin[..., None]
    covs = eta_diag - eta @ eta.transpose(-1, -2)
    covs = torch.tril(covs) + torch.tril(covs, -1).transpose(-1, -2)  # the kernel reads lower triangles
    covs_sub = eta_sub.transpose(-1, -2) - eta[:, :-1] @ eta[:, 1:].transpose(-1, -2)
    chols = torch.linalg.cholesky(covs)
    a = torch.cholesky_solve(covs_sub, chols[:, :-1]).transpose(-1, -2)
    offsets = (eta[:, 1:] - a @ eta[:, :-1])[..., 0]
    cond = covs[:, 1:] - a @ covs[:, :-1] @ a.transpose(-1, -2)
    cond = torch.tril(cond) + torch.tril(cond, -1).transpose(-1, -2)
    chol_q = torch.linalg.cholesky(cond)
    return (a, torch.cat([eta_lin[:, :1], offsets], dim=1), torch.cat([chols[:, :1], chol_q], dim=1))


def _sym_lower(m: torch.Tensor) -> torch.Tensor:
    """The symmetric matrix the kernels see when they read lower triangles."""
    return torch.tril(m) + torch.tril(m, -1).transpose(-1, -2)


def naturals_to_ssm_diff(theta_lin, theta_diag, theta_sub, smoothing: bool = True):
    """Differentiable ``naturals_to_ssm_params`` (``ssm_gaussian_transformations.py:332-593``) in the concatenated
    layout of ``mf_nat_to_ssm``: ``(a [B,T-1,D,D], offsets [B,T,D], chols [B,T,D,D])``.

    The backward ``U D U^T`` recursion ``D_k = P_kk - P_{k+1,k}^T D_{k+1}^{-1} P_{k+1,k}``,
    ``z_k = theta_k + A_k^T z_{k+1}`` IS the block Cholesky + forward substitution of the TIME-REVERSED
    precision: with ``C_k = chol(D_k)`` the reversed factor has diagonal blocks ``C_k`` and its forward solve
    gives ``x_k = C_k^{-1} z_k``.  So the sequential part runs on the CUDA sweeps that have adjoint sweeps
    (``CholeskyFn``, ``SolveFn``); what is left are per-step maps in batched torch ops:
    ``A_k = D_{k+1}^{-1} theta_sub_k``, ``offset_k = C_k^{-T} x_k``, ``chol_k = chol(D_k^{-1})``."""
    diag = _sym_lower(-2.0 * theta_diag)
    if not smoothing:  # per-step map (:514-593)
        c = torch.linalg.cholesky(diag)
        off = torch.cholesky_solve(theta_lin[..., None], c)[..., 0]
        a = torch.cholesky_solve(theta_sub, c[:, 1:])
        chols = torch.linalg.cholesky(_sym_lower(torch.cholesky_inverse(c)))
        return a, off, chols
    t = theta_lin.shape[1]
    if t == 1:
        sub_r = None
    else:
        sub_r = torch.flip(-theta_sub.transpose(-1, -2), dims=(1,)).contiguous()  # P_{k+1,k}^T, reversed
    ld_r, ls_r, _ = CholeskyFn.apply(torch.flip(diag, dims=(1,)).contiguous(), sub_r)
    x_r = SolveFn.apply(ld_r, ls_r, torch.flip(theta_lin, dims=(1,)).contiguous(), False)
    c = torch.tril(torch.flip(ld_r, dims=(1,)))  # C_k = chol(D_k)
    x = torch.flip(x_r, dims=(1,))
    off = torch.linalg.solve_triangular(c.transpose(-1, -2), x[..., None], upper=True)[..., 0]
    a = torch.cholesky_solve(theta_sub, c[:, 1:]) if t > 1 else theta_sub
    chols = torch.linalg.cholesky(_sym_lower(torch.cholesky_inverse(c)))
    return a, off, chols


# ---------------------------------------------------------------------------------------------------
# composites, assembled from the primitives with the reference's formulas
# ---------------------------------------------------------------------------------------------------
def _precision(l0, a, lq, h=None, r_inv=None):
    from .state_space_model import _precision_blocks_cuda

    if h is None:
        return _RecomputeFn.apply(lambda *x: _precision_blocks_cuda(*x, None, None),
                                  lambda *x: precision_blocks_torch(*x), 2, l0, a, lq)
    return _RecomputeFn.apply(_precision_blocks_cuda, precision_blocks_torch, 2, l0, a, lq, h, r_inv)


def kalman_log_likelihood_diff(mu0, l0, a, b, lq, h, y, chol_r) -> torch.Tensor:
    """Differentiable ``BaseKalmanFilter.log_likelihood`` per chain, the reference's SpInGP form
    (``kalman_filter.py:184-255``): flat operands ``a [B,T-1,D,D]`` ..., ``h [1|B,T,m,D]``, ``y [B,T,m]``,
    ``chol_r [m,m]`` or ``[T,m,m]``.  Returns ``[B]``."""
    bsz, n, d, _ = a.shape
    t, m = n + 1, h.shape[-2]
    eye_m = torch.eye(m, dtype=a.dtype, device=a.device)
    chol_r = torch.tril(chol_r)
    r_inv = torch.cholesky_solve(eye_m.expand(chol_r.shape), chol_r)  # [m,m] or [T,m,m]
    diag, sub = _precision(l0, a, lq, h, r_inv)
    ld, ls, info = CholeskyFn.apply(diag, sub)
    mean, _, _ = MarginalsFn.apply(mu0, l0, a, b, lq)
    disp = y - (h @ mean[..., None])[..., 0]
    rd = (r_inv @ disp[..., None])[..., 0]
    term1 = -0.5 * torch.sum(rd * disp, dim=(-1, -2))
    obs_proj = (h.transpose(-1, -2) @ rd[..., None])[..., 0]
    z = SolveFn.apply(ld, ls, obs_proj.contiguous(), False)
    term2 = 0.5 * torch.sum(z * z, dim=(-1, -2))
    log_det_prior = -2.0 * (torch.log(torch.diagonal(l0, dim1=-2, dim2=-1).abs()).sum(-1)
                            + torch.log(torch.diagonal(lq, dim1=-2, dim2=-1).abs()).sum((-1, -2)))
    log_det_l = torch.log(torch.diagonal(ld, dim1=-2, dim2=-1).abs()).sum((-1, -2))
    ldr = -2.0 * torch.log(torch.diagonal(chol_r, dim1=-2, dim2=-1).abs()).sum(-1)  # log|R^-1| per step
    log_det_obs = t * ldr if chol_r.dim() == 2 else ldr.sum(-1)
    cst = -0.5 * math.log(2.0 * math.pi) * (m * t)
    return cst + term1 + term2 + 0.5 * log_det_prior - log_det_l + 0.5 * log_det_obs, info


def kl_divergence_diff(q, p) -> torch.Tensor:
    """Differentiable ``KL(q || p)`` per chain for two flat SSM parameter tuples ``(mu0, l0, a, b, lq)``, by the
    reference's closed form (``state_space_model.py:528-593``):
    ``1/2 [tr(K_p^-1 Sigma_q) + (mu_q-mu_p)^T K_p^-1 (mu_q-mu_p) - N - log|K_p^-1| + log|K_q^-1|]``
    where the trace needs the block-tridiagonal part of ``Sigma_q`` only."""
    mean_q, cov_q, sub_q = MarginalsFn.apply(*q)
    mean_p, _, _ = MarginalsFn.apply(*p)
    diag_p, sub_p = _precision(p[1], p[2], p[4])
    bsz, t, d = mean_q.shape
    sym_p = torch.tril(diag_p) + torch.tril(diag_p, -1).transpose(-1, -2)
    trace = torch.sum(sym_p * cov_q, dim=(-1, -2, -3)) + 2.0 * torch.sum(sub_p * sub_q, dim=(-1, -2, -3))
    dm = (mean_q - mean_p)[..., None]
    kd = sym_p @ dm
    kd = kd + torch.cat([torch.zeros_like(kd[:, :1]), sub_p @ dm[:, :-1]], dim=1)
    kd = kd + torch.cat([sub_p.transpose(-1, -2) @ dm[:, 1:], torch.zeros_like(kd[:, :1])], dim=1)
    maha = torch.sum(dm * kd, dim=(-1, -2, -3))

    def log_det_prec(l0, lq):
        return -2.0 * (torch.log(torch.diagonal(l0, dim1=-2, dim2=-1).abs()).sum(-1)
                       + torch.log(torch.diagonal(lq, dim1=-2, dim2=-1).abs()).sum((-1, -2)))

    return 0.5 * (trace + maha - t * d - log_det_prec(p[1], p[4]) + log_det_prec(q[1], q[4]))
