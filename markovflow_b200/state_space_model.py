"""``StateSpaceModel`` with the API of ``markovflow/state_space_model.py:35-664``, executed by the
sm_100a kernels behind the C ABI (``mf_ssm_*`` in ``include/markovflow_b200.h``).

The model is ``x₀ ~ N(μ₀, P₀)``, ``x_{k+1} = A_k x_k + b_k + q_k``, ``q_k ~ N(0, Q_k)``, parameterised
by ``μ₀``, ``chol P₀``, ``A_k``, ``b_k``, ``chol Q_k`` with an arbitrary leading batch shape.
Differences from the reference, by design: tensors are CUDA tensors; marginal covariances come
from the forward recursion ``P_{k+1} = A_k P_k A_kᵀ + Q_k`` in one sweep (the reference factors the
precision and takes the sparse inverse subset -- same quantity, three sweeps and less accurate);
``kl_divergence`` is one fused sweep over both models.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CholeskyError, check, current_stream, dtype_code, i64, ptr
from .config import check_numerics
from .block_tri_diag import LowerTriangularBlockTriDiagonal, SymmetricBlockTriDiagonal, _prod
from .gauss_markov import GaussMarkovDistribution, check_compatible
from .autograd import MarginalsFn, needs_grad
from .interop import framework_of, as_torch, boundary, require_cuda


class StateSpaceModel(GaussMarkovDistribution):
    """Reference ``state_space_model.py:35-609``."""

    def __init__(self, initial_mean, chol_initial_covariance, state_transitions, state_offsets,
                 chol_process_covariances) -> None:
        self._fw = framework_of(initial_mean, chol_initial_covariance, state_transitions, state_offsets,
                                chol_process_covariances)
        mu0 = as_torch(initial_mean)
        dev = mu0.device
        l0 = as_torch(chol_initial_covariance, dev)
        a = as_torch(state_transitions, dev)
        b = as_torch(state_offsets, dev)
        lq = as_torch(chol_process_covariances, dev)
        if a.dim() < 3 or a.shape[-1] != a.shape[-2]:
            raise ValueError("state_transitions must be [..., num_transitions, state_dim, state_dim]")
        d, n = int(a.shape[-1]), int(a.shape[-3])
        batch = tuple(a.shape[:-3])
        # exact batch-shape match, as the reference asserts (state_space_model.py:101-116)
        want = {
            "initial_mean": (mu0, batch + (d,)),
            "chol_initial_covariance": (l0, batch + (d, d)),
            "state_offsets": (b, batch + (n, d)),
            "chol_process_covariances": (lq, batch + (n, d, d)),
        }
        for name, (t, shape) in want.items():
            if tuple(t.shape) != shape:
                raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {shape}")
            if t.dtype != a.dtype:
                raise ValueError(f"{name} must have dtype {a.dtype}")
        if n == 0:
            raise ValueError("a StateSpaceModel needs at least one transition")
        self._mu_0, self._chol_P_0, self._A_s, self._b_s, self._chol_Q_s = mu0, l0, a, b, lq

    @property
    def _dev(self) -> torch.device:
        return self._A_s.device

    # -- shapes / accessors (reference :126-229) -------------------------------------------------
    @property
    def event_shape(self) -> torch.Size:
        return torch.Size((self.num_transitions + 1, self.state_dim))

    @property
    def batch_shape(self) -> torch.Size:
        return self._A_s.shape[:-3]

    @property
    def state_dim(self) -> int:
        return int(self._A_s.shape[-2])

    @property
    def num_transitions(self) -> int:
        return int(self._A_s.shape[-3])

    @property
    @boundary
    def cholesky_process_covariances(self) -> torch.Tensor:
        return self._chol_Q_s

    @property
    @boundary
    def cholesky_initial_covariance(self) -> torch.Tensor:
        return self._chol_P_0

    @property
    @boundary
    def initial_covariance(self) -> torch.Tensor:
        return self._chol_P_0 @ self._chol_P_0.transpose(-1, -2)

    @property
    @boundary
    def concatenated_cholesky_process_covariance(self) -> torch.Tensor:
        return torch.cat([self._chol_P_0[..., None, :, :], self._chol_Q_s], dim=-3)

    @property
    @boundary
    def state_offsets(self) -> torch.Tensor:
        return self._b_s

    @property
    @boundary
    def initial_mean(self) -> torch.Tensor:
        return self._mu_0

    @property
    @boundary
    def concatenated_state_offsets(self) -> torch.Tensor:
        return torch.cat([self._mu_0[..., None, :], self._b_s], dim=-2)

    @property
    @boundary
    def state_transitions(self) -> torch.Tensor:
        return self._A_s

    # -- flattened contiguous parameter views for the C ABI ---------------------------------------
    def _flat(self):
        """Dense ``[B, ...]`` views of the parameters for the C ABI.  Parameters that arrive as strided
        views (``naturals_to_ssm_params`` returns slices of its concatenated outputs) are compacted
        ONCE per object -- the parameters are immutable, as in the reference -- not once per method."""
        flat = getattr(self, "_flat_cache", None)
        if flat is None:
            flat = self._flat_cache = self._flatten()
        return flat

    def _flatten(self):
        require_cuda(self._A_s, "state_transitions")
        bsz, n, d = _prod(self.batch_shape), self.num_transitions, self.state_dim
        return (
            self._mu_0.reshape(bsz, d).contiguous(),
            self._chol_P_0.reshape(bsz, d, d).contiguous(),
            self._A_s.reshape(bsz, n, d, d).contiguous(),
            self._b_s.reshape(bsz, n, d).contiguous(),
            self._chol_Q_s.reshape(bsz, n, d, d).contiguous(),
            bsz, n + 1, d,
        )

    def _marginals(self, mean: bool, cov: bool, sub: bool):
        mu0, l0, a, b, lq, bsz, t, d = self._flat()
        if needs_grad(mu0, l0, a, b, lq):
            bs = tuple(self.batch_shape)
            o_mean, o_cov, o_sub = MarginalsFn.apply(mu0, l0, a, b, lq)
            return (o_mean.reshape(bs + (t, d)) if mean else None, o_cov.reshape(bs + (t, d, d)) if cov else None,
                    o_sub.reshape(bs + (t - 1, d, d)) if sub else None)
        o_mean = torch.empty(bsz, t, d, dtype=a.dtype, device=a.device) if mean else None
        o_cov = torch.empty(bsz, t, d, d, dtype=a.dtype, device=a.device) if cov else None
        o_sub = torch.empty(bsz, t - 1, d, d, dtype=a.dtype, device=a.device) if sub else None
        check(
            _lib.lib().mf_ssm_marginals(
                dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(o_mean),
                ptr(o_cov), ptr(o_sub), i64(bsz), i64(t), i64(d), current_stream()),
            "mf_ssm_marginals",
        )
        bs = tuple(self.batch_shape)
        return (
            None if o_mean is None else o_mean.reshape(bs + (t, d)),
            None if o_cov is None else o_cov.reshape(bs + (t, d, d)),
            None if o_sub is None else o_sub.reshape(bs + (t - 1, d, d)),
        )

    # -- moments (reference :231-275, :326-341) ----------------------------------------------------
    @property
    @boundary
    def marginal_means(self) -> torch.Tensor:
        if needs_grad(*self._flat()[:5]):
            return self._marginals(True, False, False)[0]
        return self._affine(None, ())

    @property
    @boundary
    def marginal_covariances(self) -> torch.Tensor:
        return self._marginals(False, True, False)[1]

    @property
    @boundary
    def marginals(self) -> Tuple[torch.Tensor, torch.Tensor]:
        mean, cov, _ = self._marginals(True, True, False)
        return mean, cov

    @boundary
    def covariance_blocks(self) -> Tuple[torch.Tensor, torch.Tensor]:
        _, cov, sub = self._marginals(False, True, True)
        return cov, sub

    @boundary
    def subsequent_covariances(self, marginal_covariances) -> torch.Tensor:
        """``Σ_{k+1,k} = A_k Σ_kk`` (reference :326-341)."""
        cov = as_torch(marginal_covariances, self._A_s.device)
        return self._A_s @ cov[..., :-1, :, :]

    @property
    @boundary
    def a_inv_block(self) -> LowerTriangularBlockTriDiagonal:
        """``A⁻¹``: identity diagonal, ``-A_k`` sub-diagonal (reference :277-296)."""
        d, t = self.state_dim, self.num_transitions + 1
        eye = torch.eye(d, dtype=self._A_s.dtype, device=self._A_s.device)
        eye = eye.expand(tuple(self.batch_shape) + (t, d, d))
        return LowerTriangularBlockTriDiagonal(eye, -self._A_s, unit_diagonal=True)

    def _affine(self, eps: Optional[torch.Tensor], sample_shape) -> torch.Tensor:
        mu0, l0, a, b, lq, bsz, t, d = self._flat()
        if eps is not None and needs_grad(mu0, l0, a, b, lq, eps):
            return self._affine_diff(eps, sample_shape)
        n = _prod(sample_shape) * bsz
        out = torch.empty(n, t, d, dtype=a.dtype, device=a.device)
        if eps is not None:
            eps = eps.reshape(n, t, d).contiguous()
        check(
            _lib.lib().mf_ssm_affine_scan(
                dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(eps), ptr(out),
                i64(n), i64(bsz), i64(t), i64(d), current_stream()),
            "mf_ssm_affine_scan",
        )
        return out.reshape(tuple(sample_shape) + tuple(self.batch_shape) + (t, d))

    def _affine_diff(self, eps: torch.Tensor, sample_shape) -> torch.Tensor:
        """Reparameterised trajectories with reverse mode (reference :298-324 under a gradient tape):
        ``x = A_inv^-1 [mu0 + L0 e_0, b_k + Lq_k e_k]`` -- the right-hand side in torch ops, the unit-diagonal
        block solve on the CUDA sweep with its adjoint sweep (``autograd.SolveFn``)."""
        from .autograd import SolveFn

        mu0, l0, a, b, lq, bsz, t, d = self._flat()
        s = _prod(sample_shape)
        e = eps.reshape(s, bsz, t, d)
        first = mu0 + (torch.tril(l0) @ e[:, :, 0, :, None])[..., 0]
        if t > 1:
            rest = b + (torch.tril(lq) @ e[:, :, 1:, :, None])[..., 0]
            rhs = torch.cat([first[:, :, None, :], rest], dim=2)
        else:
            rhs = first[:, :, None, :]
        x = SolveFn.apply(None, -a if t > 1 else None, rhs.reshape(s * bsz, t, d).contiguous(), False)
        return x.reshape(tuple(sample_shape) + tuple(self.batch_shape) + (t, d))

    @boundary
    def sample(self, sample_shape, generator: Optional[torch.Generator] = None,
               seed: Optional[int] = None) -> torch.Tensor:
        """Trajectories ``sample_shape + batch_shape + [T, D]`` (reference :298-324).

        By default the standard normals are drawn INSIDE the sweep (``mf_ssm_sample``: Philox4x32-10 keyed by
        ``seed``, trajectory and step -- nothing is written or read for them); ``seed`` defaults to a draw from
        torch's global CPU generator, so ``torch.manual_seed`` makes samples reproducible.
        :meth:`sample_epsilons` returns the same stream.  With a ``generator`` the draws come from
        ``torch.randn`` and are streamed through ``mf_ssm_affine_scan``."""
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        sample_shape = tuple(int(s) for s in sample_shape)
        full = sample_shape + tuple(self.batch_shape) + tuple(self.event_shape)
        if generator is not None:
            eps = torch.randn(full, dtype=self._A_s.dtype, device=self._A_s.device, generator=generator)
            if eps.numel() == 0:
                return eps
            return self._affine(eps, sample_shape)
        mu0, l0, a, b, lq, bsz, t, d = self._flat()
        if needs_grad(mu0, l0, a, b, lq):
            # reverse mode: the same Philox draws, written out once, then the differentiable affine solve
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64).item())
            return self._affine_diff(self.sample_epsilons(sample_shape, seed), sample_shape)
        n = _prod(sample_shape) * bsz
        out = torch.empty(n, t, d, dtype=a.dtype, device=a.device)
        if n == 0:
            return out.reshape(full)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64).item())
        check(
            _lib.lib().mf_ssm_sample(
                dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), _lib.ctypes.c_uint64(int(seed)),
                ptr(out), i64(n), i64(bsz), i64(t), i64(d), current_stream()),
            "mf_ssm_sample",
        )
        return out.reshape(full)

    def sample_epsilons(self, sample_shape, seed: int) -> torch.Tensor:
        """The standard normals :meth:`sample` draws in its kernel for ``seed``, written out
        (``mf_philox_normal``): ``sample_from_epsilons(sample_epsilons(shape, seed)) == sample(shape, seed=seed)``."""
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        sample_shape = tuple(int(s) for s in sample_shape)
        t, d = self.num_transitions + 1, self.state_dim
        n = _prod(sample_shape) * _prod(self.batch_shape)
        out = torch.empty(n, t, d, dtype=self._A_s.dtype, device=self._A_s.device)
        check(
            _lib.lib().mf_philox_normal(dtype_code(out.dtype), _lib.ctypes.c_uint64(int(seed)), ptr(out), i64(n),
                                        i64(t), i64(d), current_stream()),
            "mf_philox_normal",
        )
        return out.reshape(sample_shape + tuple(self.batch_shape) + (t, d))

    @boundary
    def sample_from_epsilons(self, epsilons) -> torch.Tensor:
        """:meth:`sample` with the standard-normal draw supplied (``[..., batch, T, D]``)."""
        eps = as_torch(epsilons, self._A_s.device)
        nb = len(self.batch_shape) + 2
        sample_shape = tuple(eps.shape[: eps.dim() - nb])
        return self._affine(eps, sample_shape)

    @boundary
    def log_det_precision(self) -> torch.Tensor:
        """``-log|P₀| - Σ log|Q_k|`` (reference :343-373)."""
        l0 = LowerTriangularBlockTriDiagonal(self._chol_P_0[..., None, :, :])
        lq = LowerTriangularBlockTriDiagonal(self._chol_Q_s)
        return -2.0 * (l0.abs_log_det() + lq.abs_log_det())

    @boundary
    def create_non_trainable_copy(self) -> "StateSpaceModel":
        return StateSpaceModel(*(t.detach() for t in (
            self._mu_0, self._chol_P_0, self._A_s, self._b_s, self._chol_Q_s)))

    @boundary
    def create_trainable_copy(self) -> "StateSpaceModel":
        """Copy whose parameters are fresh leaf tensors (reference :396-429; the reference wraps
        them in ``gpflow.Parameter`` with a triangular bijector, which is optimiser plumbing)."""
        return StateSpaceModel(*(t.detach().clone().requires_grad_(True) for t in (
            self._mu_0, self._chol_P_0, self._A_s, self._b_s, self._chol_Q_s)))

    @property
    def trainable_variables(self):
        """``(A_s, b_s, chol_P_0, chol_Q_s, mu_0)`` -- the order ``ssm_natgrad.py:159-166`` unpacks."""
        return (self._A_s, self._b_s, self._chol_P_0, self._chol_Q_s, self._mu_0)

    @boundary
    def _build_precision(self) -> SymmetricBlockTriDiagonal:
        """``K⁻¹`` blocks (reference :431-483)."""
        diag, sub = self._precision_blocks(None, None)
        return SymmetricBlockTriDiagonal(diag, sub)

    def _precision_blocks(self, h: Optional[torch.Tensor], r_inv: Optional[torch.Tensor]):
        """Precision blocks, optionally fused with ``+ HᵀR⁻¹H`` (``kalman_filter.py:85-101``).
        ``h``: ``[T,m,D]`` or ``batch + [T,m,D]``; ``r_inv``: ``[m,m]`` or ``[T,m,m]``."""
        mu0, l0, a, b, lq, bsz, t, d = self._flat()
        m, hb, rs = 0, 1, 1
        if h is not None:
            # same operand handling as kalman_log_likelihood: dtype of the model, emission batch
            # expanded to the model's batch, shapes checked before any pointer is handed over
            h = as_torch(h, a.device)
            r_inv = as_torch(r_inv, a.device)
            m = int(h.shape[-2])
            if tuple(h.shape[-3:]) != (t, m, d):
                raise ValueError(f"emission matrix must be [..., {t}, m, {d}], got {tuple(h.shape)}")
            if tuple(r_inv.shape) not in ((m, m), (t, m, m)):
                raise ValueError("observation precision must be [m, m] or [T, m, m]")
            hb = 1 if h.dim() == 3 else bsz
            if h.dim() > 3 and tuple(h.shape[:-3]) != tuple(self.batch_shape):
                h = h.expand(tuple(self.batch_shape) + (t, m, d))
            h = h.reshape(hb, t, m, d).contiguous().to(a.dtype)
            rs = 1 if r_inv.dim() == 2 else t
            r_inv = r_inv.reshape(rs, m, m).contiguous().to(a.dtype)
        if needs_grad(l0, a, lq, h, r_inv):
            from .autograd import _precision

            diag, sub = _precision(l0, a, lq, h, r_inv)
        else:
            diag, sub = _precision_blocks_cuda(l0, a, lq, h, r_inv)
        bs = tuple(self.batch_shape)
        return diag.reshape(bs + (t, d, d)), sub.reshape(bs + (t - 1, d, d))

    @boundary
    def log_pdf(self, states) -> torch.Tensor:
        """``log p(states)``: ``[..., batch, T, D] -> [..., batch]`` (reference :485-526)."""
        mu0, l0, a, b, lq, bsz, t, d = self._flat()
        x = as_torch(states, a.device)
        if needs_grad(mu0, l0, a, b, lq, x):
            return self._log_pdf_torch(x)
        nb = len(self.batch_shape)
        if x.dim() < nb + 2 or tuple(x.shape[x.dim() - nb - 2:]) != tuple(self.batch_shape) + (t, d):
            raise ValueError(
                f"states must have shape [...] + {tuple(self.batch_shape) + (t, d)}, got {tuple(x.shape)}")
        lead = tuple(x.shape[: x.dim() - nb - 2])
        n = _prod(lead) * bsz
        x = x.reshape(n, t, d).contiguous()
        out = torch.empty(n, dtype=a.dtype, device=a.device)
        check(
            _lib.lib().mf_ssm_log_pdf(
                dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(x), ptr(out),
                i64(n), i64(bsz), i64(t), i64(d), current_stream()),
            "mf_ssm_log_pdf",
        )
        return out.reshape(lead + tuple(self.batch_shape))

    @boundary
    def _log_pdf_torch(self, x: torch.Tensor) -> torch.Tensor:
        """``log_pdf`` in differentiable torch ops (per-step Gaussian factors, no recursion), used when an
        operand requires a gradient (reference ``_log_pdf_factors`` :485-526)."""
        def mvn(res, chol):
            z = torch.linalg.solve_triangular(chol, res[..., None], upper=False)[..., 0]
            half_log_det = torch.log(torch.diagonal(chol, dim1=-2, dim2=-1).abs()).sum(-1)
            return -0.5 * torch.sum(z * z, dim=-1) - half_log_det - 0.5 * res.shape[-1] * math.log(2.0 * math.pi)

        init = mvn(x[..., 0, :] - self._mu_0, torch.tril(self._chol_P_0))
        cond = (self._A_s @ x[..., :-1, :, None])[..., 0] + self._b_s
        rest = mvn(x[..., 1:, :] - cond, torch.tril(self._chol_Q_s))
        return init + rest.sum(-1)

    def kl_divergence(self, dist: GaussMarkovDistribution) -> torch.Tensor:
        """``KL(self ‖ dist)`` with shape ``batch_shape`` (reference :528-593)."""
        check_compatible(self, dist)
        q = self._flat()
        p = dist._flat()
        bsz, t, d = q[5], q[6], q[7]
        if needs_grad(*q[:5], *p[:5]):
            from .autograd import kl_divergence_diff

            return kl_divergence_diff(q[:5], p[:5]).reshape(tuple(self.batch_shape))
        out = torch.empty(bsz, dtype=q[2].dtype, device=q[2].device)
        check(
            _lib.lib().mf_ssm_kl_divergence(
                dtype_code(q[2].dtype), *(ptr(x) for x in q[:5]), *(ptr(x) for x in p[:5]),
                ptr(out), i64(bsz), i64(t), i64(d), current_stream()),
            "mf_ssm_kl_divergence",
        )
        return out.reshape(tuple(self.batch_shape))

    @boundary
    def normalizer(self) -> torch.Tensor:
        """Reference :595-609."""
        dim = (self.num_transitions + 1) * self.state_dim
        mean = self.marginal_means
        prec = self.precision
        mahalanobis = torch.sum(mean * prec.dense_mult(mean), dim=(-2, -1))
        return 0.5 * (dim * math.log(2.0 * math.pi) - self.log_det_precision() + mahalanobis)


@boundary
def _precision_blocks_cuda(l0, a, lq, h=None, r_inv=None):
    """``mf_ssm_build_precision`` on flat operands: ``l0 [B,D,D]``, ``a / lq [B,T-1,D,D]``, optional
    ``h [1|B,T,m,D]``, ``r_inv [1|T,m,m]``  ->  ``(diag [B,T,D,D], sub [B,T-1,D,D])``."""
    bsz, n, d, _ = a.shape
    t = n + 1
    diag = torch.empty(bsz, t, d, d, dtype=a.dtype, device=a.device)
    sub = torch.empty(bsz, n, d, d, dtype=a.dtype, device=a.device)
    m, hb, rs = 0, 1, 1
    if h is not None:
        m, hb, rs = int(h.shape[-2]), int(h.shape[0]), int(r_inv.shape[0])
    check(
        _lib.lib().mf_ssm_build_precision(
            dtype_code(a.dtype), ptr(l0), ptr(a), ptr(lq), ptr(h), ptr(r_inv), ptr(diag),
            ptr(sub), i64(bsz), i64(t), i64(d), i64(m), i64(hb), i64(rs), current_stream()),
        "mf_ssm_build_precision",
    )
    return diag, sub


def cholesky_or_zero(covariance) -> torch.Tensor:
    """Cholesky factor of every ``[D,D]`` block, an all-zero block mapping to zero (reference
    ``state_space_model.py:634-656``)."""
    cov = as_torch(covariance)
    require_cuda(cov, "covariance")
    d = int(cov.shape[-1])
    flat = cov.reshape(-1, d, d).contiguous()
    out = torch.empty_like(flat)
    info = torch.empty(1, dtype=torch.int32, device=flat.device)
    check(
        _lib.lib().mf_block_cholesky_or_zero(
            dtype_code(flat.dtype), ptr(flat), ptr(out), ptr(info), i64(flat.shape[0]), i64(d),
            current_stream()),
        "mf_block_cholesky_or_zero",
    )
    if check_numerics() and int(info[0]) != 0:
        raise CholeskyError(f"cholesky_or_zero: block {int(info[0]) - 1} is not positive definite")
    return out.reshape(cov.shape)


@boundary
def state_space_model_from_covariances(initial_mean, initial_covariance, state_transitions,
                                       state_offsets, process_covariances) -> StateSpaceModel:
    """Reference ``state_space_model.py:613-664``."""
    return StateSpaceModel(
        initial_mean=initial_mean,
        chol_initial_covariance=cholesky_or_zero(initial_covariance),
        state_transitions=state_transitions,
        state_offsets=state_offsets,
        chol_process_covariances=cholesky_or_zero(process_covariances),
    )
