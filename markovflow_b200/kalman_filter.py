"""Kalman filters with the API of ``markovflow/kalman_filter.py:32-626``.

``log_likelihood`` runs as ONE fused CUDA sweep per chain (``mf_kalman_log_likelihood``): many
chains -> one thread per chain; few long chains -> the parallel-in-time scan (segment summaries,
prefix over summaries, seeded local filters).  The reference evaluates the same quantity in SpInGP
form through five banded operations and two block<->band repacks (``kalman_filter.py:184-255``).
``posterior_state_space_model`` keeps the reference's algebra (``:109-182``) on the O(T) CUDA
``upper_diagonal_lower`` / solve kernels.
"""
from __future__ import annotations

import abc
import math
from typing import Optional

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr
from .block_tri_diag import LowerTriangularBlockTriDiagonal, SymmetricBlockTriDiagonal, _prod
from .emission_model import EmissionModel
from .autograd import needs_grad
from .interop import framework_of, as_torch, boundary, require_cuda
from .state_space_model import StateSpaceModel, cholesky_or_zero

def _workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    """Scratch for the parallel-in-time path (the caller-provided workspace of the ABI), allocated per
    call: torch's caching allocator is stream-ordered and CUDA-graph safe, so concurrent streams never
    share a buffer and a captured graph keeps its own (a process-wide cache did neither)."""
    if nbytes == 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


@boundary
def kalman_log_likelihood(ssm: StateSpaceModel, emission_matrix: torch.Tensor,
                          observations: torch.Tensor, chol_obs_covariance: torch.Tensor
                          ) -> torch.Tensor:
    """Per-chain marginal log-likelihood ``[batch_shape]``.

    ``emission_matrix``: ``[T,m,D]`` or ``batch + [T,m,D]``; ``observations``: ``batch + [T,m]``;
    ``chol_obs_covariance``: ``[m,m]`` or ``[T,m,m]`` (Cholesky of ``R`` / ``R_k``; an infinite
    entry with m = 1 marks a step without observation).
    """
    mu0, l0, a, b, lq, bsz, t, d = ssm._flat()
    h = as_torch(emission_matrix, a.device)
    y = as_torch(observations, a.device)
    lr = as_torch(chol_obs_covariance, a.device)
    m = int(h.shape[-2])
    if tuple(h.shape[-3:]) != (t, m, d):
        raise ValueError(f"emission matrix must be [..., {t}, m, {d}], got {tuple(h.shape)}")
    if tuple(y.shape) != tuple(ssm.batch_shape) + (t, m):
        raise ValueError(
            "The shape of the observations and the state-space-model parameters are not compatible")
    if tuple(lr.shape) not in ((m, m), (t, m, m)):
        raise ValueError(
            "The shape of the observation covariance matrix and the emission matrix are not compatible")
    hb = 1 if h.dim() == 3 else bsz
    if h.dim() > 3 and tuple(h.shape[:-3]) != tuple(ssm.batch_shape):
        h = h.expand(tuple(ssm.batch_shape) + (t, m, d))
    h = h.reshape(hb, t, m, d).contiguous().to(a.dtype)
    y = y.reshape(bsz, t, m).contiguous().to(a.dtype)
    rs = 1 if lr.dim() == 2 else t
    lr = lr.reshape(rs, m, m).contiguous().to(a.dtype)
    if needs_grad(mu0, l0, a, b, lq, h, y, lr):
        from .autograd import kalman_log_likelihood_diff
        from .block_tri_diag import _raise_if_failed

        ll, info = kalman_log_likelihood_diff(mu0, l0, a, b, lq, h, y, lr if lr.shape[0] > 1 else lr[0])
        _raise_if_failed(info, "KalmanFilter.log_likelihood")
        return ll.reshape(tuple(ssm.batch_shape))
    out = torch.empty(bsz, dtype=a.dtype, device=a.device)
    lib = _lib.lib()
    lib.mf_kalman_workspace_bytes.restype = _lib.ctypes.c_size_t
    nbytes = int(lib.mf_kalman_workspace_bytes(dtype_code(a.dtype), i64(bsz), i64(t), i64(d)))
    ws = _workspace(nbytes, a.device)
    check(
        lib.mf_kalman_log_likelihood(
            dtype_code(a.dtype), ptr(mu0), ptr(l0), ptr(a), ptr(b), ptr(lq), ptr(h), ptr(y), ptr(lr),
            ptr(out), i64(bsz), i64(t), i64(d), i64(m), i64(hb), i64(rs), ptr(ws),
            _lib.ctypes.c_size_t(nbytes), current_stream()),
        "mf_kalman_log_likelihood",
    )
    return out.reshape(tuple(ssm.batch_shape))


class BaseKalmanFilter(abc.ABC):
    """Reference ``kalman_filter.py:32-271``."""

    def __init__(self, state_space_model: StateSpaceModel, emission_model: EmissionModel) -> None:
        self.prior_ssm = state_space_model
        self.emission = emission_model
        self._fw = framework_of(state_space_model, emission_model)

    @property
    def _dev(self) -> torch.device:
        return self.prior_ssm._dev

    @property
    @abc.abstractmethod
    def _r_inv(self) -> torch.Tensor:
        """Precision of the observation model, ``[m,m]`` or ``[T,m,m]``."""

    @property
    @abc.abstractmethod
    def _chol_r(self) -> torch.Tensor:
        """Cholesky factor of the observation covariance, ``[m,m]`` or ``[T,m,m]``."""

    @property
    @abc.abstractmethod
    def observations(self) -> torch.Tensor: ...

    @property
    @boundary
    def _k_inv_prior(self) -> SymmetricBlockTriDiagonal:
        return self.prior_ssm.precision

    @property
    @boundary
    def _k_inv_post(self) -> SymmetricBlockTriDiagonal:
        """``K⁻¹ + GᵀΣ⁻¹G`` (reference :85-101), built in one kernel."""
        return SymmetricBlockTriDiagonal(
            *self.prior_ssm._precision_blocks(self.emission.emission_matrix, self._r_inv))

    @property
    def _log_det_observation_precision(self) -> torch.Tensor:
        r_inv = self._r_inv
        t = self.prior_ssm.num_transitions + 1
        if r_inv.dim() == 2:
            return t * torch.logdet(r_inv)
        return torch.sum(torch.logdet(r_inv), dim=-1)

    @boundary
    def _back_project_y_to_state(self, observations) -> torch.Tensor:
        """``HᵀR⁻¹y`` (reference :257-271)."""
        h = self.emission.emission_matrix
        r_inv = self._r_inv
        ry = (r_inv @ as_torch(observations, h.device)[..., None])
        return (h.transpose(-1, -2) @ ry)[..., 0]

    @boundary
    def log_likelihood_per_chain(self) -> torch.Tensor:
        return kalman_log_likelihood(self.prior_ssm, self.emission.emission_matrix,
                                     self.observations, self._chol_r)

    @boundary
    def log_likelihood(self) -> torch.Tensor:
        """Marginal log-likelihood, summed over the batch (reference :184-255)."""
        per_chain = self.log_likelihood_per_chain()
        # one series: the sum is the value itself (no reduction launch behind a 0.2 ms kernel)
        return per_chain.reshape(()) if per_chain.numel() == 1 else torch.sum(per_chain)

    @boundary
    def posterior_state_space_model(self) -> StateSpaceModel:
        """The posterior as a state-space model (reference :109-182).

        The reference's chain -- ``K_post^-1 = U D U^T``, ``A^-T`` solve, two triangular solves, Cholesky of the
        block inverses -- is exactly ``naturals_to_ssm_params`` applied to the posterior's natural parameters
        ``theta_lin = H^T R^-1 y + K^-1 mu_prior``, ``theta_diag = -1/2 diag(K_post^-1)``,
        ``theta_sub = -sub(K_post^-1)`` (``ssm_gaussian_transformations.py:332-511`` states the same recursion),
        which is ONE backward sweep here (``mf_nat_to_ssm``): a handful of launches instead of ~25.  Under
        autograd the operator-by-operator composition below is kept (its pieces have adjoint sweeps)."""
        from .autograd import needs_grad
        from .ssm_gaussian_transformations import naturals_to_ssm_params, ssm_to_naturals

        prior = self.prior_ssm
        if not needs_grad(prior._A_s, prior._b_s, prior._chol_Q_s, prior._chol_P_0, prior._mu_0,
                          self.emission.emission_matrix, as_torch(self.observations), self._r_inv):
            post = self._k_inv_post
            theta_lin = ssm_to_naturals(prior)[0] + self._back_project_y_to_state(self.observations)
            a_s, offsets, chol_p0, chol_qs, mu0 = naturals_to_ssm_params(
                theta_lin, -0.5 * post.block_diagonal, -post.block_sub_diagonal)
            return StateSpaceModel(initial_mean=mu0, chol_initial_covariance=chol_p0, state_transitions=a_s,
                                   state_offsets=offsets, chol_process_covariances=chol_qs)
        a_inv_post, chol_q_inv_post = self._k_inv_post.upper_diagonal_lower()
        obs_proj = self._back_project_y_to_state(self.observations)
        k_inv_mu_prior = self._k_inv_prior.dense_mult(self.prior_ssm.marginal_means)
        rhs = obs_proj + k_inv_mu_prior
        m_post = chol_q_inv_post.solve(
            chol_q_inv_post.solve(a_inv_post.solve(rhs, transpose_left=True)), transpose_left=True)
        concatted_qs = chol_q_inv_post.cholesky_of_block_inverses()
        return StateSpaceModel(
            initial_mean=m_post[..., 0, :],
            chol_initial_covariance=concatted_qs[..., 0, :, :],
            state_transitions=-a_inv_post.block_sub_diagonal,
            state_offsets=m_post[..., 1:, :],
            chol_process_covariances=concatted_qs[..., 1:, :, :],
        )


class KalmanFilter(BaseKalmanFilter):
    """Gaussian observations with a shared covariance (reference :275-353)."""

    def __init__(self, state_space_model: StateSpaceModel, emission_model: EmissionModel,
                 observations, chol_obs_covariance) -> None:
        super().__init__(state_space_model, emission_model)
        obs = as_torch(observations)
        lr = as_torch(chol_obs_covariance, obs.device)
        m = emission_model.output_dim
        if tuple(lr.shape) != (m, m):
            raise ValueError(
                "The shape of the observation covariance matrix and the emission matrix are not compatible")
        want = tuple(state_space_model.batch_shape) + (state_space_model.num_transitions + 1, m)
        if tuple(obs.shape) != want:
            raise ValueError(
                "The shape of the observations and the state-space-model parameters are not compatible")
        self._chol_obs_covariance = lr
        self._observations = obs
        self._fw = framework_of(state_space_model, emission_model, observations, chol_obs_covariance)

    @property
    def _chol_r(self) -> torch.Tensor:
        return self._chol_obs_covariance

    @property
    def _r_inv(self) -> torch.Tensor:
        eye = torch.eye(self.emission.output_dim, dtype=self._chol_obs_covariance.dtype,
                        device=self._chol_obs_covariance.device)
        return torch.cholesky_solve(eye, self._chol_obs_covariance)

    @property
    def observations(self) -> torch.Tensor:
        return self._observations


class GaussianSites(abc.ABC):
    """Reference :356-379."""

    @property
    @abc.abstractmethod
    def means(self) -> torch.Tensor: ...

    @property
    @abc.abstractmethod
    def precisions(self) -> torch.Tensor: ...

    @property
    @abc.abstractmethod
    def log_det_precisions(self) -> torch.Tensor: ...


class UnivariateGaussianSitesNat(GaussianSites):
    """Univariate sites in natural parameters ``nat1 [T,1]``, ``nat2 [T,1,1]`` (reference :382-433)."""

    def __init__(self, nat1, nat2, log_norm=None) -> None:
        self._fw = framework_of(nat1, nat2, log_norm)
        self.num_data, self.output_dim = nat1.shape
        if tuple(nat2.shape) != (self.num_data, 1, 1) or self.output_dim != 1:
            raise ValueError("nat1 must be [N, 1] and nat2 [N, 1, 1]")
        self.nat1 = as_torch(nat1)
        self.nat2 = as_torch(nat2, self.nat1.device)
        self.log_norm = torch.zeros_like(self.nat1) if log_norm is None else as_torch(log_norm)

    @property
    @boundary
    def means(self) -> torch.Tensor:
        return -0.5 * self.nat1 / self.nat2[..., 0]

    @property
    @boundary
    def precisions(self) -> torch.Tensor:
        return -2.0 * self.nat2

    @property
    @boundary
    def log_det_precisions(self) -> torch.Tensor:
        return torch.log(-2.0 * self.nat2)


def _chol_cov_from_precisions(prec: torch.Tensor) -> torch.Tensor:
    """Cholesky of ``R_k = prec_k⁻¹`` for ``[T,m,m]`` precisions."""
    if prec.shape[-1] == 1:
        return torch.rsqrt(prec)
    return torch.linalg.cholesky(torch.linalg.inv(prec))


class KalmanFilterWithSites(BaseKalmanFilter):
    """Per-step Gaussian sites as observations (reference :436-497)."""

    def __init__(self, state_space_model: StateSpaceModel, emission_model: EmissionModel,
                 sites: GaussianSites) -> None:
        self.sites = sites
        super().__init__(state_space_model, emission_model)
        self._fw = framework_of(state_space_model, emission_model, sites)

    @property
    def _r_inv(self) -> torch.Tensor:
        return self.sites.precisions

    @property
    def _chol_r(self) -> torch.Tensor:
        return _chol_cov_from_precisions(self.sites.precisions)

    @property
    def _log_det_observation_precision(self) -> torch.Tensor:
        return torch.sum(torch.logdet(self._r_inv), dim=-1)

    @property
    def observations(self) -> torch.Tensor:
        return self.sites.means


class KalmanFilterWithSparseSites(BaseKalmanFilter):
    """Sites on a subset of a time grid (reference :500-626).  Grid points without data carry no
    observation: in the fused kernel they are steps whose noise scale is infinite."""

    def __init__(self, state_space_model: StateSpaceModel, emission_model: EmissionModel,
                 sites: GaussianSites, num_grid_points: int, observations_index, observations) -> None:
        self.sites = sites
        self.observations_index = as_torch(observations_index).reshape(-1).long()
        self.sparse_observations = self._drop_batch_shape(as_torch(observations))
        self.grid_shape = (int(num_grid_points), 1)
        super().__init__(state_space_model, emission_model)
        self._fw = framework_of(state_space_model, emission_model, sites, observations_index, observations)

    @staticmethod
    def _drop_batch_shape(tensor: torch.Tensor) -> torch.Tensor:
        if tensor.dim() < 3:
            return tensor
        if tensor.shape[0] != 1:
            raise Exception("KalmanFilterWithSparseSites doesn't support batches")
        return tensor.squeeze(0)

    def sparse_to_dense(self, tensor: torch.Tensor, output_shape, fill: float = 0.0) -> torch.Tensor:
        dense = torch.full(tuple(output_shape), fill, dtype=tensor.dtype, device=tensor.device)
        dense[self.observations_index.to(tensor.device)] = tensor
        return dense

    def dense_to_sparse(self, tensor: torch.Tensor) -> torch.Tensor:
        expand = tensor.dim() == 3
        out = tensor.reshape(-1, 1)[self.observations_index.to(tensor.device)]
        return out[..., None] if expand else out

    @property
    def _r_inv_data(self) -> torch.Tensor:
        return self.sites.precisions

    @property
    def _r_inv(self) -> torch.Tensor:
        return self.sparse_to_dense(self.sites.precisions, self.grid_shape + (1,))

    @property
    def _chol_r(self) -> torch.Tensor:
        return self.sparse_to_dense(torch.rsqrt(self.sites.precisions), self.grid_shape + (1,),
                                    fill=math.inf)

    @property
    def _log_det_observation_precision(self) -> torch.Tensor:
        return torch.sum(torch.logdet(self._r_inv_data), dim=-1)

    @property
    def observations(self) -> torch.Tensor:
        return self.sparse_to_dense(self.sparse_observations, self.grid_shape)

    @boundary
    def log_likelihood_per_chain(self) -> torch.Tensor:
        ssm = self.prior_ssm
        obs = self.observations
        if len(ssm.batch_shape):
            obs = obs.expand(tuple(ssm.batch_shape) + tuple(obs.shape))
        return kalman_log_likelihood(ssm, self.emission.emission_matrix, obs, self._chol_r)
