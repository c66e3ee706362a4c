"""Abstract Gauss-Markov distribution (reference ``markovflow/gauss_markov.py:29-216``)."""
from __future__ import annotations

import abc
from typing import Tuple

import torch

from .interop import boundary

from .block_tri_diag import SymmetricBlockTriDiagonal


class GaussMarkovDistribution(abc.ABC):
    """Interface of distributions over ``[num_transitions + 1, state_dim]`` trajectories whose
    precision is block-tridiagonal (reference ``gauss_markov.py:29-201``)."""

    @property
    @abc.abstractmethod
    def event_shape(self) -> torch.Size: ...

    @property
    @abc.abstractmethod
    def batch_shape(self) -> torch.Size: ...

    @property
    @abc.abstractmethod
    def state_dim(self) -> int: ...

    @property
    @abc.abstractmethod
    def num_transitions(self) -> int: ...

    @abc.abstractmethod
    def _build_precision(self) -> SymmetricBlockTriDiagonal: ...

    @property
    @boundary
    def precision(self) -> SymmetricBlockTriDiagonal:
        """Block-tridiagonal precision ``K⁻¹`` (reference ``gauss_markov.py:72-78``)."""
        return self._build_precision()

    @property
    @abc.abstractmethod
    def marginal_means(self) -> torch.Tensor: ...

    @property
    @abc.abstractmethod
    def marginal_covariances(self) -> torch.Tensor: ...

    @abc.abstractmethod
    def covariance_blocks(self) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @property
    @boundary
    def marginals(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(μ_k, Σ_kk)`` (reference ``gauss_markov.py:107-117``)."""
        return self.marginal_means, self.marginal_covariances

    @abc.abstractmethod
    def sample(self, sample_shape) -> torch.Tensor: ...

    @abc.abstractmethod
    def log_det_precision(self) -> torch.Tensor: ...

    @abc.abstractmethod
    def log_pdf(self, states) -> torch.Tensor: ...

    @abc.abstractmethod
    def create_trainable_copy(self) -> "GaussMarkovDistribution": ...

    @abc.abstractmethod
    def create_non_trainable_copy(self) -> "GaussMarkovDistribution": ...

    @abc.abstractmethod
    def kl_divergence(self, dist: "GaussMarkovDistribution") -> torch.Tensor: ...


def check_compatible(dist_1: GaussMarkovDistribution, dist_2: GaussMarkovDistribution) -> None:
    """Same representation, state dim, batch shape and length (reference ``gauss_markov.py:204-216``)."""
    if not isinstance(dist_2, type(dist_1)):
        raise TypeError("`dist_2` has different representation than `dist_1`")
    if dist_1.state_dim != dist_2.state_dim:
        raise ValueError("state_dim differs")
    if tuple(dist_1.batch_shape) != tuple(dist_2.batch_shape):
        raise ValueError("batch_shape differs")
    if dist_1.num_transitions != dist_2.num_transitions:
        raise ValueError("num_transitions differs")
    d1, d2 = getattr(dist_1, "_A_s", None), getattr(dist_2, "_A_s", None)
    if d1 is not None and d2 is not None and (d1.dtype != d2.dtype or d1.device != d2.device):
        raise ValueError(f"dtype / device differ: {d1.dtype} on {d1.device} vs {d2.dtype} on {d2.device}")
