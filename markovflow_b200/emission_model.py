"""``EmissionModel`` (reference ``markovflow/emission_model.py:25-153``): the linear map ``f = H x``
from states to observations.  It is the operand container the Kalman filter API takes; the
projections are per-step elementwise maps (no recursion) and are left to torch."""
from __future__ import annotations

from typing import Tuple

import torch

from .interop import framework_of, as_torch, boundary


class EmissionModel:
    """``emission_matrix``: ``batch_shape + [num_data, output_dim, state_dim]`` (reference :40-51)."""

    def __init__(self, emission_matrix) -> None:
        self._fw = framework_of(emission_matrix)
        h = as_torch(emission_matrix)
        if h.dim() < 3:
            raise ValueError("emission_matrix must be [..., num_data, output_dim, state_dim]")
        self._H = h

    @property
    def _dev(self) -> torch.device:
        return self._H.device

    @property
    def batch_shape(self) -> torch.Size:
        return self._H.shape[:-3]

    @property
    def num_data(self) -> int:
        return int(self._H.shape[-3])

    @property
    def output_dim(self) -> int:
        return int(self._H.shape[-2])

    @property
    def state_dim(self) -> int:
        return int(self._H.shape[-1])

    @property
    @boundary
    def emission_matrix(self) -> torch.Tensor:
        return self._H

    @boundary
    def project_state_to_f(self, state) -> torch.Tensor:
        """``H x`` (reference :115-128)."""
        state = as_torch(state, self._H.device)
        return (self._H @ state[..., None])[..., 0]

    @boundary
    def project_state_covariance_to_f(self, covariance, full_output_cov: bool = False) -> torch.Tensor:
        """``H S Hᵀ`` or its diagonal (reference :130-153)."""
        cov = as_torch(covariance, self._H.device)
        if tuple(cov.shape[-3:]) != (self.num_data, self.state_dim, self.state_dim):
            raise ValueError("covariance must be [..., num_data, state_dim, state_dim]")
        hs = self._H @ cov
        if full_output_cov:
            return hs @ self._H.transpose(-1, -2)
        return torch.sum(hs * self._H, dim=-1)

    @boundary
    def project_state_marginals_to_f(self, means, covariances, full_output_cov: bool = False
                                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        return (self.project_state_to_f(means),
                self.project_state_covariance_to_f(covariances, full_output_cov))
