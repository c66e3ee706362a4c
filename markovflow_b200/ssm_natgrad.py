"""``SSMNaturalGradient``: natural-gradient descent on a :class:`StateSpaceModel` variational posterior, with
the interface and update rule of ``markovflow/ssm_natgrad.py:31-218`` (Salimbeni, Eleftheriadis & Hensman,
AISTATS 2018, eq. 10).

One step (``_natgrad_step``, reference :121-218):

1. ``dL/d(ssm)`` -- the ordinary gradient of the loss with respect to ``(A_s, b_s, chol_P0, chol_Qs, mu0)``;
2. ``dL/d(eta) = (d ssm / d eta)^T dL/d(ssm)`` -- chain rule through ``expectations_to_ssm_params`` evaluated at
   ``eta = ssm_to_expectations(ssm)``;
3. ``theta_new = theta - gamma * dL/d(eta)`` with ``theta = ssm_to_naturals(ssm)``;
4. ``ssm <- naturals_to_ssm_params(theta_new)``, written in place into the model's parameters.

Every piece runs on the CUDA operators: the forward sweeps, the adjoint sweeps behind ``autograd.py`` for (1)-(2),
and the backward ``U D U^T`` sweep of ``mf_nat_to_ssm`` for (4).  The momentum variant of the reference (:173-203)
needs ``dL/d(theta)`` through ``naturals_to_ssm_params``, whose adjoint sweep does not exist yet.
"""
from __future__ import annotations

from typing import Callable

import torch

from .ssm_gaussian_transformations import (
    expectations_to_ssm_params,
    naturals_to_ssm_params,
    ssm_to_expectations,
    ssm_to_naturals,
)
from .state_space_model import StateSpaceModel


class SSMNaturalGradient:
    """Reference ``ssm_natgrad.py:31-218``: ``minimize(loss_fn, ssm)`` performs one natural-gradient step on
    ``ssm`` (a :class:`StateSpaceModel` whose parameters require gradients, e.g. ``create_trainable_copy()``)."""

    def __init__(self, gamma: float = 0.1, momentum: bool = False, beta1: float = 0.9, beta2: float = 0.99,
                 epsilon: float = 1e-8, name: str = "SSMNaturalGradient") -> None:
        if momentum:
            raise NotImplementedError(
                "momentum needs dL/d(theta) through naturals_to_ssm_params (ssm_natgrad.py:173-176), whose "
                "adjoint sweep is not built; momentum=False is the reference's own integration-test setting")
        self.gamma, self._name = float(gamma), name

    def minimize(self, loss_fn: Callable[[], torch.Tensor], ssm: StateSpaceModel) -> None:
        self._natgrad_step(loss_fn, ssm)

    def _natgrad_step(self, loss_fn: Callable[[], torch.Tensor], ssm: StateSpaceModel) -> None:
        params = ssm.trainable_variables  # (A_s, b_s, chol_P0, chol_Qs, mu0): the order of the transforms' outputs
        if not all(p.requires_grad for p in params):
            raise ValueError("the state-space model's parameters must require gradients "
                             "(StateSpaceModel.create_trainable_copy())")
        with torch.enable_grad():
            loss = loss_fn()
            dl_dssm = torch.autograd.grad(loss, params, allow_unused=True)
            dl_dssm = [torch.zeros_like(p) if g is None else g for p, g in zip(params, dl_dssm)]
            # the Cholesky factors are lower triangular: only those entries are parameters (reference: the
            # FillTriangular transform of :160-161)
            dl_dssm[2], dl_dssm[3] = torch.tril(dl_dssm[2]), torch.tril(dl_dssm[3])
            etas = [e.detach().requires_grad_(True) for e in ssm_to_expectations(ssm)]
            ssm_params = expectations_to_ssm_params(*etas)
            dl_detas = torch.autograd.grad(ssm_params, etas, grad_outputs=dl_dssm, allow_unused=True)
        with torch.no_grad():
            thetas = ssm_to_naturals(_detached(ssm))
            thetas_new = [th - self.gamma * g for th, g in zip(thetas, dl_detas)]
            new = naturals_to_ssm_params(*thetas_new)
            for p, v in zip(params, new):
                p.copy_(v)
        ssm._flat_cache = None  # the parameters changed in place


def _detached(ssm: StateSpaceModel) -> StateSpaceModel:
    return StateSpaceModel(ssm._mu_0.detach(), ssm._chol_P_0.detach(), ssm._A_s.detach(), ssm._b_s.detach(),
                           ssm._chol_Q_s.detach())
