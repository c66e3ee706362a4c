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
and the backward ``U D U^T`` sweep of ``mf_nat_to_ssm`` for (4).  The momentum variant (:173-203, Adam-style
moving averages of the natural gradient and of its norm ``<dL/d(eta), dL/d(theta)>``) takes ``dL/d(theta)``
through the differentiable ``naturals_to_ssm_params`` (``autograd.naturals_to_ssm_diff``).
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
        self.gamma, self._name = float(gamma), name
        self._momentum = bool(momentum)
        self._beta1, self._beta2, self._epsilon = float(beta1), float(beta2), float(epsilon)
        self._ms = None      # moving averages of dL/d(eta) (reference :68-71, 104-108)
        self._v = 0.0        # moving average of the natural-gradient norm
        self._step_counter = 1
        self._effective_lr = None

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
            dl_detas = [torch.zeros_like(e) if g is None else g for e, g in zip(etas, dl_detas)]
            dl_dthetas = None
            if self._momentum:  # dL/d(theta) by the chain rule through naturals_to_ssm_params (:173-176, 190)
                thetas_g = [th.detach().requires_grad_(True) for th in ssm_to_naturals(_detached(ssm))]
                ssm_params_2 = naturals_to_ssm_params(*thetas_g)
                dl_dthetas = torch.autograd.grad(ssm_params_2, thetas_g, grad_outputs=dl_dssm, allow_unused=True)
                dl_dthetas = [torch.zeros_like(t) if g is None else g for t, g in zip(thetas_g, dl_dthetas)]
        with torch.no_grad():
            thetas = ssm_to_naturals(_detached(ssm))
            if self._momentum:
                if self._ms is None:
                    self._ms = [torch.zeros_like(g) for g in dl_detas]
                lr = (self.gamma * (1.0 - self._beta2 ** self._step_counter) ** 0.5
                      / (1.0 - self._beta1 ** self._step_counter))
                ms_new = [m * self._beta1 + (1.0 - self._beta1) * g for m, g in zip(self._ms, dl_detas)]
                comps = [torch.sum(g * gt) for g, gt in zip(dl_detas, dl_dthetas)]
                comps[-1] = comps[-1] * 2.0  # the sub-diagonal blocks appear twice in the symmetric precision
                v_new = self._v * self._beta2 + (1.0 - self._beta2) * float(sum(comps))
                denom = v_new ** 0.5 + self._epsilon
                thetas_new = [th - lr * m / denom for th, m in zip(thetas, ms_new)]
                self._ms, self._v = ms_new, v_new
                self._step_counter += 1
                self._effective_lr = lr / denom
            else:
                thetas_new = [th - self.gamma * g for th, g in zip(thetas, dl_detas)]
            new = naturals_to_ssm_params(*thetas_new)
            for p, v in zip(params, new):
                p.copy_(v)
        ssm._flat_cache = None  # the parameters changed in place


def _detached(ssm: StateSpaceModel) -> StateSpaceModel:
    return StateSpaceModel(ssm._mu_0.detach(), ssm._chol_P_0.detach(), ssm._A_s.detach(), ssm._b_s.detach(),
                           ssm._chol_Q_s.detach())
