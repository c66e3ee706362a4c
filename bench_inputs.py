"""Synthetic inputs of the BASELINE.json configurations, generated ON DEVICE with torch (setup code,
never timed).  Formulas restate the reference's closed forms (``markovflow/kernels/matern.py:434-501``,
``kernels/sde_kernel.py:421-446``, ``state_space_model.py:431-483``, ``kalman_filter.py:85-101``); the
numpy oracle (``oracle/np_oracle.py``) builds the same inputs on the host for parity checks."""
from __future__ import annotations

import math

import torch

SEED = 71892305  # the reference's test seed (tests/conftest.py:22)


def matern52_posterior_precision(b: int, t: int, device, seed: int = SEED, r_inv: float = 100.0,
                                 chunk: int = 512):
    """Config 2: per chain Matern52 with lengthscale ~U(0.5,2), variance ~U(0.5,2),
    dt_k = lengthscale*U(0.2,1.0); returns the posterior precision K^-1 + H^T R^-1 H as blocks
    (diag [b,t,3,3], sub [b,t-1,3,3]) and a N(0,1) right-hand side [b,t,3], float64."""
    g = torch.Generator(device=device).manual_seed(seed)
    f64 = torch.float64
    diag = torch.empty(b, t, 3, 3, dtype=f64, device=device)
    sub = torch.empty(b, t - 1, 3, 3, dtype=f64, device=device)
    eye = torch.eye(3, dtype=f64, device=device)
    for b0 in range(0, b, chunk):
        nb = min(chunk, b - b0)
        ell = 0.5 + 1.5 * torch.rand(nb, generator=g, dtype=f64, device=device)
        var = 0.5 + 1.5 * torch.rand(nb, generator=g, dtype=f64, device=device)
        dt = ell[:, None] * (0.2 + 0.8 * torch.rand(nb, t - 1, generator=g, dtype=f64, device=device))
        lam = math.sqrt(5.0) / ell  # [nb]
        l2, l3, l4 = lam ** 2, lam ** 3, lam ** 4
        feedback = torch.zeros(nb, 3, 3, dtype=f64, device=device)
        feedback[:, 0, 1] = 1.0
        feedback[:, 1, 2] = 1.0
        feedback[:, 2, 0] = -l3
        feedback[:, 2, 1] = -3.0 * l2
        feedback[:, 2, 2] = -3.0 * lam
        pinf = torch.zeros(nb, 3, 3, dtype=f64, device=device)
        pinf[:, 0, 0] = 1.0
        pinf[:, 0, 2] = -l2 / 3.0
        pinf[:, 1, 1] = l2 / 3.0
        pinf[:, 2, 0] = -l2 / 3.0
        pinf[:, 2, 2] = l4
        pinf = var[:, None, None] * pinf
        flt = (feedback + lam[:, None, None] * eye)[:, None] * dt[..., None, None]  # [nb,t-1,3,3]
        a = torch.exp(-lam[:, None] * dt)[..., None, None] * (eye + flt + flt @ flt / 2.0)
        q = pinf[:, None] - a @ pinf[:, None] @ a.transpose(-1, -2)
        chol_q = torch.linalg.cholesky(q)
        chol_p0 = torch.linalg.cholesky(pinf)
        inv_q_a = torch.cholesky_solve(a, chol_q)
        aqa = a.transpose(-1, -2) @ inv_q_a
        chols = torch.cat([chol_p0[:, None], chol_q], dim=1)
        inv_q = torch.cholesky_solve(eye.expand(nb, t, 3, 3), chols)
        d = inv_q
        d[:, :-1] += aqa
        d[:, :, 0, 0] += r_inv  # H = [1, 0, 0], R^-1 = r_inv
        diag[b0:b0 + nb] = d
        sub[b0:b0 + nb] = -inv_q_a
        del flt, a, q, chol_q, inv_q_a, aqa, chols, inv_q, d
    rhs = torch.randn(b, t, 3, generator=g, dtype=f64, device=device)
    return diag, sub, rhs
