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


def matern32_ssm(b: int, t: int, device, seed: int = SEED, dt_lo: float = 0.05, dt_hi: float = 0.15,
                 jitter_hyper: bool = False, dtype=torch.float64, chunk_t: int = 1 << 21):
    """Configs 1/3/5: Matern32 (D=2) state-space parameters for ``b`` chains of ``t`` states with
    dt_k ~ U(dt_lo, dt_hi) (reference closed form ``kernels/matern.py:299-356``).  Returns
    (mu0 [b,2], chol_p0 [b,2,2], a [b,t-1,2,2], offsets [b,t-1,2], chol_q [b,t-1,2,2], h [1,t,1,2])."""
    g = torch.Generator(device=device).manual_seed(seed)
    f64 = torch.float64
    if jitter_hyper:
        ell = 0.8 + 0.4 * torch.rand(b, generator=g, dtype=f64, device=device)
        var = 0.8 + 0.4 * torch.rand(b, generator=g, dtype=f64, device=device)
    else:
        ell = torch.ones(b, dtype=f64, device=device)
        var = torch.ones(b, dtype=f64, device=device)
    lam = math.sqrt(3.0) / ell
    a = torch.empty(b, t - 1, 2, 2, dtype=dtype, device=device)
    chol_q = torch.zeros(b, t - 1, 2, 2, dtype=dtype, device=device)
    for k0 in range(0, t - 1, chunk_t):
        n = min(chunk_t, t - 1 - k0)
        dt = dt_lo + (dt_hi - dt_lo) * torch.rand(b, n, generator=g, dtype=f64, device=device)
        la = lam[:, None]
        e = torch.exp(-la * dt)
        a00, a01 = e * (1.0 + la * dt), e * dt
        a10, a11 = -e * la * la * dt, e * (1.0 - la * dt)
        p0, p1 = var[:, None], var[:, None] * la * la  # Pinf = diag(p0, p1)
        q00 = p0 - (a00 * a00 * p0 + a01 * a01 * p1)
        q10 = -(a10 * a00 * p0 + a11 * a01 * p1)
        q11 = p1 - (a10 * a10 * p0 + a11 * a11 * p1)
        l00 = torch.sqrt(q00)
        l10 = q10 / l00
        l11 = torch.sqrt(q11 - l10 * l10)
        sl = slice(k0, k0 + n)
        a[:, sl, 0, 0], a[:, sl, 0, 1], a[:, sl, 1, 0], a[:, sl, 1, 1] = (
            a00.to(dtype), a01.to(dtype), a10.to(dtype), a11.to(dtype))
        chol_q[:, sl, 0, 0], chol_q[:, sl, 1, 0], chol_q[:, sl, 1, 1] = (
            l00.to(dtype), l10.to(dtype), l11.to(dtype))
        del dt, e, a00, a01, a10, a11, q00, q10, q11, l00, l10, l11
    mu0 = torch.zeros(b, 2, dtype=dtype, device=device)
    chol_p0 = torch.zeros(b, 2, 2, dtype=dtype, device=device)
    chol_p0[:, 0, 0] = torch.sqrt(var).to(dtype)
    chol_p0[:, 1, 1] = (torch.sqrt(var) * lam).to(dtype)
    offsets = torch.zeros(b, t - 1, 2, dtype=dtype, device=device)
    h = torch.zeros(1, t, 1, 2, dtype=dtype, device=device)
    h[..., 0] = 1.0
    return mu0, chol_p0, a, offsets, chol_q, h


def matern32_time_deltas(b: int, t: int, device, seed: int = SEED, dt_lo: float = 0.05,
                         dt_hi: float = 0.15, chunk_t: int = 1 << 21):
    """The time deltas ``[b, t-1]`` that :func:`matern32_ssm` (``jitter_hyper=False``: lengthscale =
    variance = 1) drew for the same ``seed`` -- the input of the in-kernel SSM construction path."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(b, t - 1, dtype=torch.float64, device=device)
    for k0 in range(0, t - 1, chunk_t):
        n = min(chunk_t, t - 1 - k0)
        out[:, k0:k0 + n] = dt_lo + (dt_hi - dt_lo) * torch.rand(b, n, generator=g, dtype=torch.float64,
                                                                 device=device)
    return out


def kalman_inputs_config3(t: int, device, seed: int = SEED, noise: float = 0.1):
    """Config 3: ONE Matern32 series of ``t`` states, observations sampled from the model + noise."""
    from markovflow_b200 import StateSpaceModel

    mu0, l0, a, off, lq, h = matern32_ssm(1, t, device, seed)
    ssm = StateSpaceModel(mu0, l0, a, off, lq)
    g = torch.Generator(device=device).manual_seed(seed + 1)
    x = ssm.sample((), generator=g)  # [1,t,2]
    y = x[..., :1] + noise * torch.randn(1, t, 1, generator=g, dtype=torch.float64, device=device)
    chol_r = torch.tensor([[noise]], dtype=torch.float64, device=device)
    return ssm, h, y, chol_r


def cvi_naturals_config5(b: int, t: int, device, seed: int = SEED, dtype=torch.float64):
    """Config 5: natural parameters of the CVI posterior: Matern32 prior precision at ``t`` inducing
    states (linspace grid, per-chain jittered hyper-parameters) + back-projected site naturals
    (``models/variational_cvi.py:106-135,423-445``).  Returns (theta_lin, theta_diag, theta_sub)."""
    from markovflow_b200 import StateSpaceModel

    mu0, l0, a, off, lq, h = matern32_ssm(b, t, device, seed, dt_lo=0.1, dt_hi=0.1, jitter_hyper=True)
    ssm = StateSpaceModel(mu0, l0, a, off, lq)
    g = torch.Generator(device=device).manual_seed(seed + 2)
    prec = 0.5 + 1.5 * torch.rand(t, 1, 1, generator=g, dtype=torch.float64, device=device)
    nat1 = torch.randn(b, t, 1, generator=g, dtype=torch.float64, device=device)
    pd, ps = ssm._precision_blocks(h[0], prec)  # K^-1 + H^T diag(prec) H
    theta_lin = torch.zeros(b, t, 2, dtype=torch.float64, device=device)
    theta_lin[..., 0] = nat1[..., 0]
    return theta_lin.to(dtype), (-0.5 * pd).to(dtype), (-ps).to(dtype)


def sum_kernel_posterior_precision(b: int, t: int, device, seed: int = SEED, r_inv: float = 100.0,
                                   jitter: float = 1e-6, chunk: int = 8):
    """Config 4: Sum([Matern52(1,1)] + [HarmonicOscillator(0.5**j, 1/j) for j=1..7], jitter) => D=17,
    dt ~ U(0.05, 0.15); returns the posterior precision K^-1 + H^T R^-1 H as blocks
    (diag [b,t,17,17], sub [b,t-1,17,17]) and a N(0,1) right-hand side [b,t,17], float64
    (``kernels/sde_kernel.py:540-687``, ``kernels/periodic.py:27-187``)."""
    g = torch.Generator(device=device).manual_seed(seed)
    f64 = torch.float64
    d = 17
    diag = torch.empty(b, t, d, d, dtype=f64, device=device)
    sub = torch.empty(b, t - 1, d, d, dtype=f64, device=device)
    eye3 = torch.eye(3, dtype=f64, device=device)
    eye = torch.eye(d, dtype=f64, device=device)
    lam = math.sqrt(5.0)
    feedback = torch.tensor([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3 * lam ** 2, -3 * lam]],
                            dtype=f64, device=device)
    pinf = torch.zeros(d, d, dtype=f64, device=device)
    pinf[:3, :3] = torch.tensor([[1.0, 0.0, -lam ** 2 / 3], [0.0, lam ** 2 / 3, 0.0],
                                 [-lam ** 2 / 3, 0.0, lam ** 4]], dtype=f64, device=device)
    h = torch.zeros(d, dtype=f64, device=device)
    h[0] = 1.0
    for j in range(1, 8):
        o = 3 + 2 * (j - 1)
        pinf[o, o] = pinf[o + 1, o + 1] = 0.5 ** j
        h[o] = 1.0
    hrh = r_inv * torch.outer(h, h)
    chol_p0 = torch.linalg.cholesky(pinf)
    for b0 in range(0, b, chunk):
        nb = min(chunk, b - b0)
        dt = 0.05 + 0.1 * torch.rand(nb, t - 1, generator=g, dtype=f64, device=device)
        a = torch.zeros(nb, t - 1, d, d, dtype=f64, device=device)
        flt = (feedback + lam * eye3) * dt[..., None, None]
        a[..., :3, :3] = torch.exp(-lam * dt)[..., None, None] * (eye3 + flt + flt @ flt / 2.0)
        for j in range(1, 8):
            o = 3 + 2 * (j - 1)
            ang = dt * (2.0 * math.pi * j)
            c, s = torch.cos(ang), torch.sin(ang)
            a[..., o, o], a[..., o, o + 1], a[..., o + 1, o], a[..., o + 1, o + 1] = c, -s, s, c
        q = pinf - a @ pinf @ a.transpose(-1, -2) + jitter * eye
        chol_q = torch.linalg.cholesky(q)
        inv_q_a = torch.cholesky_solve(a, chol_q)
        aqa = a.transpose(-1, -2) @ inv_q_a
        chols = torch.cat([chol_p0.expand(nb, 1, d, d), chol_q], dim=1)
        dd = torch.cholesky_solve(eye.expand(nb, t, d, d), chols)
        dd[:, :-1] += aqa
        dd += hrh
        diag[b0:b0 + nb] = dd
        sub[b0:b0 + nb] = -inv_q_a
        del dt, a, flt, q, chol_q, inv_q_a, aqa, chols, dd
    rhs = torch.randn(b, t, d, generator=g, dtype=f64, device=device)
    return diag, sub, rhs
