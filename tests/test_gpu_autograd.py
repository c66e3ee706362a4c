"""Reverse mode and the natural-gradient step (§8 a20).

The adjoint CUDA sweeps (``mf_btd_cholesky_bwd``, ``mf_ssm_marginals_bwd``, the transposed solves) are checked
against torch's own reverse mode on DENSE restatements of the same quantities (float64), and the composites
against the identities the reference's tests use:

* gradient of the Kalman log-likelihood with respect to kernel hyper-parameters == gradient of the dense GP
  marginal likelihood (``tests/integration/models/test_gaussian_process_regression.py:117-130``);
* one natural-gradient step with ``gamma = 1`` on a Gaussian likelihood lands on the exact posterior: the ELBO
  equals the GPR log marginal likelihood (``tests/integration/test_ssm_natgrad.py:47-70``).
"""
import math

import numpy as np
import pytest
import torch

from tests.helpers import random_ssm_arrays, random_well_conditioned_spd_btd

pytestmark = pytest.mark.gpu
GTOL = 1e-9  # gradients, float64, max-abs relative to max-abs of the dense-autograd gradient


def dev():
    return torch.device("cuda:0")


def tt(x, grad=False):
    t = torch.as_tensor(np.array(x, dtype=np.float64), device=dev())
    return t.requires_grad_(True) if grad else t


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def dense_from_blocks(diag, sub, symmetric):
    """[T,D,D] (+[T-1,D,D]) -> dense [TD,TD], lower triangles of the diagonal blocks only (differentiable)."""
    t, d, _ = diag.shape
    rows = []
    for i in range(t):
        row = []
        for j in range(t):
            if i == j:
                blk = torch.tril(diag[i])
                if symmetric:
                    blk = blk + torch.tril(diag[i], -1).T
            elif i == j + 1 and sub is not None:
                blk = sub[j]
            elif j == i + 1 and sub is not None and symmetric:
                blk = sub[i].T
            else:
                blk = torch.zeros(d, d, dtype=diag.dtype, device=diag.device)
            row.append(blk)
        rows.append(torch.cat(row, dim=1))
    return torch.cat(rows, dim=0)


@pytest.mark.parametrize("d,t,with_sub", [(1, 5, True), (2, 7, True), (3, 6, True), (5, 4, True), (3, 1, False),
                                          (2, 5, False), (8, 3, True)])
def test_cholesky_solve_logdet_gradients_match_dense_autograd(d, t, with_sub):
    """cholesky + solve + log-det through the adjoint sweeps vs torch.linalg on the dense matrix."""
    from markovflow_b200 import SymmetricBlockTriDiagonal

    b = 3
    if t == 1:
        diag_np = random_well_conditioned_spd_btd((b,), 2, d, rng=d)[0][:, :1]
        sub_np = None
    else:
        diag_np, sub_np, _, _ = random_well_conditioned_spd_btd((b,), t, d, rng=10 * d + t)
        if not with_sub:
            sub_np = None
    rng = np.random.default_rng(d + t)
    rhs_np = rng.standard_normal((b, t, d))
    w_ld, w_ls, w_x = (tt(rng.standard_normal(s)) for s in ((b, t, d, d), (b, max(t - 1, 1), d, d), (b, t, d)))

    def loss_ours(diag, sub, rhs):
        chol, x, logdet = SymmetricBlockTriDiagonal(diag, sub).cholesky_and_solve(rhs, want_log_det=True)
        out = (w_ld * chol.block_diagonal).sum() + (w_x * x).sum() + 1.7 * logdet.sum()
        if sub is not None:
            out = out + (w_ls[:, :t - 1] * chol.block_sub_diagonal).sum()
        # a second, transposed solve through the same factor
        return out + (w_x * chol.solve(x, transpose_left=True)).sum()

    def loss_dense(diag, sub, rhs):
        out = 0.0
        for i in range(b):
            m = dense_from_blocks(diag[i], None if sub is None else sub[i], True)
            low = torch.linalg.cholesky(m)
            x = torch.linalg.solve_triangular(low, rhs[i].reshape(-1, 1), upper=False)
            x2 = torch.linalg.solve_triangular(low.T, x, upper=True)
            ld = torch.stack([low[k * d:(k + 1) * d, k * d:(k + 1) * d] for k in range(t)])
            out = out + (w_ld[i] * ld).sum() + (w_x[i].reshape(-1, 1) * (x + x2)).sum()
            out = out + 1.7 * torch.log(torch.diagonal(low)).sum()
            if sub is not None:
                ls = torch.stack([low[(k + 1) * d:(k + 2) * d, k * d:(k + 1) * d] for k in range(t - 1)])
                out = out + (w_ls[i, :t - 1] * ls).sum()
        return out

    args1 = (tt(diag_np, True), None if sub_np is None else tt(sub_np, True), tt(rhs_np, True))
    args2 = (tt(diag_np, True), None if sub_np is None else tt(sub_np, True), tt(rhs_np, True))
    l1, l2 = loss_ours(*args1), loss_dense(*args2)
    assert abs(float(l1) - float(l2)) < 1e-10 * abs(float(l2))
    l1.backward()
    l2.backward()
    for a1, a2, name in zip(args1, args2, ("diag", "sub", "rhs")):
        if a1 is None:
            continue
        g2 = a2.grad
        if name == "diag":  # both paths read the lower triangles only
            assert float(torch.triu(a1.grad, 1).abs().max()) == 0.0
        assert rel(a1.grad, g2) < GTOL, name


def _propagate_torch(mu0, l0, a, b, lq):
    means, covs, subs = [mu0], [l0 @ l0.transpose(-1, -2)], []
    for k in range(a.shape[-3]):
        ak = a[..., k, :, :]
        subs.append(ak @ covs[-1])
        means.append((ak @ means[-1][..., None])[..., 0] + b[..., k, :])
        covs.append(ak @ covs[-1] @ ak.transpose(-1, -2) + lq[..., k, :, :] @ lq[..., k, :, :].transpose(-1, -2))
    return torch.stack(means, -2), torch.stack(covs, -3), torch.stack(subs, -3)


@pytest.mark.parametrize("d,n", [(1, 5), (2, 9), (3, 6), (6, 3)])
def test_marginals_gradients_match_torch_propagation(batch_shape, d, n):
    import markovflow_b200 as mf

    np.random.seed(d * 10 + n)
    arrays = random_ssm_arrays(batch_shape, n, d)
    rng = np.random.default_rng(n)
    wm, wc, ws = (tt(rng.standard_normal(batch_shape + s)) for s in ((n + 1, d), (n + 1, d, d), (n, d, d)))
    p1 = [tt(x, True) for x in arrays]
    p2 = [tt(x, True) for x in arrays]
    ssm = mf.StateSpaceModel(*p1)
    mean, cov = ssm.marginals
    cov2, sub = ssm.covariance_blocks()
    l1 = (wm * mean).sum() + (wc * cov).sum() + (ws * sub).sum() + (wm * ssm.marginal_means).sum() * 0.5
    tm, tc, tsub = _propagate_torch(p2[0], torch.tril(p2[1]), p2[2], p2[3], torch.tril(p2[4]))
    l2 = (wm * tm).sum() * 1.5 + (wc * tc).sum() + (ws * tsub).sum()
    assert abs(float(l1) - float(l2)) < 1e-10 * abs(float(l2))
    l1.backward()
    l2.backward()
    for q1, q2 in zip(p1, p2):
        assert rel(q1.grad, q2.grad) < GTOL


def _dense_joint(mu0, l0, a, b, lq):
    """Dense joint mean / covariance of ONE chain (differentiable)."""
    n, d = a.shape[0], a.shape[-1]
    mean, cov, _ = _propagate_torch(mu0, l0, a, b, lq)
    rows = []
    for i in range(n + 1):
        row = []
        for j in range(n + 1):
            lo, hi = min(i, j), max(i, j)
            c = cov[lo]
            for k in range(lo, hi):
                c = a[k] @ c
            row.append(c if i >= j else c.T)
        rows.append(torch.cat(row, dim=1))
    return mean.reshape(-1), torch.cat(rows, dim=0)


def _mvn_kl(m1, c1, m2, c2):
    k = m1.shape[0]
    l2 = torch.linalg.cholesky(c2)
    sol = torch.cholesky_solve(c1, l2)
    dm = (m2 - m1)[:, None]
    return 0.5 * (torch.trace(sol) + (dm.T @ torch.cholesky_solve(dm, l2))[0, 0] - k
                  + torch.logdet(c2) - torch.logdet(c1))


@pytest.mark.parametrize("d,n", [(1, 4), (2, 6), (3, 4)])
def test_kl_divergence_gradients_match_dense_gaussians(d, n):
    import markovflow_b200 as mf

    np.random.seed(d + 7 * n)
    qa, pa = random_ssm_arrays((2,), n, d), random_ssm_arrays((2,), n, d)
    q1, p1 = [tt(x, True) for x in qa], [tt(x, True) for x in pa]
    q2, p2 = [tt(x, True) for x in qa], [tt(x, True) for x in pa]
    kl = mf.StateSpaceModel(*q1).kl_divergence(mf.StateSpaceModel(*p1))
    want = torch.stack([_mvn_kl(*_dense_joint(*(x[i] if j not in (1, 4) else torch.tril(x[i]) for j, x in enumerate(q2))),
                                *_dense_joint(*(x[i] if j not in (1, 4) else torch.tril(x[i]) for j, x in enumerate(p2))))
                        for i in range(2)])
    assert rel(kl, want) < 1e-10
    w = tt([1.0, -0.6])
    (w * kl).sum().backward()
    (w * want).sum().backward()
    for a1, a2 in zip(q1 + p1, q2 + p2):
        assert rel(a1.grad, a2.grad) < 1e-8


def _matern32_ssm_torch(ell, var, tp):
    """Matern32 state-space model from hyper-parameters in differentiable torch ops
    (kernels/matern.py:299-356, kernels/sde_kernel.py:421-446)."""
    lam = math.sqrt(3.0) / ell
    dt = (tp[1:] - tp[:-1])[:, None, None]
    eye = torch.eye(2, dtype=tp.dtype, device=tp.device)
    n_mat = torch.stack([torch.stack([lam, torch.ones_like(lam)]), torch.stack([-lam * lam, -lam])])
    a = torch.exp(-lam * dt) * (eye + n_mat * dt)
    pinf = torch.diag(torch.stack([var, var * lam * lam]))
    q = pinf - a @ pinf @ a.transpose(-1, -2)
    return (torch.zeros(2, dtype=tp.dtype, device=tp.device), torch.linalg.cholesky(pinf), a,
            torch.zeros(tp.shape[0] - 1, 2, dtype=tp.dtype, device=tp.device), torch.linalg.cholesky(q))


def _dense_gp_loglik(ell, var, noise, tp, y):
    r = (tp[:, None] - tp[None, :]).abs()
    lam = math.sqrt(3.0) / ell
    k = var * (1.0 + lam * r) * torch.exp(-lam * r) + noise ** 2 * torch.eye(tp.shape[0], dtype=tp.dtype,
                                                                              device=tp.device)
    low = torch.linalg.cholesky(k)
    alpha = torch.cholesky_solve(y[:, None], low)
    return -0.5 * (y[None] @ alpha)[0, 0] - torch.log(torch.diagonal(low)).sum() - 0.5 * tp.shape[0] * math.log(2 * math.pi)


def test_log_likelihood_hyperparameter_gradients_match_the_dense_gp():
    """tests/integration/models/test_gaussian_process_regression.py:117-130: d log p(y) / d(lengthscale,
    variance, noise) from the state-space form equals the dense GP's."""
    import markovflow_b200 as mf

    rng = np.random.default_rng(3)
    tp = tt(np.cumsum(rng.uniform(0.05, 0.4, size=40)))
    y = tt(np.sin(tp.cpu().numpy()) + 0.1 * rng.standard_normal(40))
    h = torch.zeros(40, 1, 2, dtype=torch.float64, device=dev())
    h[..., 0] = 1.0
    hp1 = [tt(v, True) for v in (0.7, 1.9, 0.3)]
    hp2 = [tt(v, True) for v in (0.7, 1.9, 0.3)]
    ssm = mf.StateSpaceModel(*_matern32_ssm_torch(hp1[0], hp1[1], tp))
    ll = mf.KalmanFilter(ssm, mf.EmissionModel(h), y[:, None], hp1[2].reshape(1, 1)).log_likelihood()
    want = _dense_gp_loglik(hp2[0], hp2[1], hp2[2], tp, y)
    assert abs(float(ll) - float(want)) < 1e-10 * abs(float(want))
    ll.backward()
    want.backward()
    for a1, a2 in zip(hp1, hp2):
        assert abs(float(a1.grad) - float(a2.grad)) < 1e-8 * abs(float(a2.grad))
    # and the forward-only fused kernel agrees with the differentiable composition
    with torch.no_grad():
        fused = mf.KalmanFilter(ssm, mf.EmissionModel(h), y[:, None], hp1[2].reshape(1, 1)).log_likelihood()
    assert abs(float(fused) - float(ll)) < 1e-10 * abs(float(ll))


def test_transform_gradients_match_torch_restatements(batch_shape):
    """ssm_to_expectations / ssm_to_naturals / expectations_to_ssm_params under autograd: the values equal the
    forward-only kernels and the gradients equal torch reverse mode through an explicit propagation."""
    import markovflow_b200 as mf

    np.random.seed(11)
    d, n = 2, 6
    arrays = random_ssm_arrays(batch_shape, n, d)
    p1 = [tt(x, True) for x in arrays]
    p2 = [tt(x, True) for x in arrays]
    ssm = mf.StateSpaceModel(*p1)
    etas = mf.ssm_to_expectations(ssm)
    with torch.no_grad():
        etas_fwd = mf.ssm_to_expectations(mf.StateSpaceModel(*(x.detach() for x in p1)))
    for e, f in zip(etas, etas_fwd):
        assert rel(e, f) < 1e-12
    tm, tc, tsub = _propagate_torch(p2[0], torch.tril(p2[1]), p2[2], p2[3], torch.tril(p2[4]))
    mu = tm[..., None]
    etas_t = (tm, tc + mu @ mu.transpose(-1, -2), tsub + mu[..., 1:, :, :] @ mu[..., :-1, :, :].transpose(-1, -2))
    rng = np.random.default_rng(0)
    ws = [tt(rng.standard_normal(tuple(e.shape))) for e in etas]
    sum(((w * e).sum() for w, e in zip(ws, etas))).backward()
    sum(((w * e).sum() for w, e in zip(ws, etas_t))).backward()
    for q1, q2 in zip(p1, p2):
        assert rel(q1.grad, q2.grad) < GTOL
    # expectations -> ssm: round trip is the identity, so the Jacobian product of the two is too
    e_in = [e.detach().requires_grad_(True) for e in etas]
    back = mf.expectations_to_ssm_params(*e_in)  # (As, offsets, chol_P0, chol_Qs, mu0)
    for got, want in zip(back, (arrays[2], arrays[3], arrays[1], arrays[4], arrays[0])):
        assert rel(got, tt(want)) < 1e-9
    g_out = [tt(rng.standard_normal(tuple(o.shape))) for o in back]
    g_out[2], g_out[3] = torch.tril(g_out[2]), torch.tril(g_out[3])
    g_eta = torch.autograd.grad(back, e_in, grad_outputs=g_out)
    # J_{ssm<-eta}^T g pushed through J_{eta<-ssm}^T must give g back (chain rule of the identity map)
    p3 = [tt(x, True) for x in arrays]
    etas3 = mf.ssm_to_expectations(mf.StateSpaceModel(*p3))
    g_ssm = torch.autograd.grad(etas3, p3, grad_outputs=list(g_eta))
    for got, want in zip(g_ssm, (g_out[4], g_out[2], g_out[0], g_out[1], g_out[3])):
        assert rel(got, want) < 1e-8


def test_natgrad_gets_the_optimal_elbo_in_one_iteration():
    """tests/integration/test_ssm_natgrad.py:47-70: with a Gaussian likelihood, ONE natural-gradient step with
    gamma = 1 moves q to the exact posterior, so the ELBO equals the GPR log marginal likelihood."""
    import markovflow_b200 as mf

    rng = np.random.default_rng(5)
    bsz, t = 3, 10
    tp = tt(np.linspace(0.0, 1.0, t))
    noise_var = 0.01
    prior_arrays = _matern32_ssm_torch(tt(0.3), tt(0.1), tp)
    prior = mf.StateSpaceModel(*(x.detach().expand((bsz,) + tuple(x.shape)).contiguous() for x in prior_arrays))
    y = tt(rng.standard_normal((bsz, t, 1)))
    h = torch.zeros(t, 1, 2, dtype=torch.float64, device=dev())
    h[..., 0] = 1.0
    em = mf.EmissionModel(h)
    gpr_ll = mf.KalmanFilter(prior, em, y, tt([[math.sqrt(noise_var)]])).log_likelihood_per_chain()

    q = prior.create_trainable_copy()

    def elbo():
        mean, cov = q.marginals
        f_mean, f_var = em.project_state_marginals_to_f(mean, cov)
        ve = -0.5 * math.log(2 * math.pi * noise_var) - 0.5 * ((y - f_mean) ** 2 + f_var) / noise_var
        return ve.sum((-1, -2)) - q.kl_divergence(prior)

    before = elbo().detach()
    assert float((gpr_ll - before).min()) > 1.0  # the prior is a poor posterior
    mf.SSMNaturalGradient(gamma=1.0, momentum=False).minimize(lambda: -elbo().sum(), q)
    after = elbo().detach()
    np.testing.assert_allclose(after.cpu().numpy(), gpr_ll.cpu().numpy(), atol=1e-5, rtol=1e-6)
    # q is now the posterior state-space model the Kalman filter computes
    post = mf.KalmanFilter(prior, em, y, tt([[math.sqrt(noise_var)]])).posterior_state_space_model()
    pm, pc = post.marginals
    with torch.no_grad():
        qm, qc = q.marginals
    assert rel(qm, pm) < 1e-7 and rel(qc, pc) < 1e-7


def _nat_to_ssm_torch_loop(th_lin, th_diag, th_sub):
    """The backward U D U^T recursion of naturals_to_ssm_params, step by step in torch (dense autograd reference)."""
    bsz, t, d = th_lin.shape
    offs, chols, a_s = [None] * t, [None] * t, [None] * (t - 1)
    z = c = None
    for k in range(t - 1, -1, -1):
        dk = -2.0 * th_diag[:, k]
        dk = torch.tril(dk) + torch.tril(dk, -1).transpose(-1, -2)
        th = th_lin[:, k]
        if k + 1 < t:
            a = torch.cholesky_solve(th_sub[:, k], c)
            dk = dk - th_sub[:, k].transpose(-1, -2) @ a
            dk = torch.tril(dk) + torch.tril(dk, -1).transpose(-1, -2)
            th = th + (a.transpose(-1, -2) @ z[..., None])[..., 0]
            a_s[k] = a
        z = th
        c = torch.linalg.cholesky(dk)
        offs[k] = torch.cholesky_solve(z[..., None], c)[..., 0]
        q = torch.cholesky_inverse(c)
        chols[k] = torch.linalg.cholesky(torch.tril(q) + torch.tril(q, -1).transpose(-1, -2))
    return torch.stack(a_s, 1), torch.stack(offs, 1), torch.stack(chols, 1)


@pytest.mark.parametrize("smoothing", [True, False])
@pytest.mark.parametrize("b,t,d", [(2, 6, 1), (3, 5, 2), (2, 7, 3), (1, 4, 5)])
def test_naturals_to_ssm_params_reverse_mode(b, t, d, smoothing):
    """naturals_to_ssm_params under autograd (ssm_natgrad.py:173-176 needs it): values equal the forward kernel,
    gradients equal torch's reverse mode through the step-by-step recursion."""
    import markovflow_b200 as mf
    from oracle import np_oracle as O

    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
    to_nat = O.ssm_to_naturals if smoothing else O.ssm_to_naturals_no_smoothing
    th_np = to_nat(O.SSM(*arrays))
    fn = mf.naturals_to_ssm_params if smoothing else mf.naturals_to_ssm_params_no_smoothing
    th1 = [tt(x, True) for x in th_np]
    th2 = [tt(x, True) for x in th_np]
    got = fn(*th1)  # (As, offsets, chol_P0, chol_Qs, mu0)
    with torch.no_grad():
        fwd = fn(*(x.detach() for x in th1))
    for g, f in zip(got, fwd):
        assert rel(g, f) < 1e-11
    if smoothing:
        a, off, chol = _nat_to_ssm_torch_loop(*th2)
    else:
        dk = -2.0 * th2[1]
        dk = torch.tril(dk) + torch.tril(dk, -1).transpose(-1, -2)
        c = torch.linalg.cholesky(dk)
        off = torch.cholesky_solve(th2[0][..., None], c)[..., 0]
        a = torch.cholesky_solve(th2[2], c[:, 1:])
        q = torch.cholesky_inverse(c)
        chol = torch.linalg.cholesky(torch.tril(q) + torch.tril(q, -1).transpose(-1, -2))
    want = (a, off[:, 1:], chol[:, 0], chol[:, 1:], off[:, 0])
    rng = np.random.default_rng(b + t + d)
    ws = [tt(rng.standard_normal(tuple(o.shape))) for o in got]
    ws[2], ws[3] = torch.tril(ws[2]), torch.tril(ws[3])
    sum(((w * o).sum() for w, o in zip(ws, got))).backward()
    sum(((w * o).sum() for w, o in zip(ws, want))).backward()
    for x1, x2, name in zip(th1, th2, ("theta_lin", "theta_diag", "theta_sub")):
        g1, g2 = x1.grad, x2.grad
        if name == "theta_diag":  # symmetric parameter: compare the symmetrised gradients
            g1, g2 = g1 + g1.transpose(-1, -2), g2 + g2.transpose(-1, -2)
        assert rel(g1, g2) < GTOL, name


@pytest.mark.parametrize("b,t,d", [(2, 6, 1), (3, 5, 2), (2, 4, 4)])
def test_sample_reverse_mode(b, t, d):
    """Reparameterised samples are differentiable in the model's parameters and in the draws: gradients equal
    torch's reverse mode through the step-by-step affine recursion; sample(seed=...) under grad returns the
    same trajectories as without."""
    import markovflow_b200 as mf

    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))  # mu0, chol_p0, a, b, chol_q
    rng = np.random.default_rng(3 * b + t + d)
    eps_np = rng.standard_normal((4, b, t, d))
    p1 = [tt(x, True) for x in arrays]
    p2 = [tt(x, True) for x in arrays]
    e1, e2 = tt(eps_np, True), tt(eps_np, True)
    x1 = mf.StateSpaceModel(*p1).sample_from_epsilons(e1)
    mu0, l0, a, bb, lq = p2
    xs = [mu0 + (torch.tril(l0) @ e2[:, :, 0, :, None])[..., 0]]
    for k in range(1, t):
        xs.append((a[:, k - 1] @ xs[-1][..., None])[..., 0] + bb[:, k - 1]
                  + (torch.tril(lq[:, k - 1]) @ e2[:, :, k, :, None])[..., 0])
    x2 = torch.stack(xs, dim=2)
    assert rel(x1, x2) < 1e-12
    w = tt(rng.standard_normal(tuple(x1.shape)))
    (w * x1).sum().backward()
    (w * x2).sum().backward()
    for q1, q2 in zip(p1 + [e1], p2 + [e2]):
        assert rel(q1.grad, q2.grad) < GTOL
    ssm = mf.StateSpaceModel(*(tt(x, True) for x in arrays))
    s_grad = ssm.sample((3,), seed=11)
    with torch.no_grad():
        s_plain = mf.StateSpaceModel(*(tt(x) for x in arrays)).sample((3,), seed=11)
    assert s_grad.requires_grad and rel(s_grad, s_plain) < 1e-12


def test_natgrad_with_momentum_follows_the_reference_update():
    """ssm_natgrad.py:173-203: Adam-style step in the naturals.  One step is compared with the update computed
    from the formulas with dense-autograd gradients; a few steps on a Gaussian-likelihood ELBO increase it."""
    import markovflow_b200 as mf

    rng = np.random.default_rng(9)
    bsz, t = 2, 8
    tp = tt(np.linspace(0.0, 1.0, t))
    noise_var = 0.05
    prior_arrays = _matern32_ssm_torch(tt(0.3), tt(0.5), tp)
    prior = mf.StateSpaceModel(*(x.detach().expand((bsz,) + tuple(x.shape)).contiguous() for x in prior_arrays))
    y = tt(rng.standard_normal((bsz, t, 1)))
    h = torch.zeros(t, 1, 2, dtype=torch.float64, device=dev())
    h[..., 0] = 1.0
    em = mf.EmissionModel(h)
    q = prior.create_trainable_copy()

    def elbo():
        mean, cov = q.marginals
        f_mean, f_var = em.project_state_marginals_to_f(mean, cov)
        ve = -0.5 * math.log(2 * math.pi * noise_var) - 0.5 * ((y - f_mean) ** 2 + f_var) / noise_var
        return ve.sum((-1, -2)) - q.kl_divergence(prior)

    # expected first step, from the formulas
    gamma, b1, b2, eps_ = 0.1, 0.9, 0.99, 1e-8
    loss = -elbo().sum()
    params = q.trainable_variables
    dl = list(torch.autograd.grad(loss, params))
    dl[2], dl[3] = torch.tril(dl[2]), torch.tril(dl[3])
    det = lambda s: mf.StateSpaceModel(s._mu_0.detach(), s._chol_P_0.detach(), s._A_s.detach(), s._b_s.detach(),
                                       s._chol_Q_s.detach())
    etas = [e.detach().requires_grad_(True) for e in mf.ssm_to_expectations(det(q))]
    g_eta = torch.autograd.grad(mf.expectations_to_ssm_params(*etas), etas, grad_outputs=dl)
    thetas = [th.detach().requires_grad_(True) for th in mf.ssm_to_naturals(det(q))]
    a_, off_, chol_ = _nat_to_ssm_torch_loop(*thetas)
    g_theta = torch.autograd.grad((a_, off_[:, 1:], chol_[:, 0], chol_[:, 1:], off_[:, 0]), thetas, grad_outputs=dl)
    lr = gamma * math.sqrt(1 - b2) / (1 - b1)
    comps = [float((g * gt).sum()) for g, gt in zip(g_eta, g_theta)]
    comps[-1] *= 2.0
    v = (1 - b2) * sum(comps)
    theta_new = [th.detach() - lr * (1 - b1) * g / (math.sqrt(v) + eps_) for th, g in zip(thetas, g_eta)]
    want = mf.naturals_to_ssm_params(*theta_new)

    opt = mf.SSMNaturalGradient(gamma=gamma, momentum=True, beta1=b1, beta2=b2, epsilon=eps_)
    before = float(elbo().sum())
    opt.minimize(lambda: -elbo().sum(), q)
    for p, w in zip(q.trainable_variables, want):
        assert rel(p.detach(), w) < 1e-8
    for _ in range(5):
        opt.minimize(lambda: -elbo().sum(), q)
    assert float(elbo().sum()) > before
