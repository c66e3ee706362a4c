"""Pin the oracle's StateSpaceModel / KalmanFilter / transform restatements.

Known answers: the golden vectors produced by the reference's own ``NumpyKalmanFilter``
(``tests/golden/make_golden.py``), dense Gaussians built by direct propagation (what
``tests/unit/test_state_space_model.py:40-233`` checks against), and round trips
(``tests/unit/test_ssm_gaussian_transformations.py:64-103``)."""
import glob
import os

import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import dense_ssm_mean_cov, random_ssm_arrays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kalman_case(path):
    g = np.load(path)
    batch = tuple(int(x) for x in g["batch_shape"]) if "batch_shape" in g else ()
    d = g["A"].shape[0]
    m = g["H"].shape[0]
    t = g["log_liks"].shape[-1]
    n = t - 1
    ssm = O.SSM(
        np.broadcast_to(g["mu0"], batch + (d,)).copy(),
        np.broadcast_to(g["chol_P0"], batch + (d, d)).copy(),
        np.broadcast_to(g["A"], batch + (n, d, d)).copy(),
        np.broadcast_to(g["b"], batch + (n, d)).copy(),
        np.broadcast_to(g["chol_Q"], batch + (n, d, d)).copy(),
    )
    h = np.broadcast_to(g["H"], batch + (t, m, d)).copy()
    return g, ssm, h


@pytest.mark.parametrize("tag", ["b3", "b0", "b21", "d5m1"])
def test_log_likelihood_matches_reference_kalman_filter(tag):
    """tests/integration/test_kalman_filter.py:131-139 with the reference's numpy filter as truth."""
    g, ssm, h = load_kalman_case(os.path.join(GOLDEN, f"kalman_{tag}.npz"))
    r_inv = O._r_inv_from_chol(g["chol_R"])
    ll = O.kalman_log_likelihood(ssm, h, g["y"], r_inv)
    np.testing.assert_allclose(ll, np.sum(g["log_liks"]), rtol=1e-7)  # reference tolerance (default rtol)
    per_chain = O.kalman_log_likelihood(ssm, h, g["y"], r_inv, per_chain=True)
    np.testing.assert_allclose(per_chain, np.sum(g["log_liks"], axis=-1), rtol=1e-7)


@pytest.mark.parametrize("tag", ["b3", "b0", "b21", "d5m1"])
def test_posterior_ssm_matches_reference_rts_smoother(tag):
    """tests/integration/test_kalman_filter.py:105-128."""
    g, ssm, h = load_kalman_case(os.path.join(GOLDEN, f"kalman_{tag}.npz"))
    r_inv = O._r_inv_from_chol(g["chol_R"])
    post = O.kalman_posterior_ssm(ssm, h, g["y"], r_inv)
    np.testing.assert_allclose(O.ssm_marginal_means(post), g["smooth_means"], rtol=1e-7, atol=1e-9)
    covs = O.ssm_marginal_covariances(post)
    np.testing.assert_allclose(
        *np.broadcast_arrays(g["smooth_covs"], covs), rtol=1e-7, atol=1e-9
    )


def test_sites_log_likelihood_and_posterior_match_reference():
    """tests/integration/test_kalman_filter_with_sites.py."""
    g = np.load(os.path.join(GOLDEN, "kalman_sites_t10.npz"))
    d = g["A"].shape[0]
    t = g["nat1"].shape[0]
    ssm = O.SSM(
        g["mu0"], g["chol_P0"], np.broadcast_to(g["A"], (t - 1, d, d)).copy(),
        np.broadcast_to(g["b"], (t - 1, d)).copy(), np.broadcast_to(g["chol_Q"], (t - 1, d, d)).copy(),
    )
    h = np.broadcast_to(g["H"], (t, 1, d)).copy()
    means, precisions, _ = O.sites_means_precisions(g["nat1"], g["nat2"])
    np.testing.assert_allclose(means, g["site_means"])
    np.testing.assert_allclose(precisions, 1.0 / g["site_vars"])
    ll = O.kalman_log_likelihood(ssm, h, means, precisions)
    np.testing.assert_allclose(ll, np.sum(g["log_liks"]), rtol=1e-7)  # reference tolerance (default rtol)
    post = O.kalman_posterior_ssm(ssm, h, means, precisions)
    np.testing.assert_allclose(O.ssm_marginal_means(post), g["smooth_means"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(O.ssm_marginal_covariances(post), g["smooth_covs"], rtol=1e-7, atol=1e-9)


def test_time_varying_filter_agrees_with_golden_and_spingp():
    g, ssm, h = load_kalman_case(os.path.join(GOLDEN, "kalman_b0.npz"))
    r = g["chol_R"] @ g["chol_R"].T
    q = ssm.chol_q_s @ np.swapaxes(ssm.chol_q_s, -1, -2)
    lls, fm, fp = O.kalman_filter_time_varying(
        ssm.mu0, ssm.chol_p0 @ ssm.chol_p0.T, ssm.a_s, ssm.b_s, q, h, r, g["y"]
    )
    np.testing.assert_allclose(lls, g["log_liks"], rtol=1e-9)
    np.testing.assert_allclose(fm, g["filter_means"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(fp, g["filter_covs"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("kernel_name", ["m32", "m52", "sum"])
def test_irregular_time_varying_loglik_spingp_equals_filter_equals_pscan(kernel_name):
    rng = np.random.default_rng(5)
    if kernel_name == "m32":
        k = O.Matern32(1.3, 0.7)
    elif kernel_name == "m52":
        k = O.Matern52(0.9, 1.4)
    else:
        k = O.Sum([O.Matern52(1.0, 1.0)] + [O.HarmonicOscillator(0.5 ** j, 1.0 / j) for j in (1, 2)],
                  jitter=1e-6)
    t = 40
    tp = np.cumsum(rng.uniform(0.05, 0.3, size=t))
    ssm = k.state_space_model(tp)
    h = k.emission_matrix(tp)
    y = rng.standard_normal((t, 1))
    chol_r = np.array([[0.3]])
    ll = O.kalman_log_likelihood(ssm, h, y, O._r_inv_from_chol(chol_r))
    q = ssm.chol_q_s @ np.swapaxes(ssm.chol_q_s, -1, -2)
    p0 = ssm.chol_p0 @ ssm.chol_p0.T
    lls, _, _ = O.kalman_filter_time_varying(ssm.mu0, p0, ssm.a_s, ssm.b_s, q, h, chol_r @ chol_r.T, y)
    np.testing.assert_allclose(ll, np.sum(lls), rtol=1e-9)
    ll_scan = O.pscan_log_likelihood(ssm.mu0, p0, ssm.a_s, ssm.b_s, q, h, chol_r @ chol_r.T, y, segment=7)
    np.testing.assert_allclose(ll_scan, np.sum(lls), rtol=1e-9)


def test_pscan_combine_is_associative_and_prefix_is_filter():
    rng = np.random.default_rng(3)
    k = O.Matern32(1.0, 1.0)
    t = 12
    tp = np.cumsum(rng.uniform(0.05, 0.3, size=t))
    ssm = k.state_space_model(tp)
    h = k.emission_matrix(tp)
    y = rng.standard_normal((t, 1))
    r = np.array([[0.05]])
    q = ssm.chol_q_s @ np.swapaxes(ssm.chol_q_s, -1, -2)
    p0 = ssm.chol_p0 @ ssm.chol_p0.T
    els = O.pscan_elements(ssm.mu0, p0, ssm.a_s, ssm.b_s, q, h, r, y)
    e = [tuple(x[i] for x in els) for i in range(t)]
    left = O.pscan_combine(O.pscan_combine(e[3], e[4]), e[5])
    right = O.pscan_combine(e[3], O.pscan_combine(e[4], e[5]))
    for a, b in zip(left, right):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    _, fm, fp = O.kalman_filter_time_varying(ssm.mu0, p0, ssm.a_s, ssm.b_s, q, h, r, y)
    acc = e[0]
    for i in range(1, t):
        acc = O.pscan_combine(acc, e[i])
        np.testing.assert_allclose(acc[1], fm[i], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(acc[2], fp[i], rtol=1e-9, atol=1e-12)


# ---- StateSpaceModel against a dense Gaussian (tests/unit/test_state_space_model.py) ----------


def _single(ssm_arrays, idx=None):
    return ssm_arrays if idx is None else tuple(a[idx] for a in ssm_arrays)


def test_ssm_precision_means_covariances_logdet_vs_dense(state_dim, transitions):
    arrs = random_ssm_arrays((2,), transitions, state_dim)
    ssm = O.SSM(*arrs)
    pd, ps = O.ssm_build_precision(ssm)
    prec = O.btd_to_dense(pd, ps, symmetric=True)
    means = O.ssm_marginal_means(ssm)
    covs = O.ssm_marginal_covariances(ssm)
    sub_covs = O.ssm_subsequent_covariances(ssm, covs)
    logdet = O.ssm_log_det_precision(ssm)
    d, t = state_dim, transitions + 1
    for b in range(2):
        mean_d, cov_d = dense_ssm_mean_cov(*_single(arrs, b))
        np.testing.assert_allclose(prec[b] @ cov_d, np.eye(t * d), atol=1e-7)
        np.testing.assert_allclose(means[b].reshape(-1), mean_d, rtol=1e-10, atol=1e-12)
        for k in range(t):
            np.testing.assert_allclose(covs[b, k], cov_d[k*d:(k+1)*d, k*d:(k+1)*d], rtol=1e-7, atol=1e-9)
        for k in range(t - 1):
            np.testing.assert_allclose(sub_covs[b, k], cov_d[(k+1)*d:(k+2)*d, k*d:(k+1)*d], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(logdet[b], -np.linalg.slogdet(cov_d)[1], rtol=1e-8, atol=1e-9)


def test_ssm_log_pdf_vs_dense(state_dim, transitions):
    arrs = random_ssm_arrays((), transitions, state_dim)
    ssm = O.SSM(*arrs)
    mean_d, cov_d = dense_ssm_mean_cov(*arrs)
    states = np.random.normal(size=(4, transitions + 1, state_dim))
    got = O.ssm_log_pdf(ssm, states)
    diff = states.reshape(4, -1) - mean_d
    want = -0.5 * (
        np.einsum("si,ij,sj->s", diff, np.linalg.inv(cov_d), diff)
        + np.linalg.slogdet(cov_d)[1]
        + mean_d.size * np.log(2 * np.pi)
    )
    np.testing.assert_allclose(got, want, rtol=1e-8)


def test_ssm_kl_vs_dense(state_dim, transitions):
    a1 = random_ssm_arrays((), transitions, state_dim)
    a2 = random_ssm_arrays((), transitions, state_dim)
    m1, c1 = dense_ssm_mean_cov(*a1)
    m2, c2 = dense_ssm_mean_cov(*a2)
    n = m1.size
    c2i = np.linalg.inv(c2)
    want = 0.5 * (
        np.trace(c2i @ c1) + (m2 - m1) @ c2i @ (m2 - m1) - n
        + np.linalg.slogdet(c2)[1] - np.linalg.slogdet(c1)[1]
    )
    got = O.ssm_kl_divergence(O.SSM(*a1), O.SSM(*a2))
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(O.ssm_kl_divergence(O.SSM(*a1), O.SSM(*a1)), 0.0, atol=1e-8)


def test_sample_from_epsilons_near_deterministic():
    """tests/unit/test_sampling_from_ssm.py:54-130: tiny noise => samples follow the mean recursion."""
    mu0, cp0, a_s, b_s, cq = random_ssm_arrays((3,), 5, 2)
    tiny = np.finfo(np.float64).tiny
    ssm = O.SSM(mu0, cp0 * 0 + tiny * np.eye(2), a_s, b_s, cq * 0 + tiny * np.eye(2))
    eps = np.random.normal(size=(4, 3, 6, 2))
    samples = O.ssm_sample_from_epsilons(ssm, eps)
    assert samples.shape == (4, 3, 6, 2)
    np.testing.assert_allclose(samples, np.broadcast_to(O.ssm_marginal_means(ssm), samples.shape))


# ---- transforms (tests/unit/test_ssm_gaussian_transformations.py) -------------------------------


def _ssm_close(a: O.SSM, params, rtol=1e-7, atol=1e-6):
    a_s, offsets, chol_p0, chol_q, mu0 = params
    np.testing.assert_allclose(a.a_s, a_s, rtol=rtol, atol=atol)
    np.testing.assert_allclose(a.b_s, offsets, rtol=rtol, atol=atol)
    np.testing.assert_allclose(a.chol_p0, chol_p0, rtol=rtol, atol=atol)
    np.testing.assert_allclose(a.chol_q_s, chol_q, rtol=rtol, atol=atol)
    np.testing.assert_allclose(a.mu0, mu0, rtol=rtol, atol=atol)


@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5)])
def test_transform_round_trips(batch_shape, d, n):
    ssm = O.SSM(*random_ssm_arrays(batch_shape, n, d))
    _ssm_close(ssm, O.expectations_to_ssm_params(*O.ssm_to_expectations(ssm)))
    _ssm_close(ssm, O.naturals_to_ssm_params(*O.ssm_to_naturals(ssm)))
    _ssm_close(ssm, O.naturals_to_ssm_params_no_smoothing(*O.ssm_to_naturals_no_smoothing(ssm)))


def test_transform_round_trip_sum_of_matern52_d30():
    """The reference's own case: Sum(10 x Matern52), D=30 (we use T=101 to stay fast)."""
    k = O.Sum([O.Matern52(1.0 + 0.1 * i, 1.0 + 0.05 * i) for i in range(10)], jitter=0.0)
    tp = np.linspace(0.0, 10.0, 101)
    ssm = k.state_space_model(tp)
    _ssm_close(ssm, O.expectations_to_ssm_params(*O.ssm_to_expectations(ssm)), rtol=1e-5, atol=1e-5)
    _ssm_close(ssm, O.naturals_to_ssm_params(*O.ssm_to_naturals(ssm)), rtol=1e-5, atol=1e-5)


def test_naturals_are_precision_and_precision_times_mean():
    arrs = random_ssm_arrays((), 4, 2)
    ssm = O.SSM(*arrs)
    mean_d, cov_d = dense_ssm_mean_cov(*arrs)
    th_lin, th_diag, th_sub = O.ssm_to_naturals(ssm)
    prec = O.btd_to_dense(-2.0 * th_diag, -th_sub, symmetric=True)
    np.testing.assert_allclose(prec, np.linalg.inv(cov_d), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(th_lin.reshape(-1), np.linalg.solve(cov_d, mean_d), rtol=1e-6, atol=1e-8)


# ---- closed-form kernels: SSM-implied covariance equals the kernel function ------------------


def test_matern_and_harmonic_ssm_covariance_equals_kernel_function():
    tp = np.array([0.0, 0.3, 0.35, 1.2, 2.0])
    r = np.abs(tp[:, None] - tp[None, :])

    def f_cov(k):
        ssm = k.state_space_model(tp)
        _, cov = dense_ssm_mean_cov(ssm.mu0, ssm.chol_p0, ssm.a_s, ssm.b_s, ssm.chol_q_s)
        d = k.state_dim
        hrow = k.emission_row()
        hh = np.kron(np.eye(len(tp)), hrow[None, :])
        return hh @ cov @ hh.T

    l, v = 0.8, 1.7
    lam = np.sqrt(3) / l
    np.testing.assert_allclose(f_cov(O.Matern32(l, v)), v * (1 + lam * r) * np.exp(-lam * r), rtol=1e-9)
    lam = np.sqrt(5) / l
    np.testing.assert_allclose(
        f_cov(O.Matern52(l, v)), v * (1 + lam * r + lam ** 2 * r ** 2 / 3) * np.exp(-lam * r), rtol=1e-8
    )
    np.testing.assert_allclose(
        f_cov(O.HarmonicOscillator(v, 0.7, jitter=1e-10)), v * np.cos(2 * np.pi / 0.7 * r), atol=1e-8
    )
    ksum = O.Sum([O.Matern32(l, v), O.HarmonicOscillator(0.5, 0.7)], jitter=1e-10)
    np.testing.assert_allclose(
        f_cov(ksum),
        v * (1 + np.sqrt(3) / l * r) * np.exp(-np.sqrt(3) / l * r) + 0.5 * np.cos(2 * np.pi / 0.7 * r),
        atol=1e-7,
    )


# ---- the kernel -> SSM -> Kalman log-likelihood chain the fused CUDA path replaces ----


@pytest.mark.parametrize("name", ["m12", "m32", "m52"])
def test_matern_kalman_log_likelihood_equals_dense_gp_marginal_likelihood(name):
    """tests/integration/models/test_gaussian_process_regression.py:78-115 restated: the SSM-based
    log-likelihood equals log N(y; 0, K + R) with the closed-form Matern kernel matrix K."""
    rng = np.random.default_rng(71892305)
    tp = np.cumsum(rng.uniform(0.05, 0.4, size=40))
    y = np.sin(tp) + 0.1 * rng.standard_normal(40)
    l, v, noise = 0.7, 1.9, 0.3
    r = np.abs(tp[:, None] - tp[None, :])
    if name == "m12":
        kern, kmat = O.Matern12(l, v), v * np.exp(-r / l)
    elif name == "m32":
        lam = np.sqrt(3) / l
        kern, kmat = O.Matern32(l, v), v * (1 + lam * r) * np.exp(-lam * r)
    else:
        lam = np.sqrt(5) / l
        kern, kmat = O.Matern52(l, v), v * (1 + lam * r + lam ** 2 * r ** 2 / 3) * np.exp(-lam * r)
    ssm = kern.state_space_model(tp)
    h = kern.emission_matrix(tp)
    got = O.kalman_log_likelihood(ssm, h, y[:, None], np.array([[1.0 / noise ** 2]]))
    c = kmat + noise ** 2 * np.eye(40)
    want = -0.5 * (y @ np.linalg.solve(c, y) + np.linalg.slogdet(c)[1] + 40 * np.log(2 * np.pi))
    np.testing.assert_allclose(got, want, rtol=1e-9)


def test_extended_precision_transition_statistics_agree_with_the_float64_restatement():
    tp = np.cumsum(np.random.default_rng(1).uniform(0.05, 0.3, 20))
    for kern, tol in ((O.Matern12(0.7, 1.3, 1e-8), 1e-14), (O.Matern32(0.7, 1.3, 1e-8), 1e-11),
                      (O.Matern52(0.7, 1.3, 1e-8), 1e-8)):
        a, b = kern.state_space_model(tp), O.stationary_ssm_extended_precision(kern, tp)
        np.testing.assert_allclose(a.a_s, b.a_s, rtol=0, atol=1e-15)
        np.testing.assert_allclose(a.chol_q_s, b.chol_q_s, rtol=0, atol=tol)
        np.testing.assert_allclose(a.chol_p0, b.chol_p0, rtol=0, atol=1e-15)
