"""State dimensions 8 < D <= 32 on the warp-per-chain kernels (csrc/mid_kernels.cuh): the natural / expectation
parameter transforms and the moment recursion, against the numpy oracle (float64 1e-10, float32 1e-4 by
tests/helpers.py::assert_parity), plus the reference's own D = 30, T = 1001 round-trip test
(/root/reference/tests/unit/test_ssm_gaussian_transformations.py:40-103: Sum of ten Matern52 kernels on
linspace(0, 1, 1001); tolerances rtol 1e-7 / atol 1e-6 as written there)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import assert_parity, ld, random_ssm_arrays

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def tt(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda:0").to(dtype)


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def f32r(arrays):
    return tuple(np.asarray(a).astype(np.float32).astype(np.float64) for a in arrays)


def case(b, n, d, dtype, seed):
    state = np.random.get_state()
    np.random.seed(seed)
    arrays = random_ssm_arrays((b,), n, d, scale_a=0.6 / np.sqrt(d))
    np.random.set_state(state)
    if dtype == torch.float32:
        arrays = f32r(arrays)
    return arrays


def adj(fn, arrays_or_inputs, i, dtype):
    """truth= / peer= keyword for assert_parity: the oracle function in long double / float32 on the same inputs."""
    if dtype == torch.float64:
        return dict(truth=lambda: fn(*ld(*arrays_or_inputs))[i])
    return dict(peer=lambda: fn(*(np.asarray(x).astype(np.float32) for x in arrays_or_inputs))[i])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("b,n,d", [(3, 7, 9), (2, 5, 12), (2, 6, 17), (1, 4, 30), (2, 3, 32)])
def test_transforms_match_oracle_above_eight_dimensions(b, n, d, dtype):
    import markovflow_b200 as mf

    arrays = case(b, n, d, dtype, 100 * d + n)
    ssm = mf.StateSpaceModel(*(tt(a, dtype) for a in arrays))
    tol = TOL[dtype]

    def o_ssm(*xs):
        return O.SSM(*xs)

    for name, fn, ofn in (("ssm_to_expectations", mf.ssm_to_expectations, O.ssm_to_expectations),
                          ("ssm_to_naturals", mf.ssm_to_naturals, O.ssm_to_naturals),
                          ("ssm_to_naturals_no_smoothing", mf.ssm_to_naturals_no_smoothing,
                           O.ssm_to_naturals_no_smoothing)):
        want = ofn(O.SSM(*arrays))
        for i, (g, w) in enumerate(zip(fn(ssm), want)):
            assert_parity(npy(g), w, tol, what=f"{name}[{i}] D={d}",
                          **adj(lambda *xs: ofn(o_ssm(*xs)), arrays, i, dtype))
    mu, cov = ssm.marginals
    assert_parity(npy(mu), O.ssm_marginal_means(O.SSM(*arrays)), tol, what=f"marginal means D={d}",
                  **adj(lambda *xs: (O.ssm_marginal_means(o_ssm(*xs)),), arrays, 0, dtype))
    assert_parity(npy(cov), O.ssm_marginal_covariances(O.SSM(*arrays)), tol, what=f"marginal covariances D={d}",
                  **adj(lambda *xs: (O.ssm_marginal_covariances(o_ssm(*xs)),), arrays, 0, dtype))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("b,n,d", [(3, 7, 9), (2, 6, 17), (1, 4, 30), (2, 3, 32)])
def test_inverse_transforms_match_oracle_above_eight_dimensions(b, n, d, dtype):
    import markovflow_b200 as mf

    arrays = case(b, n, d, dtype, 200 * d + n)
    ref = O.SSM(*arrays)
    tol = TOL[dtype]
    names = ["As", "offsets", "chol_P0", "chol_Qs", "mu0"]
    for label, to_params, o_to_params, o_from in (
            ("naturals", mf.naturals_to_ssm_params, O.naturals_to_ssm_params, O.ssm_to_naturals),
            ("naturals_no_smoothing", mf.naturals_to_ssm_params_no_smoothing,
             O.naturals_to_ssm_params_no_smoothing, O.ssm_to_naturals_no_smoothing),
            ("expectations", mf.expectations_to_ssm_params, O.expectations_to_ssm_params, O.ssm_to_expectations)):
        inputs = o_from(ref)
        if dtype == torch.float32:
            inputs = f32r(inputs)
        want = o_to_params(*inputs)
        got = to_params(*(tt(x, dtype) for x in inputs))
        for i, (g, w) in enumerate(zip(got, want)):
            assert_parity(npy(g), w, tol, what=f"{label} -> {names[i]} D={d}", **adj(o_to_params, inputs, i, dtype))


def test_failure_is_reported_above_eight_dimensions():
    import markovflow_b200 as mf

    arrays = case(2, 5, 12, torch.float64, 7)
    th = list(O.ssm_to_naturals(O.SSM(*arrays)))
    th[1] = th[1].copy()
    th[1][1, 3] *= -1.0
    with pytest.raises(mf.CholeskyError):
        mf.naturals_to_ssm_params(*(tt(x) for x in th))
    with pytest.raises(mf.CholeskyError):
        mf.naturals_to_ssm_params_no_smoothing(*(tt(x) for x in th))


@pytest.fixture(scope="module")
def sum_of_matern52_ssm():
    """The set-up of the reference's test (`_setup`, :40-63): ten Matern52(lengthscale 0.01, variance 0.01)."""
    import markovflow_b200 as mf
    from markovflow_b200.kernels import Matern52, Sum

    kern = Sum([Matern52(lengthscale=0.01, variance=0.01) for _ in range(10)])
    x = torch.linspace(0, 1, 1001, dtype=torch.float64, device="cuda:0")
    ssm = kern.state_space_model(x)
    params = (ssm.state_transitions, ssm.state_offsets, ssm.cholesky_initial_covariance,
              ssm.cholesky_process_covariances, ssm.initial_mean)
    assert ssm.state_dim == 30 and ssm.num_transitions == 1000
    return ssm, params


RTOL, ATOL = 1e-7, 1e-6  # the reference's RELATIVE_TOLERANCE / ABSOLUTE_TOLERANCE (:31-32)


def _round_trip(params, back):
    for p, q in zip(params, back):
        np.testing.assert_allclose(npy(p), npy(q), rtol=RTOL, atol=ATOL)


def test_expectation_transformations_d30_t1001(sum_of_matern52_ssm):
    """tests/unit/test_ssm_gaussian_transformations.py:66-77."""
    import markovflow_b200 as mf

    ssm, params = sum_of_matern52_ssm
    _round_trip(params, mf.expectations_to_ssm_params(*mf.ssm_to_expectations(ssm)))


def test_natural_transformations_d30_t1001(sum_of_matern52_ssm):
    """tests/unit/test_ssm_gaussian_transformations.py:80-91."""
    import markovflow_b200 as mf

    ssm, params = sum_of_matern52_ssm
    _round_trip(params, mf.naturals_to_ssm_params(*mf.ssm_to_naturals(ssm)))


def test_natural_transformations_no_smoothing_d30_t1001(sum_of_matern52_ssm):
    """tests/unit/test_ssm_gaussian_transformations.py:94-103."""
    import markovflow_b200 as mf

    ssm, params = sum_of_matern52_ssm
    _round_trip(params, mf.naturals_to_ssm_params_no_smoothing(*mf.ssm_to_naturals_no_smoothing(ssm)))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("b,n,d", [(3, 6, 9), (2, 5, 17), (2, 4, 32)])
def test_block_operators_above_eight_dimensions(b, n, d, dtype):
    """block_diagonal_of_inverse / upper_diagonal_lower / _build_precision / marginal_means / sample_from_epsilons
    (block_tri_diag.py:331,438-545; state_space_model.py:231-251,298-324,431-483) on the warp-per-chain kernels."""
    import markovflow_b200 as mf

    arrays = case(b, n, d, dtype, 300 * d + n)
    ref = O.SSM(*arrays)
    ssm = mf.StateSpaceModel(*(tt(a, dtype) for a in arrays))
    tol = TOL[dtype]
    hi = O.SSM(*ld(*arrays))
    lo = O.SSM(*(a.astype(np.float32) for a in arrays))

    def kw(fn):
        return dict(truth=lambda: fn(hi)) if dtype == torch.float64 else dict(peer=lambda: fn(lo))

    prec = ssm.precision
    pd_, ps_ = O.ssm_build_precision(ref)
    assert_parity(npy(prec.block_diagonal), pd_, tol, what=f"precision diag D={d}",
                  **kw(lambda s: O.ssm_build_precision(s)[0]))
    assert_parity(npy(prec.block_sub_diagonal), ps_, tol, what=f"precision sub D={d}",
                  **kw(lambda s: O.ssm_build_precision(s)[1]))
    assert_parity(npy(ssm.marginal_means), O.ssm_marginal_means(ref), tol, what=f"marginal means D={d}",
                  **kw(O.ssm_marginal_means))
    # inverse subset of the precision's Cholesky factor = marginal covariances
    chol = prec.cholesky
    want_cov = O.ssm_marginal_covariances(ref)
    got_cov = chol.block_diagonal_of_inverse()
    assert_parity(npy(got_cov), want_cov, max(tol, 1e-9) if dtype == torch.float64 else tol,
                  what=f"block_diagonal_of_inverse D={d}", **kw(O.ssm_marginal_covariances))
    # U D U^T of the precision: U^T = A^{-1} has -A_k below the diagonal ... checked through the identity
    # K = U D U^T on dense matrices
    u, chol_d = prec.upper_diagonal_lower()
    dense_k = npy(prec.to_dense())
    ut = npy(u.to_dense())
    cd = npy(chol_d.to_dense())
    rebuilt = np.swapaxes(ut, -1, -2) @ (cd @ np.swapaxes(cd, -1, -2)) @ ut
    scale = np.abs(dense_k).max()
    assert np.abs(rebuilt - dense_k).max() <= (1e-10 if dtype == torch.float64 else 2e-4) * scale
    # sampling from given draws
    eps = np.random.default_rng(d).standard_normal((2, b, n + 1, d))
    if dtype == torch.float32:
        eps = eps.astype(np.float32).astype(np.float64)
    got = ssm.sample_from_epsilons(tt(eps, dtype))
    assert_parity(npy(got), O.ssm_sample_from_epsilons(ref, eps), tol, what=f"sample D={d}",
                  **kw(lambda s: O.ssm_sample_from_epsilons(s, eps.astype(s.a_s.dtype))))
    # the same Philox stream as below eight dimensions: sample(seed) == sample_from_epsilons(sample_epsilons(seed))
    draws = ssm.sample_epsilons((3,), 5)
    assert tuple(draws.shape) == (3, b, n + 1, d)
    assert torch.equal(ssm.sample((3,), seed=5), ssm.sample_from_epsilons(draws))
    assert abs(float(draws.mean())) < 0.2 and abs(float(draws.std()) - 1.0) < 0.2


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("b,n,d,m", [(3, 8, 9, 1), (2, 6, 17, 2), (1, 5, 30, 1)])
def test_kalman_log_likelihood_above_eight_dimensions(b, n, d, m, dtype):
    import markovflow_b200 as mf

    rng = np.random.default_rng(10 * d + m)
    arrays = case(b, n, d, dtype, 400 * d + n)
    h = rng.standard_normal((n + 1, m, d))
    y = rng.standard_normal((b, n + 1, m))
    lr = np.tril(rng.standard_normal((m, m))) * 0.3 + np.eye(m)
    if dtype == torch.float32:
        h, y, lr = f32r((h, y, lr))
    want = O.kalman_log_likelihood(O.SSM(*arrays), h, y, O._r_inv_from_chol(lr), per_chain=True)
    ssm = mf.StateSpaceModel(*(tt(a, dtype) for a in arrays))
    got = mf.kalman_log_likelihood(ssm, tt(h, dtype), tt(y, dtype), tt(lr, dtype))
    kw = (dict(truth=lambda: O.kalman_log_likelihood(O.SSM(*ld(*arrays)), *ld(h, y), O._r_inv_from_chol(ld(lr)[0]),
                                                     per_chain=True))
          if dtype == torch.float64 else
          dict(peer=lambda: O.kalman_log_likelihood(O.SSM(*(a.astype(np.float32) for a in arrays)),
                                                    h.astype(np.float32), y.astype(np.float32),
                                                    O._r_inv_from_chol(lr.astype(np.float32)), per_chain=True)))
    assert_parity(npy(got), want, TOL[dtype], what=f"Kalman log-likelihood D={d} m={m}", **kw)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("b,n,d", [(2, 6, 9), (2, 5, 17), (1, 4, 32)])
def test_log_pdf_dense_mult_and_posterior_above_eight_dimensions(b, n, d, dtype):
    """log_pdf (state_space_model.py:485-526), dense_mult (block_tri_diag.py:189) and the posterior state-space
    model of the Kalman filter (kalman_filter.py:109-182: U D U^T + dense_mult + solves + block maps)."""
    import markovflow_b200 as mf

    rng = np.random.default_rng(d)
    arrays = case(b, n, d, dtype, 500 * d + n)
    ref = O.SSM(*arrays)
    ssm = mf.StateSpaceModel(*(tt(a, dtype) for a in arrays))
    tol = TOL[dtype]
    hi = O.SSM(*ld(*arrays))
    lo = O.SSM(*(a.astype(np.float32) for a in arrays))
    states = rng.standard_normal((3, b, n + 1, d))
    if dtype == torch.float32:
        states = states.astype(np.float32).astype(np.float64)
    kw = (dict(truth=lambda: O.ssm_log_pdf(hi, ld(states)[0])) if dtype == torch.float64 else
          dict(peer=lambda: O.ssm_log_pdf(lo, states.astype(np.float32))))
    assert_parity(npy(ssm.log_pdf(tt(states, dtype))), O.ssm_log_pdf(ref, states), tol, what=f"log_pdf D={d}", **kw)
    # dense_mult of the precision with a vector = K^{-1} x on dense matrices
    prec = ssm.precision
    x = rng.standard_normal((b, n + 1, d))
    if dtype == torch.float32:
        x = x.astype(np.float32).astype(np.float64)
    dense = npy(prec.to_dense())
    want = np.einsum("bij,bj->bi", dense, x.reshape(b, -1)).reshape(b, n + 1, d)
    got = npy(prec.dense_mult(tt(x, dtype)))
    assert np.abs(got - want).max() <= (1e-10 if dtype == torch.float64 else 1e-4) * np.abs(want).max()
    # posterior SSM: marginal means / covariances equal the oracle's smoothed moments
    m = 1
    h = rng.standard_normal((n + 1, m, d))
    y = rng.standard_normal((b, n + 1, m))
    lr = np.array([[0.7]])
    if dtype == torch.float32:
        h, y = f32r((h, y))
    kf = mf.KalmanFilter(ssm, mf.EmissionModel(tt(h, dtype)), tt(y, dtype), tt(lr, dtype))
    post = kf.posterior_state_space_model()
    o_post = O.kalman_posterior_ssm(ref, h, y, O._r_inv_from_chol(lr))
    ptol = 1e-9 if dtype == torch.float64 else 1e-3
    mu, cov = post.marginals
    assert_parity(npy(mu), O.ssm_marginal_means(o_post), ptol, what=f"posterior means D={d}",
                  **(dict(truth=lambda: O.ssm_marginal_means(O.kalman_posterior_ssm(hi, *ld(h, y), O._r_inv_from_chol(ld(lr)[0]))))
                     if dtype == torch.float64 else {}))
    assert_parity(npy(cov), O.ssm_marginal_covariances(o_post), ptol, what=f"posterior covariances D={d}",
                  **(dict(truth=lambda: O.ssm_marginal_covariances(O.kalman_posterior_ssm(hi, *ld(h, y), O._r_inv_from_chol(ld(lr)[0]))))
                     if dtype == torch.float64 else {}))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("b,n,d", [(3, 6, 9), (2, 5, 17), (1, 4, 32)])
def test_kl_divergence_above_eight_dimensions(b, n, d, dtype):
    """kl_divergence (state_space_model.py:528-593) between two random models, chain-rule form on the warp-per-chain
    kernel; KL(q || q) = 0."""
    import markovflow_b200 as mf

    qa = case(b, n, d, dtype, 600 * d + n)
    pa = case(b, n, d, dtype, 700 * d + n)
    q = mf.StateSpaceModel(*(tt(a, dtype) for a in qa))
    p = mf.StateSpaceModel(*(tt(a, dtype) for a in pa))
    want = O.ssm_kl_divergence(O.SSM(*qa), O.SSM(*pa))
    kw = (dict(truth=lambda: O.ssm_kl_divergence(O.SSM(*ld(*qa)), O.SSM(*ld(*pa)))) if dtype == torch.float64 else
          dict(peer=lambda: O.ssm_kl_divergence(O.SSM(*(a.astype(np.float32) for a in qa)),
                                                O.SSM(*(a.astype(np.float32) for a in pa)))))
    assert_parity(npy(q.kl_divergence(p)), want, TOL[dtype], what=f"KL D={d}", **kw)
    assert float(q.kl_divergence(q).abs().max()) <= (1e-9 if dtype == torch.float64 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [9, 17, 32])
def test_conditionals_above_eight_dimensions(d, dtype):
    """pairwise_marginals, conditional statistics and conditional prediction (conditionals.py:29-205,380-485): the
    bodies of tests/test_gpu_conditionals.py at state dimensions served by the warp-per-point kernels."""
    from tests import test_gpu_conditionals as tc

    tc.test_pairwise_marginals((2,), 7, d, dtype)
    tc.test_pairwise_marginals((), 3, d, dtype)
    tc.test_conditional_statistics_and_predict(d, dtype)


def test_posterior_process_with_the_config4_kernel():
    """ConditionalProcess.predict_f for the D = 17 Sum(Matern52 + 7 harmonics)-like kernel of BASELINE config 4 against
    the dense GP posterior at new time points (posterior.py:207-258 on top of conditionals.py)."""
    import markovflow_b200 as mf
    from markovflow_b200.kernels import HarmonicOscillator, Matern52, Sum

    rng = np.random.default_rng(4)
    # (the jitter -- 1e-6 on every Q_k and on P0 -- keeps the oscillators' process noise factorisable; it is part of
    #  the model the operators see, hence the tolerances below, as in tests/test_gpu_posterior.py)
    kern = Sum([Matern52(lengthscale=0.7, variance=1.1)] +
               [HarmonicOscillator(variance=0.3 / (i + 1), period=1.0 / (i + 1)) for i in range(7)], jitter=1e-6)
    assert kern.state_dim == 17
    x = torch.as_tensor(np.sort(rng.uniform(0.0, 3.0, size=25)), device="cuda:0")
    y = torch.as_tensor(rng.standard_normal((25, 1)), device="cuda:0")
    noise = 0.2
    ssm = kern.state_space_model(x)
    em = kern.generate_emission_model(x)
    kf = mf.KalmanFilter(ssm, em, y, torch.as_tensor([[noise ** 0.5]], device="cuda:0", dtype=torch.float64))
    post = mf.ConditionalProcess(kf.posterior_state_space_model(), kern, x)
    xs = torch.as_tensor(np.sort(rng.uniform(0.2, 2.8, size=9)), device="cuda:0")
    mean, var = post.predict_f(xs)
    # dense GP regression with the same covariance function
    xn, xsn, yn = x.cpu().numpy(), xs.cpu().numpy(), y.cpu().numpy()

    def k(a, b):
        r = np.abs(a[:, None] - b[None, :])
        s5 = np.sqrt(5.0) * r / 0.7
        out = 1.1 * (1 + s5 + s5 ** 2 / 3.0) * np.exp(-s5)
        for i in range(7):
            out = out + (0.3 / (i + 1)) * np.cos(2 * np.pi * (i + 1) * r)
        return out

    kxx = k(xn, xn) + noise * np.eye(25)
    ksx = k(xsn, xn)
    want_mean = ksx @ np.linalg.solve(kxx, yn)
    want_var = np.diag(k(xsn, xsn) - ksx @ np.linalg.solve(kxx, ksx.T))
    inside = (xsn > xn[0]) & (xsn < xn[-1])
    assert inside.sum() >= 5
    got_mean, got_var = npy(mean).reshape(-1), npy(var).reshape(-1)
    # eight components each carry the jitter: the dense answer (no jitter) is matched to 1e-3
    assert np.abs(got_mean - want_mean[:, 0])[inside].max() < 1e-3 * max(1.0, np.abs(want_mean).max())
    assert np.abs(got_var - want_var)[inside].max() < 1e-3 * max(1.0, np.abs(want_var).max())
