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
