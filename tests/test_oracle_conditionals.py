"""CPU tests of the oracle's restatement of ``markovflow/conditionals.py`` against
dense Gaussian algebra -- the identities the reference's ``tests/integration/test_posterior.py:196-246``
use for the pairwise marginals, and direct conditioning of the joint for the conditional statistics."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import dense_ssm_mean_cov, random_ssm_arrays


def _spd(rng, d):
    a = rng.standard_normal((d, d))
    return a @ a.T + d * np.eye(d)


@pytest.mark.parametrize("d", [1, 2, 3])
def test_pairwise_marginals_match_the_dense_joint(d):
    np.random.seed(11 + d)
    arrays = random_ssm_arrays((), 5, d)
    ssm = O.SSM(*arrays)
    mean, cov = dense_ssm_mean_cov(*arrays)
    t = 6
    im, ic = np.random.normal(size=d), _spd(np.random.default_rng(d), d)
    jm, jc = O.pairwise_marginals(ssm, im, ic)
    assert jm.shape == (t + 1, 2 * d) and jc.shape == (t + 1, 2 * d, 2 * d)
    for k in range(1, t):  # interior pairs (x_{k-1}, x_k) are blocks of the dense joint
        sl = slice((k - 1) * d, (k + 1) * d)
        np.testing.assert_allclose(jm[k], mean[sl], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(jc[k], cov[sl, sl], rtol=1e-8, atol=1e-10)
    # the two end pairs involve the (independent) initial state
    np.testing.assert_allclose(jm[0], np.concatenate([im, mean[:d]]))
    np.testing.assert_allclose(jc[0][:d, :d], ic)
    np.testing.assert_allclose(jc[0][:d, d:], 0.0)
    np.testing.assert_allclose(jc[t][d:, d:], ic)
    np.testing.assert_allclose(jc[t][:d, :d], cov[-d:, -d:], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("d", [1, 2, 3])
def test_conditional_statistics_equal_direct_conditioning(d):
    """p(x_t | x_-, x_+) from the joint of (x_-, x_t, x_+) under x_t = A1 x_- + e1, x_+ = A2 x_t + e2."""
    rng = np.random.default_rng(5 + d)
    a1, a2 = 0.7 * rng.standard_normal((d, d)), 0.7 * rng.standard_normal((d, d))
    q1, q2, p0 = _spd(rng, d), _spd(rng, d), _spd(rng, d)
    dm, em, tm = O.conditional_statistics_from_transitions(a1, q1, a2, q2)
    _, _, tinv = O.conditional_statistics_from_transitions(a1, q1, a2, q2, return_precision=True)
    np.testing.assert_allclose(tinv, np.linalg.inv(tm), rtol=1e-9)
    # joint covariance of (x_-, x_t, x_+) with x_- ~ N(0, p0)
    s_mm = p0
    s_tm = a1 @ p0
    s_tt = a1 @ p0 @ a1.T + q1
    s_pm = a2 @ s_tm
    s_pt = a2 @ s_tt
    s_pp = a2 @ s_tt @ a2.T + q2
    s_cc = np.block([[s_mm, s_pm.T], [s_pm, s_pp]])  # conditioning set (x_-, x_+)
    s_tc = np.concatenate([s_tm, s_pt.T], axis=1)
    gain = s_tc @ np.linalg.inv(s_cc)
    np.testing.assert_allclose(np.concatenate([dm, em], axis=1), gain, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(tm, s_tt - gain @ s_tc.T, rtol=1e-8, atol=1e-10)


def test_base_conditional_predict():
    rng = np.random.default_rng(2)
    d, n = 2, 7
    proj = rng.standard_normal((n, d, 2 * d))
    tc = np.stack([_spd(rng, d) for _ in range(n)])
    m = rng.standard_normal((n, 2 * d))
    s = np.stack([_spd(rng, 2 * d) for _ in range(n)])
    mean, cov = O.base_conditional_predict(proj, tc, m, s)
    for i in range(n):
        np.testing.assert_allclose(mean[i], proj[i] @ m[i])
        np.testing.assert_allclose(cov[i], tc[i] + proj[i] @ s[i] @ proj[i].T)
    mean2, cov2 = O.base_conditional_predict(proj, tc, m)
    np.testing.assert_allclose(mean2, mean)
    np.testing.assert_allclose(cov2, tc)
