"""Posterior processes, conditional prediction with the reference signature, and in-kernel sampling
 against closed-form dense Gaussian-process answers -- the checks of the reference's
``tests/integration/test_posterior.py`` / ``tests/unit/test_conditionals.py``: predictions from the state-space
posterior equal the dense GP posterior; samples have the predicted moments; with the exact posterior as proposal
every importance weight equals the marginal likelihood."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def tt(x):
    return torch.as_tensor(np.array(x, dtype=np.float64), device=dev())


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def matern_cov(nu_dim, ell, var, a, b):
    r = np.abs(a[:, None] - b[None, :])
    lam = math.sqrt(2 * nu_dim - 1) / ell
    poly = {1: 1.0, 2: 1 + lam * r, 3: 1 + lam * r + lam ** 2 * r ** 2 / 3}[nu_dim]
    return var * poly * np.exp(-lam * r)


class GaussianLikelihood:
    def __init__(self, variance):
        self.variance = variance

    def log_prob(self, f, y):
        return (-0.5 * math.log(2 * math.pi * self.variance) - 0.5 * (y - f) ** 2 / self.variance).sum(-1)

    def predict_mean_and_var(self, f_mean, f_var):
        return f_mean, f_var + self.variance


def _gpr(kernel_cls, d, ell, var, noise, n=40, seed=0, batch=()):
    import markovflow_b200 as mf

    rng = np.random.default_rng(seed)
    # gaps of 0.1 .. 0.3: tiny gaps make Q_k = Pinf - A Pinf A^T cancel catastrophically (Matern52: O(dt^5)
    # against O(1)), in the reference as here, and would blur the comparison with the dense closed form
    tp = np.cumsum(rng.uniform(0.1, 0.3, size=batch + (n,)), axis=-1)
    y = np.sin(tp) + noise * rng.standard_normal(batch + (n,))
    kern = kernel_cls(ell, var)
    prior = kern.state_space_model(tt(tp))
    kf = mf.KalmanFilter(prior, kern.generate_emission_model(tt(tp)), tt(y[..., None]), tt([[noise]]))
    return kern, tp, y, kf


@pytest.mark.parametrize("d", [1, 2, 3])
def test_predict_f_and_state_match_the_dense_gp_posterior(d):
    """ConditionalProcess.predict_f / predict_state on the Kalman posterior (posterior.py:207-258; checked in
    the reference by tests/integration/test_posterior.py) == dense GP regression, inside, between and far
    outside the data."""
    import markovflow_b200 as mf

    ell, var, noise = 0.9, 1.4, 0.2
    kern, tp, y, kf = _gpr({1: mf.Matern12, 2: mf.Matern32, 3: mf.Matern52}[d], d, ell, var, noise)
    post = mf.ConditionalProcess(kf.posterior_state_space_model(), kern, tt(tp))
    new = np.sort(np.concatenate([np.random.default_rng(1).uniform(-3.0, tp[-1] + 3.0, size=25), tp[3:6], [-40.0, 60.0]]))
    f_mean, f_var = post.predict_f(tt(new))
    kxx = matern_cov(d, ell, var, tp, tp) + noise ** 2 * np.eye(len(tp))
    ksx = matern_cov(d, ell, var, new, tp)
    want_mean = ksx @ np.linalg.solve(kxx, y)
    want_var = var - np.einsum("ij,ji->i", ksx, np.linalg.solve(kxx, ksx.T))
    assert tuple(f_mean.shape) == (len(new), 1) and tuple(f_var.shape) == (len(new), 1)
    assert rel(npy(f_mean)[:, 0], want_mean) < 1e-9
    assert rel(npy(f_var)[:, 0], want_var) < 1e-9
    s_mean, s_cov = post.predict_state(tt(new))
    assert tuple(s_cov.shape) == (len(new), d, d)
    assert rel(npy(s_mean)[:, 0], want_mean) < 1e-9 and rel(npy(s_cov)[:, 0, 0], want_var) < 1e-9
    # far outside the data the posterior reverts to the prior
    assert abs(float(f_var[-1, 0]) - var) < 1e-9 and abs(float(f_mean[-1, 0])) < 1e-9
    full = post.predict_f(tt(new), full_output_cov=True)[1]
    assert tuple(full.shape) == (len(new), 1, 1) and rel(npy(full)[:, 0, 0], want_var) < 1e-9
    apost = mf.AnalyticPosteriorProcess(kf.posterior_state_space_model(), kern, tt(tp), GaussianLikelihood(noise ** 2))
    y_mean, y_var = apost.predict_y(tt(new))
    assert rel(npy(y_var)[:, 0], want_var + noise ** 2) < 1e-9


def test_conditional_predict_reference_signature_batched_and_sum_kernel():
    """conditional_predict(new_time_points, training_time_points, kernel, pairwise means[, covariances])
    (conditionals.py:29-83) with a batch of series and a Sum kernel (Matern32 + HarmonicOscillator, D = 4)."""
    import markovflow_b200 as mf

    kern = mf.Sum([mf.Matern32(0.8, 1.1), mf.HarmonicOscillator(0.6, 2.5)], jitter=1e-6)
    rng = np.random.default_rng(3)
    tp = np.cumsum(rng.uniform(0.08, 0.25, size=(3, 30)), axis=-1)
    y = rng.standard_normal((3, 30, 1))
    noise = 0.3
    prior = kern.state_space_model(tt(tp))
    assert prior.state_dim == 4
    kf = mf.KalmanFilter(prior, kern.generate_emission_model(tt(tp)), tt(y), tt([[noise]]))
    post = kf.posterior_state_space_model()
    new = np.sort(rng.uniform(-1.0, 6.0, size=(3, 17)), axis=-1)
    pw_mu, pw_cov = mf.pairwise_marginals(post, kern.initial_mean((3,)), kern.initial_covariance(tt(new[..., :1])))
    mean, cov = mf.conditional_predict(tt(new), tt(tp), kern, pw_mu, pw_cov)
    assert tuple(mean.shape) == (3, 17, 4) and tuple(cov.shape) == (3, 17, 4, 4)
    for b in range(3):
        r_tt = np.abs(tp[b][:, None] - tp[b][None, :])
        r_st = np.abs(new[b][:, None] - tp[b][None, :])
        lam = math.sqrt(3.0) / 0.8
        kf_ = lambda r: 1.1 * (1 + lam * r) * np.exp(-lam * r) + 0.6 * np.cos(2 * math.pi * r / 2.5)  # noqa: E731
        kxx = kf_(r_tt) + noise ** 2 * np.eye(30)
        ksx = kf_(r_st)
        want_mean = ksx @ np.linalg.solve(kxx, y[b, :, 0])
        want_var = kf_(np.zeros(17)) - np.einsum("ij,ji->i", ksx, np.linalg.solve(kxx, ksx.T))
        f_mean = npy(mean)[b][:, 0] + npy(mean)[b][:, 2]   # H = [1, 0, 1, 0]
        c = npy(cov)[b]
        f_var = c[:, 0, 0] + c[:, 2, 2] + 2 * c[:, 0, 2]
        # compare between the data points: beyond them the reference's scheme conditions on a phantom state
        # "at infinity" carrying the prior, which a (nearly) deterministic oscillator never reverts to.  The
        # kernel's jitter (1e-6 on every Q_k and on P0) is part of the model the operators see.
        inside = (new[b] > tp[b][0]) & (new[b] < tp[b][-1])
        assert inside.sum() >= 8
        assert rel(f_mean[inside], want_mean[inside]) < 1e-4 and rel(f_var[inside], want_var[inside]) < 1e-3
    # without pairwise covariances: the conditional density
    mean2, cov2 = mf.conditional_predict(tt(new), tt(tp), kern, pw_mu)
    proj, tcov = mf.conditional_statistics(tt(new), tt(tp), kern)
    assert rel(npy(mean2), npy(mean)) < 1e-12 and rel(npy(cov2), npy(tcov)) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,b,t", [(1, 3, 50), (2, 2, 700), (3, 1, 4000), (4, 5, 33), (6, 4, 20)])
def test_in_kernel_sampling_reproduces_its_own_stream(d, b, t, dtype):
    """sample(shape, seed) draws the standard normals inside the sweep (mf_ssm_sample, Philox keyed by seed /
    trajectory / step): equal to sample_from_epsilons on the stream mf_philox_normal writes out -- for the
    sequential sweep, the parallel-in-time path of few long chains and the direct kernels (D > 4) alike."""
    import markovflow_b200 as mf
    from markovflow_b200 import _lib
    from tests.helpers import random_ssm_arrays

    np.random.seed(d * 100 + t)
    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
    ssm = mf.StateSpaceModel(*(tt(a).to(dtype) for a in arrays))
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    for shape in ((), (3,), (2, 2)):
        got = ssm.sample(shape, seed=1234)
        eps = ssm.sample_epsilons(shape, seed=1234)
        assert tuple(got.shape) == shape + (b, t, d) and got.dtype == dtype
        want = ssm.sample_from_epsilons(eps)
        assert rel(npy(got), npy(want)) < tol
        assert torch.equal(got, ssm.sample(shape, seed=1234))          # reproducible
        assert not torch.equal(got, ssm.sample(shape, seed=1235))      # and seed dependent
    lib = _lib.lib()
    lib.mf_set_tuning(2, 1)  # sequential sweep only
    try:
        seq = ssm.sample((3,), seed=77)
    finally:
        lib.mf_set_tuning(2, 0)
    assert rel(npy(seq), npy(ssm.sample((3,), seed=77))) < tol
    e = npy(ssm.sample_epsilons((64,), seed=5))
    assert abs(e.mean()) < 5.0 / math.sqrt(e.size) and abs(e.var() - 1.0) < 8.0 / math.sqrt(e.size)
    torch.manual_seed(11)
    a1 = ssm.sample((2,))
    torch.manual_seed(11)
    assert torch.equal(a1, ssm.sample((2,)))  # default seed comes from torch's global generator


def test_posterior_samples_have_the_predicted_moments():
    """ConditionalProcess.sample_state_trajectories / sample_f (posterior.py:260-410): the sample mean and
    variance at new points approach predict_f."""
    import markovflow_b200 as mf

    kern, tp, y, kf = _gpr(mf.Matern32, 2, 0.9, 1.4, 0.2, n=25, seed=4)
    post = mf.ConditionalProcess(kf.posterior_state_space_model(), kern, tt(tp))
    new = np.sort(np.random.default_rng(2).uniform(-1.0, tp[-1] + 1.0, size=12))
    n = 40000
    s_new, s_z = post.sample_state_trajectories(tt(new), (n,), seed=3)
    assert tuple(s_new.shape) == (n, 12, 2) and tuple(s_z.shape) == (n, 25, 2)
    f = post.sample_f(tt(new), (n,), seed=3)
    f_mean, f_var = post.predict_f(tt(new))
    se = npy(torch.sqrt(f_var / n))[:, 0]
    assert np.all(np.abs(npy(f.mean(0))[:, 0] - npy(f_mean)[:, 0]) < 5 * se + 1e-12)
    assert rel(npy(f.var(0))[:, 0], npy(f_var)[:, 0]) < 0.05
    zm, zc = post.gauss_markov_model.marginals
    assert np.all(np.abs(npy(s_z.mean(0)) - npy(zm)) < 5 * npy(torch.sqrt(torch.diagonal(zc, dim1=-2, dim2=-1) / n)) + 1e-12)


def test_importance_weights_are_constant_for_the_exact_posterior():
    """ImportanceWeightedPosteriorProcess (posterior.py:470-700): w = p(Y|s) p(u) / q(u).  With the exact
    Gaussian posterior as proposal and the data points as conditioning points every weight equals the marginal
    likelihood p(Y) -- the Kalman log-likelihood -- whatever sample is drawn."""
    import markovflow_b200 as mf

    noise = 0.25
    kern, tp, y, kf = _gpr(mf.Matern52, 3, 1.1, 0.8, noise, n=30, seed=6)
    lik = GaussianLikelihood(noise ** 2)
    iw = mf.ImportanceWeightedPosteriorProcess(8, kf.posterior_state_space_model(), kern, tt(tp), lik)
    data = (tt(tp), tt(y[:, None]))
    new = np.sort(np.random.default_rng(8).uniform(0.0, tp[-1], size=9))
    s_new, log_w, u = iw._iwvi_samples_and_weights(tt(new), data, (5, 8), seed=2)
    assert tuple(s_new.shape) == (5, 8, 9, 3) and tuple(log_w.shape) == (5, 8) and tuple(u.shape) == (5, 8, 30, 3)
    want = float(kf.log_likelihood())
    assert rel(npy(log_w), np.full((5, 8), want)) < 1e-8
    samples, cond = iw.sample_state_trajectories(tt(new), (6,), input_data=data, seed=1)
    assert tuple(samples.shape) == (6, 9, 3) and tuple(cond.shape) == (6, 8, 30, 3)
    f = iw.sample_f(tt(new), 4, input_data=data, seed=1)
    assert tuple(f.shape) == (4, 9, 1)
    ev = iw.expected_value(tt(new), data, seed=3)
    post = mf.ConditionalProcess(kf.posterior_state_space_model(), kern, tt(tp))
    f_mean, f_var = post.predict_f(tt(new))
    assert np.all(np.abs(npy(ev)[:, 0] - npy(f_mean)[:, 0]) < 6 * npy(torch.sqrt(f_var / 8))[:, 0] + 1e-9)
    with pytest.raises(ValueError):
        iw.sample_state_trajectories(tt(new), 2)
