"""GPU parity of ``markovflow_b200.StateSpaceModel`` and the Kalman filters (CUDA kernels through
the C ABI) against the numpy oracle, the golden vectors produced by the reference's own
``NumpyKalmanFilter`` (``tests/golden``) and dense Gaussians.  Mirrors the reference's
``tests/unit/test_state_space_model.py``, ``tests/unit/test_sampling_from_ssm.py``,
``tests/integration/test_kalman_filter.py`` and ``test_kalman_filter_with_sites.py``.
float64 tolerance 1e-10 (max-abs relative), float32 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import assert_parity, dense_ssm_mean_cov, ld, max_rel_err, random_ssm_arrays
from tests.test_oracle_ssm_kalman import GOLDEN, load_kalman_case

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def dev():
    return torch.device("cuda:0")


def tt(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), device=dev()).to(dtype)


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def make_ssm(arrays, dtype=torch.float64):
    from markovflow_b200 import StateSpaceModel

    return StateSpaceModel(*(tt(a, dtype) for a in arrays))


def to_gpu_ssm(ssm: O.SSM, dtype=torch.float64):
    return make_ssm((ssm.mu0, ssm.chol_p0, ssm.a_s, ssm.b_s, ssm.chol_q_s), dtype)


# ------------------------------------------------------------------------------------------------
# StateSpaceModel
# ------------------------------------------------------------------------------------------------

def test_rejects_bad_shapes():
    from markovflow_b200 import StateSpaceModel

    mu0, l0, a, b, lq = (tt(x) for x in random_ssm_arrays((2,), 3, 2))
    with pytest.raises(ValueError):
        StateSpaceModel(mu0, l0, a[:, :0], b[:, :0], lq[:, :0])  # num_transitions = 0
    with pytest.raises(ValueError):
        StateSpaceModel(mu0[0], l0, a, b, lq)
    with pytest.raises(ValueError):
        StateSpaceModel(mu0, l0, a, b[:, :2], lq)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_precision_means_covariances_logdet(batch_shape, state_dim, transitions, dtype):
    arrays = random_ssm_arrays(batch_shape, transitions, state_dim)
    if dtype == torch.float32:  # identical inputs on both sides: the float32-representable values
        arrays = tuple(a.astype(np.float32).astype(np.float64) for a in arrays)
    ref = O.SSM(*arrays)
    ssm = make_ssm(arrays, dtype)
    tol = TOL[dtype]
    # extended-precision / same-arithmetic evaluations of the restated reference, used only when a
    # result is further than `tol` from the float64 oracle (tests/helpers.py::assert_parity)
    adj = (dict(truth=lambda: O.ssm_marginal_covariances(O.SSM(*ld(*arrays)))) if dtype == torch.float64 else
           dict(peer=lambda: O.ssm_marginal_covariances(O.SSM(*(a.astype(np.float32) for a in arrays)))))
    prec = ssm.precision
    o_d, o_s = O.ssm_build_precision(ref)
    assert max_rel_err(npy(prec.block_diagonal), o_d) < tol
    assert max_rel_err(npy(prec.block_sub_diagonal), o_s) < tol
    mean, cov = ssm.marginals
    assert max_rel_err(npy(mean), O.ssm_marginal_means(ref)) < tol
    assert max_rel_err(npy(ssm.marginal_means), O.ssm_marginal_means(ref)) < tol
    o_cov = O.ssm_marginal_covariances(ref)
    assert_parity(npy(cov), o_cov, tol, what="marginal_covariances", **adj)
    cov2, sub = ssm.covariance_blocks()
    adj_sub = {k: (lambda f=f: O.ssm_subsequent_covariances(
        O.SSM(*(ld(*arrays) if k == "truth" else tuple(a.astype(np.float32) for a in arrays))), f()))
        for k, f in adj.items()}
    assert_parity(npy(sub), O.ssm_subsequent_covariances(ref, o_cov), tol, what="subsequent_covariances", **adj_sub)
    assert max_rel_err(npy(ssm.subsequent_covariances(cov2)), npy(sub)) < tol
    assert max_rel_err(npy(ssm.log_det_precision()), O.ssm_log_det_precision(ref)) < tol
    assert tuple(ssm.batch_shape) == batch_shape and ssm.state_dim == state_dim
    assert ssm.num_transitions == transitions


def test_covariances_match_dense_propagation(state_dim, transitions):
    """The forward recursion is at least as accurate as the reference route (precision ->
    Cholesky -> inverse subset): compare both with the dense joint covariance."""
    arrays = random_ssm_arrays((), transitions, state_dim)
    _, dense_cov = dense_ssm_mean_cov(*arrays)
    d, t = state_dim, transitions + 1
    want = np.stack([dense_cov[k * d:(k + 1) * d, k * d:(k + 1) * d] for k in range(t)])
    cov = npy(make_ssm(arrays).marginal_covariances)
    assert max_rel_err(cov, want) < 1e-12


@pytest.mark.parametrize("sample_shape", [(), (4,), (2, 3)])
def test_log_pdf(batch_shape, state_dim, transitions, sample_shape):
    arrays = random_ssm_arrays(batch_shape, transitions, state_dim)
    ref = O.SSM(*arrays)
    ssm = make_ssm(arrays)
    states = np.random.normal(size=sample_shape + batch_shape + (transitions + 1, state_dim))
    got = npy(ssm.log_pdf(tt(states)))
    want = O.ssm_log_pdf(ref, states)
    assert got.shape == sample_shape + batch_shape
    assert max_rel_err(got, want) < 1e-10


def test_log_pdf_long_chain_uses_segments():
    arrays = random_ssm_arrays((2,), 9000, 2, scale_a=0.3)
    states = np.random.normal(size=(3, 2, 9001, 2))
    got = npy(make_ssm(arrays).log_pdf(tt(states)))
    assert max_rel_err(got, O.ssm_log_pdf(O.SSM(*arrays), states)) < 1e-10


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_kl_divergence(batch_shape, state_dim, transitions, dtype):
    q_arr = random_ssm_arrays(batch_shape, transitions, state_dim)
    p_arr = random_ssm_arrays(batch_shape, transitions, state_dim)
    if dtype == torch.float32:
        q_arr = tuple(a.astype(np.float32).astype(np.float64) for a in q_arr)
        p_arr = tuple(a.astype(np.float32).astype(np.float64) for a in p_arr)
    want = O.ssm_kl_divergence(O.SSM(*q_arr), O.SSM(*p_arr))
    got = npy(make_ssm(q_arr, dtype).kl_divergence(make_ssm(p_arr, dtype)))
    assert got.shape == batch_shape
    f32 = lambda arr: O.SSM(*(a.astype(np.float32) for a in arr))
    adj = (dict(truth=lambda: O.ssm_kl_divergence(O.SSM(*ld(*q_arr)), O.SSM(*ld(*p_arr)))) if dtype == torch.float64
           else dict(peer=lambda: O.ssm_kl_divergence(f32(q_arr), f32(p_arr))))
    assert_parity(got, want, TOL[dtype], what="kl_divergence", **adj)
    # KL(q || q) = 0: the terms that cancel are O(T D) each, so the bar is tol relative to their size
    same = npy(make_ssm(q_arr, dtype).kl_divergence(make_ssm(q_arr, dtype)))
    scale = 0.5 * (transitions + 1) * state_dim
    assert np.max(np.abs(same)) < TOL[dtype] * scale


@pytest.mark.parametrize("sample_shape", [(5,), (2, 2), ()])
def test_sample_from_epsilons(batch_shape, state_dim, transitions, sample_shape):
    arrays = random_ssm_arrays(batch_shape, transitions, state_dim)
    eps = np.random.normal(size=sample_shape + batch_shape + (transitions + 1, state_dim))
    got = npy(make_ssm(arrays).sample_from_epsilons(tt(eps)))
    want = O.ssm_sample_from_epsilons(O.SSM(*arrays), eps)
    assert max_rel_err(got, want) < 1e-10


def test_sample_shapes_and_moments():
    """tests/unit/test_sampling_from_ssm.py: shapes, and sample moments approach the marginals."""
    arrays = random_ssm_arrays((2,), 4, 2)
    ssm = make_ssm(arrays)
    assert tuple(ssm.sample((3, 2)).shape) == (3, 2, 2, 5, 2)
    assert tuple(ssm.sample(0).shape) == (0, 2, 5, 2)
    gen = torch.Generator(device=dev()).manual_seed(1)
    s = npy(ssm.sample(20000, generator=gen))
    mean, cov = (npy(x) for x in ssm.marginals)
    assert np.max(np.abs(s.mean(0) - mean)) < 0.05 * max(1.0, np.abs(mean).max())
    emp = np.einsum("sbti,sbtj->btij", s - mean, s - mean) / s.shape[0]
    assert np.max(np.abs(emp - cov)) < 0.08 * np.abs(cov).max()


def test_state_space_model_from_covariances_with_zero_blocks():
    from markovflow_b200 import state_space_model_from_covariances

    mu0, l0, a, b, lq = random_ssm_arrays((3,), 4, 3)
    p0 = l0 @ np.swapaxes(l0, -1, -2)
    q = lq @ np.swapaxes(lq, -1, -2)
    q[1, 2] = 0.0
    p0[2] = 0.0
    ssm = state_space_model_from_covariances(tt(mu0), tt(p0), tt(a), tt(b), tt(q))
    ref = O.ssm_from_covariances(mu0, p0, a, b, q)
    assert max_rel_err(npy(ssm.cholesky_process_covariances), ref.chol_q_s) < 1e-12
    assert max_rel_err(npy(ssm.cholesky_initial_covariance), ref.chol_p0) < 1e-12
    assert max_rel_err(npy(ssm.marginal_means), O.ssm_marginal_means(ref)) < 1e-12


def test_normalizer_and_a_inv_block():
    arrays = random_ssm_arrays((2,), 5, 3)
    ref = O.SSM(*arrays)
    ssm = make_ssm(arrays)
    mu = O.ssm_marginal_means(ref)
    pd, ps = O.ssm_build_precision(ref)
    maha = np.sum(mu * O.btd_dense_mult(pd, ps, mu, symmetric=True), axis=(-1, -2))
    want = 0.5 * (6 * 3 * np.log(2 * np.pi) - O.ssm_log_det_precision(ref) + maha)
    assert max_rel_err(npy(ssm.normalizer()), want) < 1e-10
    c = np.random.normal(size=(4, 2, 6, 3))
    got = npy(ssm.a_inv_block.solve(tt(c)))
    assert max_rel_err(got, O.affine_recurrence(ref.a_s, c)) < 1e-10
    got_t = npy(ssm.a_inv_block.solve(tt(c), transpose_left=True))
    eye = np.broadcast_to(np.eye(3), (2, 6, 3, 3))
    assert max_rel_err(got_t, O.btd_solve(eye, -ref.a_s, c, transpose_left=True)) < 1e-10


# ------------------------------------------------------------------------------------------------
# Kalman filter: golden vectors of the reference's own numpy filter
# ------------------------------------------------------------------------------------------------

def _filter_from_golden(tag, dtype=torch.float64):
    from markovflow_b200 import EmissionModel, KalmanFilter

    g, ssm, h = load_kalman_case(os.path.join(GOLDEN, f"kalman_{tag}.npz"))
    kf = KalmanFilter(to_gpu_ssm(ssm, dtype), EmissionModel(tt(h, dtype)), tt(g["y"], dtype),
                      tt(g["chol_R"], dtype))
    return g, ssm, h, kf


@pytest.mark.parametrize("tag", ["b3", "b0", "b21", "d5m1"])
def test_log_likelihood_matches_reference_kalman_filter(tag):
    """tests/integration/test_kalman_filter.py:131-139."""
    g, ssm, h, kf = _filter_from_golden(tag)
    assert max_rel_err(npy(kf.log_likelihood()), np.sum(g["log_liks"])) < 1e-10
    per_chain = npy(kf.log_likelihood_per_chain())
    assert max_rel_err(per_chain, np.sum(g["log_liks"], axis=-1)) < 1e-10
    # the SpInGP restatement (kalman_filter.py:184-255) against the same inputs; where it is further
    # than 1e-10 away, the long-double evaluation of the same restatement decides
    want = O.kalman_log_likelihood(ssm, h, g["y"], O._r_inv_from_chol(g["chol_R"]), per_chain=True)
    ssm_ld = O.SSM(*ld(ssm.mu0, ssm.chol_p0, ssm.a_s, ssm.b_s, ssm.chol_q_s))
    assert_parity(per_chain, want, 1e-10, what=f"log_likelihood[{tag}] vs SpInGP restatement",
                  truth=lambda: O.kalman_log_likelihood(ssm_ld, ld(h), ld(g["y"]),
                                                        O._r_inv_from_chol(ld(g["chol_R"])), per_chain=True))


@pytest.mark.parametrize("tag", ["b3", "b0"])
def test_log_likelihood_float32(tag):
    g, _, _, kf = _filter_from_golden(tag, torch.float32)
    assert max_rel_err(npy(kf.log_likelihood()), np.sum(g["log_liks"])) < 1e-4


@pytest.mark.parametrize("tag", ["b3", "b0", "b21", "d5m1"])
def test_posterior_ssm_matches_reference_rts_smoother(tag):
    """tests/integration/test_kalman_filter.py:105-128."""
    g, ssm, h, kf = _filter_from_golden(tag)
    post = kf.posterior_state_space_model()
    mean, cov = post.marginals
    # golden vectors of the reference's float64 numpy RTS smoother; beyond 1e-10 the long-double
    # evaluation of the restated posterior (kalman_filter.py:109-182) decides which side is off
    ssm_ld = O.SSM(*ld(ssm.mu0, ssm.chol_p0, ssm.a_s, ssm.b_s, ssm.chol_q_s))
    post_ld = lambda: O.kalman_posterior_ssm(ssm_ld, ld(h), ld(g["y"]), O._r_inv_from_chol(ld(g["chol_R"])))
    assert_parity(npy(mean), g["smooth_means"], 1e-10, what=f"posterior means[{tag}]",
                  truth=lambda: O.ssm_marginal_means(post_ld()))
    assert_parity(npy(cov), np.broadcast_to(g["smooth_covs"], npy(cov).shape), 1e-10,
                  what=f"posterior covariances[{tag}]", truth=lambda: O.ssm_marginal_covariances(post_ld()))
    ref_post = O.kalman_posterior_ssm(ssm, h, g["y"], O._r_inv_from_chol(g["chol_R"]))
    assert max_rel_err(npy(post.state_transitions), ref_post.a_s) < 1e-10
    assert max_rel_err(npy(post.state_offsets), ref_post.b_s) < 1e-10
    assert max_rel_err(npy(post.cholesky_process_covariances), ref_post.chol_q_s) < 1e-10
    assert max_rel_err(npy(post.initial_mean), ref_post.mu0) < 1e-10


def test_sites_log_likelihood_and_posterior_match_reference():
    """tests/integration/test_kalman_filter_with_sites.py."""
    from markovflow_b200 import EmissionModel, KalmanFilterWithSites, UnivariateGaussianSitesNat

    g = np.load(os.path.join(GOLDEN, "kalman_sites_t10.npz"))
    d, t = g["A"].shape[0], g["nat1"].shape[0]
    ssm = O.SSM(g["mu0"], g["chol_P0"], np.broadcast_to(g["A"], (t - 1, d, d)).copy(),
                np.broadcast_to(g["b"], (t - 1, d)).copy(), np.broadcast_to(g["chol_Q"], (t - 1, d, d)).copy())
    h = np.broadcast_to(g["H"], (t, 1, d)).copy()
    sites = UnivariateGaussianSitesNat(tt(g["nat1"]), tt(g["nat2"]))
    kf = KalmanFilterWithSites(to_gpu_ssm(ssm), EmissionModel(tt(h)), sites)
    assert max_rel_err(npy(kf.log_likelihood()), np.sum(g["log_liks"])) < 1e-10
    mean, cov = kf.posterior_state_space_model().marginals
    _, prec_k, _ = O.sites_means_precisions(g["nat1"], g["nat2"])
    post_ld = lambda: O.kalman_posterior_ssm(O.SSM(*ld(ssm.mu0, ssm.chol_p0, ssm.a_s, ssm.b_s, ssm.chol_q_s)),
                                             ld(h), ld(npy(sites.means)), ld(prec_k))
    assert_parity(npy(mean), g["smooth_means"], 1e-10, what="sites posterior means",
                  truth=lambda: O.ssm_marginal_means(post_ld()))
    assert_parity(npy(cov), g["smooth_covs"], 1e-10, what="sites posterior covariances",
                  truth=lambda: O.ssm_marginal_covariances(post_ld()))


def test_sparse_sites_equal_dense_filter_on_the_data_points():
    """KalmanFilterWithSparseSites (kalman_filter.py:500-626): grid points without data carry no
    observation, so the value equals the oracle filter that skips those steps."""
    from markovflow_b200 import EmissionModel, KalmanFilterWithSparseSites, UnivariateGaussianSitesNat

    rng = np.random.default_rng(11)
    k = O.Matern32(0.8, 1.2)
    t = 30
    tp = np.cumsum(rng.uniform(0.05, 0.3, size=t))
    ssm, h = k.state_space_model(tp), k.emission_matrix(tp)
    idx = np.sort(rng.choice(t, size=12, replace=False))
    prec = rng.uniform(0.5, 2.0, size=(12, 1, 1))
    y = rng.standard_normal((12, 1))
    sites = UnivariateGaussianSitesNat(tt(prec[..., 0] * y), tt(-0.5 * prec))
    kf = KalmanFilterWithSparseSites(to_gpu_ssm(ssm), EmissionModel(tt(h)), sites, t,
                                     tt(idx[:, None]).long(), tt(y))
    # oracle: time-varying filter with a huge variance at the grid points without data
    r = np.full((t, 1, 1), 1e30)
    r[idx] = 1.0 / prec
    obs = np.zeros((t, 1))
    obs[idx] = y
    q = ssm.chol_q_s @ np.swapaxes(ssm.chol_q_s, -1, -2)
    lls, _, _ = O.kalman_filter_time_varying(ssm.mu0, ssm.chol_p0 @ ssm.chol_p0.T, ssm.a_s, ssm.b_s,
                                             q, h, r, obs)
    assert max_rel_err(npy(kf.log_likelihood()), np.sum(lls[idx])) < 1e-10


@pytest.mark.parametrize("m", [2, 3])
def test_multi_output_time_varying_batched(m):
    """output_dim > 1 with a full observation covariance and per-chain emission matrices."""
    rng = np.random.default_rng(m)
    b, n, d = 5, 17, 3
    arrays = random_ssm_arrays((b,), n, d)
    ref = O.SSM(*arrays)
    h = rng.standard_normal((b, n + 1, m, d))
    y = rng.standard_normal((b, n + 1, m))
    lr = np.tril(rng.standard_normal((m, m))) * 0.3 + np.eye(m)
    want = O.kalman_log_likelihood(ref, h, y, O._r_inv_from_chol(lr), per_chain=True)
    from markovflow_b200 import kalman_log_likelihood

    got = npy(kalman_log_likelihood(make_ssm(arrays), tt(h), tt(y), tt(lr)))
    assert max_rel_err(got, want) < 1e-10


# ------------------------------------------------------------------------------------------------
# parallel-in-time path (few long chains)
# ------------------------------------------------------------------------------------------------

def _long_case(kernel, t, rng, b=1):
    tps = [np.cumsum(rng.uniform(0.05, 0.15, size=t)) for _ in range(b)]
    ssms = [kernel.state_space_model(tp) for tp in tps]
    stack = lambda name: np.stack([getattr(s, name) for s in ssms])
    ssm = O.SSM(stack("mu0"), stack("chol_p0"), stack("a_s"), stack("b_s"), stack("chol_q_s"))
    h = np.stack([kernel.emission_matrix(tp) for tp in tps])
    y = np.sin(np.stack(tps))[..., None] + 0.1 * rng.standard_normal((b, t, 1))
    return ssm, h, y


@pytest.mark.parametrize("name,t,b", [("m32", 5000, 1), ("m52", 3001, 2), ("m32", 200000, 1), ("sum", 2500, 1)])
def test_parallel_in_time_equals_sequential_and_oracle(name, t, b):
    from markovflow_b200 import _lib, kalman_log_likelihood

    rng = np.random.default_rng(t)
    kern = {"m32": O.Matern32(1.0, 1.0), "m52": O.Matern52(0.9, 1.3),
            "sum": O.Sum([O.Matern32(1.0, 1.0), O.HarmonicOscillator(0.5, 1.0)], jitter=1e-6)}[name]
    ssm, h, y = _long_case(kern, t, rng, b)
    lr = np.array([[0.1]])
    gssm = to_gpu_ssm(ssm)
    lib = _lib.lib()
    try:
        lib.mf_set_tuning(2, 1)
        seq = npy(kalman_log_likelihood(gssm, tt(h), tt(y), tt(lr)))
        lib.mf_set_tuning(2, 2)
        par = npy(kalman_log_likelihood(gssm, tt(h), tt(y), tt(lr)))
        lib.mf_set_tuning(3, 37)  # odd segment length, ragged tail
        par2 = npy(kalman_log_likelihood(gssm, tt(h), tt(y), tt(lr)))
    finally:
        lib.mf_set_tuning(2, 0)
        lib.mf_set_tuning(3, 0)
    assert max_rel_err(par, seq) < 1e-10
    assert max_rel_err(par2, seq) < 1e-10
    if t <= 5000:
        want = O.kalman_log_likelihood(ssm, h, y, O._r_inv_from_chol(lr), per_chain=True)
        assert max_rel_err(seq, want) < 1e-10
        assert max_rel_err(par, want) < 1e-10


def test_time_sharded_segments_reproduce_the_whole_series():
    """The multi-GPU protocol on one device: segment summaries -> fold -> seeded local filters."""
    from markovflow_b200.parallel import time_sharded_log_likelihood_local

    rng = np.random.default_rng(4)
    ssm, h, y = _long_case(O.Matern32(1.0, 1.0), 10007, rng)
    lr = np.array([[0.1]])
    from markovflow_b200 import kalman_log_likelihood

    whole = npy(kalman_log_likelihood(to_gpu_ssm(ssm), tt(h), tt(y), tt(lr)))
    for world in (2, 3, 8):
        got = time_sharded_log_likelihood_local(to_gpu_ssm(ssm), tt(h), tt(y), tt(lr), world)
        assert max_rel_err(npy(got), whole) < 1e-10
        got = time_sharded_log_likelihood_local(to_gpu_ssm(ssm), tt(h), tt(y), tt(lr), world, seeded=True)
        assert max_rel_err(npy(got), whole) < 1e-10


def test_segment_element_matches_oracle_element():
    """mf_kalman_segment_summary against the oracle's element with log-normaliser."""
    from markovflow_b200.parallel import CudaKalmanEngine, time_sharded_segments

    rng = np.random.default_rng(9)
    ssm, h, y, lr = (*_long_case(O.Matern52(0.9, 1.3), 301, rng), np.array([[0.2]]))
    segs = time_sharded_segments(to_gpu_ssm(ssm), tt(h), tt(y), tt(lr), 3)
    q = ssm.chol_q_s @ np.swapaxes(ssm.chol_q_s, -1, -2)
    lo = 0
    for r, seg in enumerate(segs):
        got = npy(CudaKalmanEngine().segment_summary(seg))[0]
        n = seg.num_steps
        tlo = 0 if r == 0 else lo - 1
        prior = (ssm.mu0[0], ssm.chol_p0[0] @ ssm.chol_p0[0].T) if r == 0 else None
        want = O.pscan_segment_element(ssm.a_s[0, tlo:lo + n - 1], ssm.b_s[0, tlo:lo + n - 1],
                                       q[0, tlo:lo + n - 1], h[0, lo:lo + n], lr @ lr.T,
                                       y[0, lo:lo + n], prior=prior)
        want = np.concatenate([np.reshape(x, -1) for x in want])
        assert max_rel_err(got, want) < 1e-10
        lo += n


# ------------------------------------------------------------------------------------------------
# few long chains: the moment recursion is evaluated parallel in time (segment elements -> seeds ->
# seeded sweeps, ssm_sweep.cuh); results must equal the sequential sweep and the direct propagation
# ------------------------------------------------------------------------------------------------

def _propagate_moments(mu0, chol_p0, a_s, b_s, chol_q):
    """Direct numpy propagation of (mu_k, Sigma_kk, A_k Sigma_kk) for batched arrays."""
    t = a_s.shape[-3] + 1
    mu = [mu0]
    p = [chol_p0 @ np.swapaxes(chol_p0, -1, -2)]
    sub = []
    for k in range(t - 1):
        a = a_s[..., k, :, :]
        sub.append(a @ p[-1])
        mu.append(np.einsum("...ij,...j->...i", a, mu[-1]) + b_s[..., k, :])
        lq = chol_q[..., k, :, :]
        p.append(a @ p[-1] @ np.swapaxes(a, -1, -2) + lq @ np.swapaxes(lq, -1, -2))
    return np.stack(mu, -2), np.stack(p, -3), np.stack(sub, -3)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
@pytest.mark.parametrize("b,t", [(1, 1000), (3, 301), (2, 129), (5, 640)])
def test_marginals_parallel_in_time(b, t, d, dtype):
    import markovflow_b200 as mf
    from markovflow_b200 import _lib

    rng = np.random.RandomState(b * 1000 + t + d)
    state = np.random.get_state()
    np.random.seed(rng.randint(1 << 30))
    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
    np.random.set_state(state)
    ssm = make_ssm(arrays, dtype)
    want_mu, want_p, want_sub = _propagate_moments(*arrays)
    tol = TOL[dtype]
    lib = _lib.lib()
    results = {}
    for knob in (0, 1):  # 0: parallel in time when it pays, 1: sequential sweep
        lib.mf_set_tuning(2, knob)
        try:
            mean, cov = ssm.marginals
            cov2, sub = ssm.covariance_blocks()
            eta = mf.ssm_to_expectations(ssm)
        finally:
            lib.mf_set_tuning(2, 0)
        results[knob] = (npy(mean), npy(cov), npy(sub), [npy(e) for e in eta])
        assert max_rel_err(npy(mean), want_mu) < tol
        assert max_rel_err(npy(cov), want_p) < tol
        assert max_rel_err(npy(cov2), want_p) < tol
        assert max_rel_err(npy(sub), want_sub) < tol
        assert max_rel_err(npy(eta[0]), want_mu) < tol
        assert max_rel_err(npy(eta[1]), want_p + want_mu[..., :, None] * want_mu[..., None, :]) < tol
        assert max_rel_err(npy(eta[2]), want_sub + want_mu[..., 1:, :, None] * want_mu[..., :-1, None, :]) < tol
    for x, y in zip(results[0][:3], results[1][:3]):
        assert max_rel_err(x, y) < tol


def test_marginals_parallel_in_time_segment_length_knob():
    """Forced short segments (knob 3) with a ragged last segment of one step."""
    from markovflow_b200 import _lib

    lib = _lib.lib()
    # more than 64 segments per chain: the seeds come from the warp-scan kernel
    for t, d, segs in ((101, 2, (2, 3, 10, 50, 100)), (1001, 3, (4, 7, 13)), (2500, 1, (2, 39))):
        arrays = random_ssm_arrays((2,), t - 1, d)
        want_mu, want_p, want_sub = _propagate_moments(*arrays)
        for seg in segs:
            lib.mf_set_tuning(3, seg)
            try:
                ssm = make_ssm(arrays)
                mean, cov = ssm.marginals
                _, sub = ssm.covariance_blocks()
            finally:
                lib.mf_set_tuning(3, 0)
            assert max_rel_err(npy(mean), want_mu) < 1e-10 and max_rel_err(npy(cov), want_p) < 1e-10
            assert max_rel_err(npy(sub), want_sub) < 1e-10


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_means_and_samples_parallel_in_time(d, dtype):
    """Few long chains: x_k = A_k x_{k-1} + b_k (+ chol_q eps) is evaluated parallel in time
    (marginal_means, sample_from_epsilons with leading sample dimensions); equal to the oracle and to
    the sequential sweep."""
    from markovflow_b200 import _lib

    lib = _lib.lib()
    # (3, 700, 7): 100 segments per chain -> warp-scan fold of the elements
    for b, t, seg in ((1, 1500, 0), (2, 301, 0), (3, 400, 9), (3, 700, 7), (2, 131, 65)):
        state = np.random.get_state()
        np.random.seed(t * 10 + d)
        arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
        eps = np.random.normal(size=(2, b, t, d))
        np.random.set_state(state)
        if dtype == torch.float32:
            arrays = tuple(a.astype(np.float32).astype(np.float64) for a in arrays)
            eps = eps.astype(np.float32).astype(np.float64)
        ref = O.SSM(*arrays)
        want_mean = O.ssm_marginal_means(ref)
        want_s = O.ssm_sample_from_epsilons(ref, eps)
        got = {}
        for knob in (0, 1):
            lib.mf_set_tuning(2, knob)
            lib.mf_set_tuning(3, seg)
            try:
                ssm = make_ssm(arrays, dtype)
                got[knob] = (npy(ssm.marginal_means), npy(ssm.sample_from_epsilons(tt(eps, dtype))))
            finally:
                lib.mf_set_tuning(2, 0)
                lib.mf_set_tuning(3, 0)
            assert max_rel_err(got[knob][0], want_mean) < TOL[dtype]
            assert max_rel_err(got[knob][1], want_s) < TOL[dtype]
        assert max_rel_err(got[0][1], got[1][1]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_kl_divergence_parallel_in_time(d, dtype):
    """Few long chains: KL(q || p) with q's marginals seeded per segment (moment elements in a
    stream-ordered workspace) and the per-segment shares added in order; equal to the oracle and to
    the sequential sweep."""
    from markovflow_b200 import _lib

    lib = _lib.lib()
    for b, t, seg in ((1, 1500, 0), (2, 301, 0), (3, 400, 9), (2, 131, 65), (1, 3000, 4)):
        state = np.random.get_state()
        np.random.seed(t * 10 + d)
        qa = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
        pa = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
        np.random.set_state(state)
        if dtype == torch.float32:
            qa = tuple(a.astype(np.float32).astype(np.float64) for a in qa)
            pa = tuple(a.astype(np.float32).astype(np.float64) for a in pa)
        want = O.ssm_kl_divergence(O.SSM(*qa), O.SSM(*pa))
        f32 = lambda arr: O.SSM(*(a.astype(np.float32) for a in arr))
        adj = (dict(truth=lambda: O.ssm_kl_divergence(O.SSM(*ld(*qa)), O.SSM(*ld(*pa)))) if dtype == torch.float64
               else dict(peer=lambda: O.ssm_kl_divergence(f32(qa), f32(pa))))
        got = {}
        for knob in (0, 1):
            lib.mf_set_tuning(2, knob)
            lib.mf_set_tuning(3, seg)
            try:
                got[knob] = npy(make_ssm(qa, dtype).kl_divergence(make_ssm(pa, dtype)))
            finally:
                lib.mf_set_tuning(2, 0)
                lib.mf_set_tuning(3, 0)
            assert_parity(got[knob], want, TOL[dtype], what=f"kl_divergence parallel-in-time knob {knob}", **adj)
        if dtype == torch.float64:  # (float32: each path is held to the oracle above)
            assert_parity(got[0], got[1], 1e-10, what="kl_divergence parallel vs sequential", truth=adj["truth"])


def test_cuda_graph_replay_of_the_single_series_job():
    """``markovflow_b200.Graphed``: log-likelihood + posterior SSM + posterior marginals of one series
    recorded into a CUDA graph; replays track in-place updates of the inputs."""
    import markovflow_b200 as mf

    g, ssm_ref, h, _ = _filter_from_golden("b0")
    y = tt(g["y"])
    kf = mf.KalmanFilter(to_gpu_ssm(ssm_ref), mf.EmissionModel(tt(h)), y, tt(g["chol_R"]))

    def job():
        post = kf.posterior_state_space_model()
        return kf.log_likelihood(), post.marginals

    graphed = mf.Graphed(job)
    ll_e, (m_e, c_e) = job()
    ll_g, (m_g, c_g) = graphed()
    assert max_rel_err(npy(ll_g), npy(ll_e)) < 1e-12
    assert max_rel_err(npy(m_g), npy(m_e)) < 1e-12 and max_rel_err(npy(c_g), npy(c_e)) < 1e-12
    y.mul_(1.5)  # new observations, same buffers
    ll_e2, (m_e2, _) = job()
    ll_g2, (m_g2, _) = graphed()
    assert max_rel_err(npy(ll_g2), npy(ll_e2)) < 1e-12 and max_rel_err(npy(m_g2), npy(m_e2)) < 1e-12
    assert abs(float(ll_e2) - float(ll_e)) > 1e-6


def test_mid_size_batch_log_likelihood_is_cut_in_time():
    """1000 series leave most SMs idle as one chain each: they are cut into 32 segments per series
    (scan elements -> warp joins -> ordered reduction); same value as the uncut filter and the oracle."""
    from markovflow_b200 import _lib, kalman_log_likelihood

    rng = np.random.default_rng(77)
    b, t = 1000, 2100
    ssm, h, y = _long_case(O.Matern32(1.0, 1.0), t, rng, b)
    lr = np.array([[0.1]])
    gssm = to_gpu_ssm(ssm)
    lib = _lib.lib()
    cut = npy(kalman_log_likelihood(gssm, tt(h), tt(y), tt(lr)))
    try:
        lib.mf_set_tuning(2, 1)
        uncut = npy(kalman_log_likelihood(gssm, tt(h), tt(y), tt(lr)))
    finally:
        lib.mf_set_tuning(2, 0)
    assert cut.shape == (b,)
    assert max_rel_err(cut, uncut) < 1e-10
    pick = [0, 499, 999]
    sub = O.SSM(ssm.mu0[pick], ssm.chol_p0[pick], ssm.a_s[pick], ssm.b_s[pick], ssm.chol_q_s[pick])
    want = O.kalman_log_likelihood(sub, h[pick], y[pick], O._r_inv_from_chol(lr), per_chain=True)
    assert max_rel_err(cut[pick], want) < 1e-10


@pytest.mark.parametrize("world", [2, 3, 8])
def test_time_sharded_exchange_inside_the_kernel_virtual_ranks(world):
    """mf_kalman_time_sharded_log_likelihood / mf_kalman_matern_time_sharded_log_likelihood: the per-rank
    elements are exchanged through peer-mapped regions and joined INSIDE the reduction kernel.  Here the `world`
    ranks are virtual: one device, one stream per rank, regions = plain device buffers (on several GPUs they are
    cudaIpc mappings of each other's allocations); every rank must return the log-likelihood of the whole series,
    call after call (the regions are double-buffered on the call parity)."""
    import markovflow_b200 as mf
    from markovflow_b200.parallel import (PeerRing, matern_time_segment, time_sharded_log_likelihood,
                                          time_sharded_segments)

    rng = np.random.default_rng(world)
    t, bsz = 40_000, 2
    k = O.Matern32(1.0, 1.0)
    tps = np.cumsum(rng.uniform(0.05, 0.15, size=(bsz, t)), axis=-1)
    ssm = k.state_space_model(tps)
    h = k.emission_matrix(tps[0])
    y = np.sin(tps)[..., None] + 0.1 * rng.standard_normal((bsz, t, 1))
    lr = np.array([[0.1]])
    gssm = to_gpu_ssm(ssm)
    whole = mf.kalman_log_likelihood(gssm, tt(h), tt(y), tt(lr))
    segs = time_sharded_segments(gssm, tt(h), tt(y), tt(lr), world)
    nbytes = PeerRing.region_bytes(torch.float64, bsz, 2, world)
    regions = [torch.zeros(nbytes, dtype=torch.uint8, device=dev()) for _ in range(world)]
    rings = [PeerRing(torch.float64, bsz, 2, regions=[r.data_ptr() for r in regions], rank=r, world=world)
             for r in range(world)]
    streams = [torch.cuda.Stream(device=dev()) for _ in range(world)]
    torch.cuda.synchronize()
    for call in range(3):
        outs = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                outs.append(time_sharded_log_likelihood(segs[r], ring=rings[r]))
        torch.cuda.synchronize()
        for o in outs:
            assert max_rel_err(npy(o), npy(whole)) < 1e-10
    # Matern prior built in the kernel from the time deltas, same series
    dts = tt(np.diff(tps, axis=-1))
    y2 = tt(y[..., 0])
    one = torch.ones(bsz, dtype=torch.float64, device=dev())
    msegs = [matern_time_segment(dts, y2, r, world) for r in range(world)]
    glr = tt(lr)
    torch.cuda.synchronize()  # the slices were cut on the default stream: the ranks' streams must see them
    for call in range(2):
        outs = []
        for r in range(world):
            first, seg_dt, seg_y = msegs[r]
            with torch.cuda.stream(streams[r]):
                outs.append(rings[r].matern(2, one, one, seg_dt, seg_y, glr, first)[0])
        torch.cuda.synchronize()
        for o in outs:
            assert max_rel_err(npy(o), npy(whole)) < 1e-10
