"""GPU parity of ``markovflow_b200.conditionals`` (CUDA kernels through the C ABI) against the numpy
oracle; mirrors ``tests/integration/test_posterior.py:196-246`` of the reference.
float64 tolerance 1e-10 (max-abs relative), float32 1e-4."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import assert_parity, ld, max_rel_err, random_ssm_arrays

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def dev():
    return torch.device("cuda:0")


def tt(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), device=dev()).to(dtype)


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def _spd(rng, shape, d):
    a = rng.standard_normal(shape + (d, d))
    return a @ np.swapaxes(a, -1, -2) + d * np.eye(d)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 5])
@pytest.mark.parametrize("batch_shape,t", [((), 6), ((3,), 2), ((2, 2), 33), ((2,), 400)])
def test_pairwise_marginals(batch_shape, t, d, dtype):
    import markovflow_b200 as mf

    np.random.seed(100 * d + t)
    arrays = random_ssm_arrays(batch_shape, t - 1, d, scale_a=0.6 / np.sqrt(d))
    if dtype == torch.float32:
        arrays = tuple(a.astype(np.float32).astype(np.float64) for a in arrays)
    ssm = mf.StateSpaceModel(*(tt(a, dtype) for a in arrays))
    rng = np.random.default_rng(d)
    for im, ic in ((np.zeros(d), np.zeros((d, d))),
                   (rng.standard_normal(batch_shape + (d,)), _spd(rng, batch_shape, d))):
        jm, jc = mf.pairwise_marginals(ssm, tt(im, dtype), tt(ic, dtype))
        want_m, want_c = O.pairwise_marginals(O.SSM(*arrays), im, ic)
        assert tuple(jm.shape) == batch_shape + (t + 1, 2 * d)
        assert tuple(jc.shape) == batch_shape + (t + 1, 2 * d, 2 * d)
        assert max_rel_err(npy(jm), want_m) < TOL[dtype]
        # the oracle takes the covariances through the precision route (state_space_model.py:253-262);
        # past the tolerance the long-double / float32 evaluation of that route decides
        adj = (dict(truth=lambda: O.pairwise_marginals(O.SSM(*ld(*arrays)), ld(im), ld(ic))[1])
               if dtype == torch.float64 else
               dict(peer=lambda: O.pairwise_marginals(O.SSM(*(a.astype(np.float32) for a in arrays)),
                                                      im.astype(np.float32), ic.astype(np.float32))[1]))
        assert_parity(npy(jc), want_c, TOL[dtype], what="pairwise marginal covariances", **adj)
        assert torch.equal(jc, jc.transpose(-1, -2))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 6])
def test_conditional_statistics_and_predict(d, dtype):
    import markovflow_b200 as mf

    rng = np.random.default_rng(7 + d)
    b, n, m = 3, 50, 11
    a_mt, a_tp = 0.7 * rng.standard_normal((2, b, n, d, d))
    q_mt, q_tp = _spd(rng, (b, n), d), _spd(rng, (b, n), d)
    if dtype == torch.float32:
        a_mt, a_tp, q_mt, q_tp = (x.astype(np.float32).astype(np.float64) for x in (a_mt, a_tp, q_mt, q_tp))
    tol = TOL[dtype]
    for prec in (False, True):
        got = mf.conditional_statistics_from_transitions(tt(a_mt, dtype), tt(q_mt, dtype), tt(a_tp, dtype),
                                                         tt(q_tp, dtype), return_precision=prec)
        want = O.conditional_statistics_from_transitions(a_mt, q_mt, a_tp, q_tp, return_precision=prec)
        for g, w in zip(got, want):
            assert max_rel_err(npy(g), w) < tol
    d_t, e_t, t_t = O.conditional_statistics_from_transitions(a_mt, q_mt, a_tp, q_tp)
    proj = np.concatenate([d_t, e_t], axis=-1)
    pm = rng.standard_normal((b, m, 2 * d))
    pc = _spd(rng, (b, m), 2 * d)
    idx = np.sort(rng.integers(0, m, size=(b, n)), axis=-1)
    if dtype == torch.float32:
        pm, pc = pm.astype(np.float32).astype(np.float64), pc.astype(np.float32).astype(np.float64)
    gm = np.take_along_axis(pm, idx[..., None], axis=1)
    gc = np.take_along_axis(pc, idx[..., None, None], axis=1)
    want_mean, want_cov = O.base_conditional_predict(proj, t_t, gm, gc)
    got_mean, got_cov = mf.conditional_predict_from_transitions(
        tt(idx, torch.int64), tt(a_mt, dtype), tt(q_mt, dtype), tt(a_tp, dtype), tt(q_tp, dtype),
        tt(pm, dtype), tt(pc, dtype))
    assert max_rel_err(npy(got_mean), want_mean) < tol and max_rel_err(npy(got_cov), want_cov) < tol
    # without pairwise covariances: the conditional density; base_conditional_predict = identity gather
    got_mean2, got_cov2 = mf.base_conditional_predict(tt(proj, dtype), tt(t_t, dtype), tt(gm, dtype))
    assert max_rel_err(npy(got_mean2), want_mean) < tol and max_rel_err(npy(got_cov2), t_t) < tol


def test_insertion_indices_match_searchsorted():
    import markovflow_b200 as mf

    train = np.sort(np.random.default_rng(0).uniform(0, 10, size=(2, 20)), axis=-1)
    new = np.sort(np.random.default_rng(1).uniform(-1, 11, size=(2, 33)), axis=-1)
    got = mf.insertion_indices(tt(new), tt(train)).cpu().numpy()
    want = np.stack([np.searchsorted(train[i], new[i]) for i in range(2)])
    np.testing.assert_array_equal(got, want)
