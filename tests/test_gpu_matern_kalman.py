"""GPU parity of the Matern-prior Kalman log-likelihood with in-kernel SSM construction
(``mf_kalman_matern_log_likelihood``) against the numpy oracle's
kernel -> SSM -> ``KalmanFilter.log_likelihood`` chain (reference ``kernels/sde_kernel.py:153-171``,
``kernels/matern.py``, ``kalman_filter.py:184-255``), the materialised-SSM CUDA path, and the
dense-GP closed form of ``tests/integration/models/test_gaussian_process_regression.py:78-115``.
float64 tolerance 1e-10 (max-abs relative), float32 1e-4."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import max_rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 1e-10, torch.float32: 1e-4}
KERNELS = {1: O.Matern12, 2: O.Matern32, 3: O.Matern52}


def tt(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda:0").to(dtype)


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def _case(d, b, t, rng, jitter=0.0):
    ls = rng.uniform(0.5, 2.0, size=b)
    var = rng.uniform(0.5, 2.0, size=b)
    tps = np.cumsum(rng.uniform(0.05, 0.3, size=(b, t)), axis=-1)
    y = np.sin(tps) + 0.1 * rng.standard_normal((b, t))
    return ls, var, tps, y


def _oracle(d, ls, var, tps, y, noise, jitter=0.0, extended=False):
    """Reference chain kernel -> SSM -> KalmanFilter.log_likelihood per series; ``extended``: the
    transition statistics in long double (``O.stationary_ssm_extended_precision``)."""
    out = []
    for c in range(len(ls)):
        kern = KERNELS[d](ls[c], var[c], jitter=jitter)
        ssm = (O.stationary_ssm_extended_precision(kern, tps[c]) if extended
               else kern.state_space_model(tps[c]))
        out.append(O.kalman_log_likelihood(ssm, kern.emission_matrix(tps[c]), y[c][:, None],
                                           np.array([[1.0 / noise ** 2]])))
    return np.array(out)


def _dense_gp(d, ls, var, tps, y, noise):
    """log N(y; 0, K + noise^2 I) with the closed-form Matern kernel matrix K -- the known answer of
    tests/integration/models/test_gaussian_process_regression.py:78-115; well conditioned, so good to
    ~1e-13 where the banded / SSM forms are not."""
    out = []
    for c in range(len(ls)):
        r = np.abs(tps[c][:, None] - tps[c][None, :])
        n = len(tps[c])
        lam = {1: 1.0, 2: np.sqrt(3), 3: np.sqrt(5)}[d] / ls[c]
        poly = {1: 1.0, 2: 1 + lam * r, 3: 1 + lam * r + lam ** 2 * r ** 2 / 3}[d]
        chol = np.linalg.cholesky(var[c] * poly * np.exp(-lam * r) + noise ** 2 * np.eye(n))
        z = np.linalg.solve(chol, y[c])
        out.append(-0.5 * (z @ z) - np.log(np.diag(chol)).sum() - 0.5 * n * np.log(2 * np.pi))
    return np.array(out)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("b,t", [(1, 1), (1, 2), (3, 17), (5, 300), (2, 1000)])
def test_matches_oracle_chain_and_dense_gp(d, b, t, dtype):
    from markovflow_b200 import matern_kalman_log_likelihood

    rng = np.random.default_rng(100 * d + t)
    ls, var, tps, y = _case(d, b, t, rng)
    if dtype == torch.float32:
        # time points up to ~300 lose the digits of dt when differenced in f32: pass the deltas
        got = matern_kalman_log_likelihood(d, tt(ls, dtype), tt(var, dtype), tt(y, dtype), 0.2,
                                           time_deltas=tt(np.diff(tps, axis=-1), dtype))
    else:
        got = matern_kalman_log_likelihood(d, tt(ls, dtype), tt(var, dtype), tt(y, dtype), 0.2,
                                           time_points=tt(tps, dtype))
    assert got.shape == (b,)
    want = _oracle(d, ls, var, tps, y, 0.2)
    dense = _dense_gp(d, ls, var, tps, y, 0.2)
    if dtype == torch.float32:
        assert max_rel_err(npy(got), dense) < TOL[dtype]
        return
    # The known answer is the dense closed form.  The restated reference chain (Q_k = Pinf - A Pinf A^T,
    # chol Q_k, precision form with Q_k^-1) is ill conditioned for Matern52 at small dt -- in numpy as
    # in TF -- and sits up to 3e-9 away from it; against that chain the bar is 1e-10 plus the chain's
    # own distance from the known answer.
    assert max_rel_err(npy(got), dense) < TOL[dtype]
    assert max_rel_err(npy(got), want) < TOL[dtype] + 2.0 * max_rel_err(want, dense)


@pytest.mark.parametrize("d", [1, 2])
def test_kernel_jitter_matches_oracle_chain(d):
    """jitter·I on P0 and every Q_k (sde_kernel.py:83-96,444-446)."""
    from markovflow_b200 import matern_kalman_log_likelihood

    rng = np.random.default_rng(5 + d)
    ls, var, tps, y = _case(d, 3, 400, rng)
    got = matern_kalman_log_likelihood(d, tt(ls), tt(var), tt(y), 0.2, time_points=tt(tps), jitter=1e-3)
    want = _oracle(d, ls, var, tps, y, 0.2, jitter=1e-3)
    assert max_rel_err(npy(got), want) < 1e-10
    plain = _oracle(d, ls, var, tps, y, 0.2)
    assert max_rel_err(want, plain) > 1e-6  # the jitter is visible at this size


@pytest.mark.parametrize("d", [1, 2, 3])
def test_dense_gp_closed_form(d):
    from markovflow_b200 import Matern12, Matern32, Matern52

    rng = np.random.default_rng(7)
    tp = np.cumsum(rng.uniform(0.05, 0.4, size=60))
    y = np.sin(tp) + 0.1 * rng.standard_normal(60)
    l, v, noise = 0.7, 1.9, 0.3
    r = np.abs(tp[:, None] - tp[None, :])
    lam = {1: 1.0, 2: np.sqrt(3), 3: np.sqrt(5)}[d] / l
    poly = {1: 1.0, 2: 1 + lam * r, 3: 1 + lam * r + lam ** 2 * r ** 2 / 3}[d]
    c = v * poly * np.exp(-lam * r) + noise ** 2 * np.eye(60)
    want = -0.5 * (y @ np.linalg.solve(c, y) + np.linalg.slogdet(c)[1] + 60 * np.log(2 * np.pi))
    kern = {1: Matern12, 2: Matern32, 3: Matern52}[d](l, v)
    got = kern.kalman_log_likelihood(tt(tp), tt(y)[:, None], tt([[noise]]))
    assert abs(float(got) - want) < 1e-10 * abs(want)


@pytest.mark.parametrize("d,b,t", [(2, 1, 200000), (3, 2, 50001), (1, 1, 100000), (2, 40, 5000)])
def test_parallel_in_time_equals_one_chain_per_series_and_materialised_ssm(d, b, t):
    """Few long series are cut into segments (range elements -> per-warp joins -> ordered reduction):
    same value as the uncut filter, for every launch geometry, and as the materialised-SSM kernel."""
    from markovflow_b200 import StateSpaceModel, _lib, kalman_log_likelihood, matern_kalman_log_likelihood

    rng = np.random.default_rng(t)
    ls, var, tps, y = _case(d, b, t, rng)
    dts = np.diff(tps, axis=-1)
    args = (d, tt(ls), tt(var), tt(y), 0.1)
    lib = _lib.lib()
    cut = npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts)))
    try:
        lib.mf_set_tuning(2, 1)
        uncut = npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts)))
        lib.mf_set_tuning(2, 0)
        lib.mf_set_tuning(3, 37)  # odd segment length, ragged tail
        odd = npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts)))
        lib.mf_set_tuning(3, 0)
        staged = []
        if d == 2:
            for variant in (1, 2, 3, 4):
                lib.mf_set_tuning(8, variant)
                staged.append(npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts))))
                lib.mf_set_tuning(2, 1)
                staged.append(npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts))))
                lib.mf_set_tuning(2, 0)
    finally:
        for k in (2, 3, 8):
            lib.mf_set_tuning(k, 0)
    assert max_rel_err(cut, uncut) < 1e-10
    assert max_rel_err(odd, uncut) < 1e-10
    for s in staged:
        assert max_rel_err(s, uncut) < 1e-10
    # the materialised path on the oracle-built SSM
    ssms = [KERNELS[d](ls[c], var[c]).state_space_model(tps[c]) for c in range(b)]
    stack = lambda f: tt(np.stack([f(s) for s in ssms]))
    gssm = StateSpaceModel(stack(lambda s: s.mu0), stack(lambda s: s.chol_p0), stack(lambda s: s.a_s),
                           stack(lambda s: s.b_s), stack(lambda s: s.chol_q_s))
    h = np.zeros((t, 1, d)); h[:, 0, 0] = 1.0
    mat = npy(kalman_log_likelihood(gssm, tt(h), tt(y)[..., None], tt([[0.1]])))
    assert max_rel_err(cut, mat) < 1e-10


def test_time_sharded_elements_fold_to_the_whole_series():
    """first_is_initial = 0 segments: elements of consecutive time segments, joined with
    mf_kalman_fold_elements, carry the whole series' log-likelihood in ell (multi-GPU protocol)."""
    from markovflow_b200 import matern_kalman_log_likelihood
    from markovflow_b200.parallel import CudaKalmanEngine

    rng = np.random.default_rng(11)
    d, t = 2, 30011
    ls, var, tps, y = _case(d, 1, t, rng)
    dts = np.diff(tps, axis=-1)
    whole = npy(matern_kalman_log_likelihood(d, tt(ls), tt(var), tt(y), 0.1, time_deltas=tt(dts)))
    for world in (2, 3, 8):
        bounds = np.linspace(0, t, world + 1).astype(int)
        elems = []
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            first = r == 0
            seg_dt = dts[:, lo:hi - 1] if first else dts[:, lo - 1:hi - 1]
            _, e = matern_kalman_log_likelihood(d, tt(ls), tt(var), tt(y[:, lo:hi]), 0.1,
                                                time_deltas=tt(seg_dt), first_is_initial=first,
                                                return_element=True)
            elems.append(e)
        total = CudaKalmanEngine().fold(torch.stack(elems), d)
        assert max_rel_err(npy(total[:, -1]), whole) < 1e-10


def test_rejects_bad_arguments():
    from markovflow_b200 import matern_kalman_log_likelihood

    y = tt(np.zeros((2, 5)))
    with pytest.raises(ValueError):
        matern_kalman_log_likelihood(4, 1.0, 1.0, y, 0.1, time_points=tt(np.zeros((2, 5))))
    with pytest.raises(ValueError):
        matern_kalman_log_likelihood(2, 1.0, 1.0, y, 0.1)
    with pytest.raises(ValueError):
        matern_kalman_log_likelihood(2, 1.0, 1.0, y, 0.1, time_points=tt(np.zeros((2, 4))))
    with pytest.raises(ValueError):
        matern_kalman_log_likelihood(2, tt(np.ones(3)), 1.0, y, 0.1, time_points=tt(np.zeros((2, 5))))
    with pytest.raises(RuntimeError):
        matern_kalman_log_likelihood(2, 1.0, 1.0, torch.zeros(2, 5, dtype=torch.float64), 0.1,
                                     time_points=torch.zeros(2, 5, dtype=torch.float64))


def test_element_of_uncut_series_and_many_series():
    """P == 1 paths: the element written directly by the summary kernel (ell copied out of it), and a
    batch large enough that series are not cut at all (plain filter core)."""
    from markovflow_b200 import _lib, matern_kalman_log_likelihood

    rng = np.random.default_rng(3)
    d, b, t = 2, 3, 700
    ls, var, tps, y = _case(d, b, t, rng)
    dts = np.diff(tps, axis=-1)
    args = (d, tt(ls), tt(var), tt(y), 0.1)
    ll_cut, el_cut = matern_kalman_log_likelihood(*args, time_deltas=tt(dts), return_element=True)
    lib = _lib.lib()
    try:
        lib.mf_set_tuning(2, 1)
        ll_one, el_one = matern_kalman_log_likelihood(*args, time_deltas=tt(dts), return_element=True)
    finally:
        lib.mf_set_tuning(2, 0)
    assert max_rel_err(npy(ll_one), npy(ll_cut)) < 1e-10
    assert max_rel_err(npy(el_one), npy(el_cut)) < 1e-10
    assert max_rel_err(npy(el_one[:, -1]), npy(ll_one)) == 0.0
    # many series: 2500 > 148 * 12 warps, one chain per series
    b2, t2 = 2500, 40
    ls, var, tps, y = _case(d, b2, t2, rng)
    got = npy(matern_kalman_log_likelihood(d, tt(ls), tt(var), tt(y), 0.2, time_points=tt(tps)))
    pick = np.array([0, 1, 777, 2499])
    want = _dense_gp(d, ls[pick], var[pick], tps[pick], y[pick], 0.2)
    assert max_rel_err(got[pick], want) < 1e-10
    assert np.all(np.isfinite(got))


def test_empty_batch():
    from markovflow_b200 import matern_kalman_log_likelihood

    out = matern_kalman_log_likelihood(2, tt(np.zeros(0)), tt(np.zeros(0)), tt(np.zeros((0, 5))), 0.1,
                                       time_deltas=tt(np.zeros((0, 4))))
    assert tuple(out.shape) == (0,)


def test_mid_size_batch_is_cut_in_time_and_huge_batch_is_not_limited():
    """2000 series cannot fill the GPU with one thread each: they are cut into 32 segments per series
    (same value as the uncut filter and the dense closed form); 70000 series run one chain each."""
    from markovflow_b200 import _lib, matern_kalman_log_likelihood

    rng = np.random.default_rng(21)
    d, b, t = 2, 2000, 2100
    ls, var, tps, y = _case(d, b, t, rng)
    dts = np.diff(tps, axis=-1)
    args = (d, tt(ls), tt(var), tt(y), 0.15)
    cut = npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts)))
    lib = _lib.lib()
    try:
        lib.mf_set_tuning(2, 1)
        uncut = npy(matern_kalman_log_likelihood(*args, time_deltas=tt(dts)))
    finally:
        lib.mf_set_tuning(2, 0)
    assert max_rel_err(cut, uncut) < 1e-10
    pick = np.array([0, 999, 1999])
    assert max_rel_err(cut[pick], _dense_gp(d, ls[pick], var[pick], tps[pick], y[pick], 0.15)) < 1e-10
    b2, t2 = 70000, 10
    ls, var, tps, y = _case(3, b2, t2, rng)
    got = npy(matern_kalman_log_likelihood(3, tt(ls), tt(var), tt(y), 0.2, time_points=tt(tps)))
    pick = np.array([0, 65535, 65536, 69999])
    assert max_rel_err(got[pick], _dense_gp(3, ls[pick], var[pick], tps[pick], y[pick], 0.2)) < 1e-10
    assert np.all(np.isfinite(got))
