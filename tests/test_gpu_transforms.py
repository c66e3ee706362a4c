"""GPU parity of ``markovflow_b200.ssm_gaussian_transformations`` against the numpy oracle and the
round-trip identities of the reference's ``tests/unit/test_ssm_gaussian_transformations.py:64-103``.
float64 tolerance 1e-10 (max-abs relative), float32 1e-4 (well-conditioned inputs)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import max_rel_err, random_ssm_arrays

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def tt(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), device=dev()).to(dtype)


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def make_ssm(arrays, dtype=torch.float64):
    from markovflow_b200 import StateSpaceModel

    return StateSpaceModel(*(tt(a, dtype) for a in arrays))


def _check_params(got, want, tol):
    names = ["As", "offsets", "chol_P0", "chol_Qs", "mu0"]
    for n, g, w in zip(names, got, want):
        assert npy(g).shape == np.asarray(w).shape, n
        assert max_rel_err(npy(g), w) < tol, n


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-4)])
@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5), (5, 3)])
def test_forward_transforms_match_oracle(batch_shape, d, n, dtype, tol):
    import markovflow_b200 as mf

    arrays = random_ssm_arrays(batch_shape, n, d)
    ref = O.SSM(*arrays)
    ssm = make_ssm(arrays, dtype)
    # expectations: the oracle takes covariances through the precision route (less accurate)
    for g, w in zip(mf.ssm_to_expectations(ssm), O.ssm_to_expectations(ref)):
        assert max_rel_err(npy(g), w) < max(tol, 1e-8)
    for g, w in zip(mf.ssm_to_naturals(ssm), O.ssm_to_naturals(ref)):
        assert max_rel_err(npy(g), w) < tol
    for g, w in zip(mf.ssm_to_naturals_no_smoothing(ssm), O.ssm_to_naturals_no_smoothing(ref)):
        assert max_rel_err(npy(g), w) < tol


@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5), (5, 3)])
def test_inverse_transforms_match_oracle(batch_shape, d, n):
    import markovflow_b200 as mf

    ref = O.SSM(*random_ssm_arrays(batch_shape, n, d))
    eta = O.ssm_to_expectations(ref)
    _check_params(mf.expectations_to_ssm_params(*(tt(x) for x in eta)),
                  O.expectations_to_ssm_params(*eta), 1e-9)
    th = O.ssm_to_naturals(ref)
    _check_params(mf.naturals_to_ssm_params(*(tt(x) for x in th)),
                  O.naturals_to_ssm_params(*th), 1e-9)
    th = O.ssm_to_naturals_no_smoothing(ref)
    _check_params(mf.naturals_to_ssm_params_no_smoothing(*(tt(x) for x in th)),
                  O.naturals_to_ssm_params_no_smoothing(*th), 1e-9)


@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5), (4, 7)])
def test_round_trips_on_device(batch_shape, d, n):
    """tests/unit/test_ssm_gaussian_transformations.py:64-103: transform there and back."""
    import markovflow_b200 as mf

    arrays = random_ssm_arrays(batch_shape, n, d)
    ssm = make_ssm(arrays)
    mu0, l0, a, b, lq = arrays
    want = (a, b, l0, lq, mu0)
    _check_params(mf.expectations_to_ssm_params(*mf.ssm_to_expectations(ssm)), want, 1e-9)
    _check_params(mf.naturals_to_ssm_params(*mf.ssm_to_naturals(ssm)), want, 1e-9)
    _check_params(mf.naturals_to_ssm_params_no_smoothing(*mf.ssm_to_naturals_no_smoothing(ssm)),
                  want, 1e-9)


def test_cvi_style_site_update_config5_shape():
    """BASELINE config 5 in miniature: prior precision + site naturals -> naturals_to_ssm_params
    (models/variational_cvi.py:106-135), float64 against the oracle and float32 against float64."""
    import markovflow_b200 as mf

    rng = np.random.default_rng(5)
    bsz, t = 6, 400
    tp = np.linspace(0.0, 40.0, t)
    lins, diags, subs = [], [], []
    for _ in range(bsz):
        k = O.Matern32(rng.uniform(0.8, 1.2), rng.uniform(0.8, 1.2))
        ssm = k.state_space_model(tp)
        h = k.emission_matrix(tp)
        pd, ps = O.ssm_build_precision(ssm)
        nat1 = rng.standard_normal((t, 1))
        prec = rng.uniform(0.5, 2.0, size=(t, 1, 1))
        # back-projected site naturals (models/variational_cvi.py:423-445)
        lins.append(np.einsum("tmd,tm->td", h, nat1))
        diags.append(-0.5 * (pd + np.einsum("tmd,tmn,tne->tde", h, prec, h)))
        subs.append(-ps)
    th = (np.stack(lins), np.stack(diags), np.stack(subs))
    want = O.naturals_to_ssm_params(*th)
    got64 = mf.naturals_to_ssm_params(*(tt(x) for x in th))
    _check_params(got64, want, 1e-10)
    got32 = mf.naturals_to_ssm_params(*(tt(x, torch.float32) for x in th))
    _check_params(got32, want, 1e-4)
    # and onwards to expectations, as the natural-gradient step does
    q = mf.StateSpaceModel(got64[4], got64[2], got64[0], got64[1], got64[3])
    ref_q = O.SSM(want[4], want[2], want[0], want[1], want[3])
    for g, w in zip(mf.ssm_to_expectations(q), O.ssm_to_expectations(ref_q)):
        assert max_rel_err(npy(g), w) < 1e-9


def test_not_positive_definite_naturals_raise():
    import markovflow_b200 as mf

    ref = O.SSM(*random_ssm_arrays((2,), 4, 2))
    th_lin, th_diag, th_sub = O.ssm_to_naturals(ref)
    th_diag[1, 2] *= -1.0
    with pytest.raises(mf.CholeskyError):
        mf.naturals_to_ssm_params(tt(th_lin), tt(th_diag), tt(th_sub))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-4)])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
@pytest.mark.parametrize("b,t", [(1, 1000), (3, 301), (2, 130), (5, 640)])
def test_naturals_to_ssm_params_parallel_in_time(b, t, d, dtype, tol):
    """Few long chains: the backward U D U^T recursion is evaluated parallel in time (segment
    elements of the linear-fractional map -> seeds -> seeded sweeps, ssm_sweep.cuh).  Must agree
    with the sequential sweep (tuning knob 2 = 1) and recover the SSM the naturals came from."""
    import markovflow_b200 as mf
    from markovflow_b200 import _lib

    state = np.random.get_state()
    np.random.seed(b * 7919 + t * 13 + d)
    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
    np.random.set_state(state)
    mu0, l0, a, bb, lq = arrays
    want = (a, bb, l0, lq, mu0)
    th = mf.ssm_to_naturals(make_ssm(arrays))  # float64 naturals on the device
    th = tuple(x.to(dtype) for x in th)
    lib = _lib.lib()
    res = {}
    for knob in (0, 1):
        lib.mf_set_tuning(2, knob)
        try:
            res[knob] = mf.naturals_to_ssm_params(*th)
        finally:
            lib.mf_set_tuning(2, 0)
        _check_params(res[knob], want, tol)
    for g, w in zip(res[0], res[1]):
        assert max_rel_err(npy(g), npy(w)) < tol


def test_naturals_to_ssm_params_parallel_in_time_short_segments_and_failure():
    import markovflow_b200 as mf
    from markovflow_b200 import _lib

    arrays = random_ssm_arrays((2,), 200, 2)
    mu0, l0, a, bb, lq = arrays
    th = mf.ssm_to_naturals(make_ssm(arrays))
    lib = _lib.lib()
    # knob 3: steps per segment (ragged last segments included); > 64 segments: warp-scan fold
    for seg in (2, 3, 4, 7, 50, 100, 199):
        lib.mf_set_tuning(3, seg)
        try:
            got = mf.naturals_to_ssm_params(*th)
        finally:
            lib.mf_set_tuning(3, 0)
        _check_params(got, (a, bb, l0, lq, mu0), 1e-9)
    bad = th[1].clone()
    bad[1, 120] *= -1.0
    with pytest.raises(mf.CholeskyError):
        mf.naturals_to_ssm_params(th[0], bad, th[2])


@pytest.mark.parametrize("b,t", [(300, 700), (8000, 33)])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-4)])
def test_naturals_to_ssm_params_many_chains_tile_geometries(b, t, dtype, tol):
    """More than 148 x 48 virtual chains at D = 2 (the config-5 regime): the float64 naturals -> SSM
    sweep runs on 4-step tiles and output-less float32 passes on 16-step tiles; tuning knob 11 = 1
    restores 8-step tiles everywhere.  Both must recover the SSM the naturals came from and agree;
    ssm_to_expectations -> expectations_to_ssm_params closes the loop on the same sizes."""
    import markovflow_b200 as mf
    from markovflow_b200 import _lib

    d = 2
    state = np.random.get_state()
    np.random.seed(b + t)
    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
    np.random.set_state(state)
    mu0, l0, a, bb, lq = arrays
    want = (a, bb, l0, lq, mu0)
    th = tuple(x.to(dtype) for x in mf.ssm_to_naturals(make_ssm(arrays)))
    lib = _lib.lib()
    res = {}
    for knob in (0, 1):
        lib.mf_set_tuning(11, knob)
        try:
            res[knob] = mf.naturals_to_ssm_params(*th)
            q = mf.StateSpaceModel(res[knob][4], res[knob][2], res[knob][0], res[knob][1], res[knob][3])
            back = mf.expectations_to_ssm_params(*mf.ssm_to_expectations(q))
        finally:
            lib.mf_set_tuning(11, 0)
        _check_params(res[knob], want, tol)
        _check_params(back, want, 10 * tol)
    for g, w in zip(res[0], res[1]):
        assert max_rel_err(npy(g), npy(w)) < tol
