"""GPU parity of ``markovflow_b200.ssm_gaussian_transformations`` against the numpy oracle and the
round-trip identities of the reference's ``tests/unit/test_ssm_gaussian_transformations.py:64-103``.
float64 tolerance 1e-10 (max-abs relative), float32 1e-4 (well-conditioned inputs)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import assert_parity, ld, max_rel_err, random_ssm_arrays

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


TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def f32r(arrays):
    """The float32-representable values of the arrays, as float64 (identical inputs on both sides)."""
    return tuple(np.asarray(a).astype(np.float32).astype(np.float64) for a in arrays)


def _check_params(got, want, tol, truth=None, peer=None, what=""):
    """Every returned array within `tol` (1e-10 float64 / 1e-4 float32) of `want`; `truth` / `peer` are
    callables returning the same five arrays in long double / in float32 arithmetic of the restated
    reference -- evaluated only if an array misses `tol` (tests/helpers.py::assert_parity)."""
    names = ["As", "offsets", "chol_P0", "chol_Qs", "mu0"]
    cache = {}

    def pick(fn, i):
        if fn is None:
            return None
        def get():
            if fn not in cache:
                cache[fn] = fn()
            return cache[fn][i]
        return get

    for i, (n, g, w) in enumerate(zip(names, got, want)):
        assert npy(g).shape == np.asarray(w).shape, n
        assert_parity(npy(g), w, tol, truth=pick(truth, i), peer=pick(peer, i), what=f"{what} {n}")


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5), (5, 3)])
def test_forward_transforms_match_oracle(batch_shape, d, n, dtype, tol):
    import markovflow_b200 as mf

    arrays = random_ssm_arrays(batch_shape, n, d)
    if dtype == torch.float32:
        arrays = f32r(arrays)
    ref = O.SSM(*arrays)
    ssm = make_ssm(arrays, dtype)
    hi = O.SSM(*ld(*arrays))
    lo = O.SSM(*(a.astype(np.float32) for a in arrays))
    for name, fn, ofn in (("ssm_to_expectations", mf.ssm_to_expectations, O.ssm_to_expectations),
                          ("ssm_to_naturals", mf.ssm_to_naturals, O.ssm_to_naturals),
                          ("ssm_to_naturals_no_smoothing", mf.ssm_to_naturals_no_smoothing,
                           O.ssm_to_naturals_no_smoothing)):
        for i, (g, w) in enumerate(zip(fn(ssm), ofn(ref))):
            adj = (dict(truth=lambda: ofn(hi)[i]) if dtype == torch.float64 else dict(peer=lambda: ofn(lo)[i]))
            assert_parity(npy(g), w, tol, what=f"{name}[{i}]", **adj)


@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5), (5, 3)])
def test_inverse_transforms_match_oracle(batch_shape, d, n):
    import markovflow_b200 as mf

    ref = O.SSM(*random_ssm_arrays(batch_shape, n, d))
    for name, fwd, fn, ofn in (
            ("expectations_to_ssm_params", O.ssm_to_expectations, mf.expectations_to_ssm_params,
             O.expectations_to_ssm_params),
            ("naturals_to_ssm_params", O.ssm_to_naturals, mf.naturals_to_ssm_params, O.naturals_to_ssm_params),
            ("naturals_to_ssm_params_no_smoothing", O.ssm_to_naturals_no_smoothing,
             mf.naturals_to_ssm_params_no_smoothing, O.naturals_to_ssm_params_no_smoothing)):
        x = fwd(ref)  # float64 inputs of the inverse transform, identical on both sides
        _check_params(fn(*(tt(v) for v in x)), ofn(*x), 1e-10, truth=lambda: ofn(*ld(*x)), what=name)


@pytest.mark.parametrize("d,n", [(1, 6), (2, 9), (3, 5), (4, 7)])
def test_round_trips_on_device(batch_shape, d, n):
    """tests/unit/test_ssm_gaussian_transformations.py:64-103: transform there and back."""
    import markovflow_b200 as mf

    arrays = random_ssm_arrays(batch_shape, n, d)
    ssm = make_ssm(arrays)
    mu0, l0, a, b, lq = arrays
    want = (a, b, l0, lq, mu0)
    # exact answer known (the SSM the chain started from); the restated reference's own round trip in
    # float64 is the yardstick where 1e-10 is out of reach of the conditioning of the random SSM
    ref = O.SSM(*arrays)
    for name, fwd, inv, ofwd, oinv in (
            ("expectations", mf.ssm_to_expectations, mf.expectations_to_ssm_params,
             O.ssm_to_expectations, O.expectations_to_ssm_params),
            ("naturals", mf.ssm_to_naturals, mf.naturals_to_ssm_params, O.ssm_to_naturals,
             O.naturals_to_ssm_params),
            ("naturals_no_smoothing", mf.ssm_to_naturals_no_smoothing, mf.naturals_to_ssm_params_no_smoothing,
             O.ssm_to_naturals_no_smoothing, O.naturals_to_ssm_params_no_smoothing)):
        _check_params(inv(*fwd(ssm)), want, 1e-10, peer=lambda: oinv(*ofwd(ref)), what=f"round trip {name}")


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
    _check_params(got64, want, 1e-10, truth=lambda: O.naturals_to_ssm_params(*ld(*th)), what="config-5 f64")
    th32 = f32r(th)
    got32 = mf.naturals_to_ssm_params(*(tt(x, torch.float32) for x in th32))
    _check_params(got32, O.naturals_to_ssm_params(*th32), 1e-4, what="config-5 f32",
                  peer=lambda: O.naturals_to_ssm_params(*(x.astype(np.float32) for x in th32)))
    # and onwards to expectations, as the natural-gradient step does
    q = mf.StateSpaceModel(got64[4], got64[2], got64[0], got64[1], got64[3])
    ref_q = O.SSM(want[4], want[2], want[0], want[1], want[3])
    hi_q = O.SSM(*ld(want[4], want[2], want[0], want[1], want[3]))
    for i, (g, w) in enumerate(zip(mf.ssm_to_expectations(q), O.ssm_to_expectations(ref_q))):
        assert_parity(npy(g), w, 1e-10, what=f"config-5 ssm_to_expectations[{i}]",
                      truth=lambda: O.ssm_to_expectations(hi_q)[i])


def test_not_positive_definite_naturals_raise():
    import markovflow_b200 as mf

    ref = O.SSM(*random_ssm_arrays((2,), 4, 2))
    th_lin, th_diag, th_sub = O.ssm_to_naturals(ref)
    th_diag[1, 2] *= -1.0
    with pytest.raises(mf.CholeskyError):
        mf.naturals_to_ssm_params(tt(th_lin), tt(th_diag), tt(th_sub))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
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
    th_np = O.ssm_to_naturals(O.SSM(*arrays))
    if dtype == torch.float32:
        th_np = f32r(th_np)
    want = O.naturals_to_ssm_params(*th_np)  # float64 oracle on the identical (rounded) naturals
    adj = (dict(truth=lambda: O.naturals_to_ssm_params(*ld(*th_np))) if dtype == torch.float64 else
           dict(peer=lambda: O.naturals_to_ssm_params(*(x.astype(np.float32) for x in th_np))))
    th = tuple(tt(x, dtype) for x in th_np)
    lib = _lib.lib()
    res = {}
    for knob in (0, 1):
        lib.mf_set_tuning(2, knob)
        try:
            res[knob] = mf.naturals_to_ssm_params(*th)
        finally:
            lib.mf_set_tuning(2, 0)
        _check_params(res[knob], want, tol, what=f"nat->ssm parallel-in-time knob {knob}", **adj)


def test_naturals_to_ssm_params_parallel_in_time_short_segments_and_failure():
    import markovflow_b200 as mf
    from markovflow_b200 import _lib

    arrays = random_ssm_arrays((2,), 200, 2)
    th_np = O.ssm_to_naturals(O.SSM(*arrays))
    want = O.naturals_to_ssm_params(*th_np)
    th = tuple(tt(x) for x in th_np)
    lib = _lib.lib()
    # knob 3: steps per segment (ragged last segments included); > 64 segments: warp-scan fold
    for seg in (2, 3, 4, 7, 50, 100, 199):
        lib.mf_set_tuning(3, seg)
        try:
            got = mf.naturals_to_ssm_params(*th)
        finally:
            lib.mf_set_tuning(3, 0)
        _check_params(got, want, 1e-10, what=f"nat->ssm segments of {seg}",
                      truth=lambda: O.naturals_to_ssm_params(*ld(*th_np)))
    bad = th[1].clone()
    bad[1, 120] *= -1.0
    with pytest.raises(mf.CholeskyError):
        mf.naturals_to_ssm_params(th[0], bad, th[2])


@pytest.mark.parametrize("b,t", [(300, 700), (8000, 33)])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
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
    th_np = O.ssm_to_naturals(O.SSM(*arrays))
    if dtype == torch.float32:
        th_np = f32r(th_np)
    want = O.naturals_to_ssm_params(*th_np)
    adj = (dict(truth=lambda: O.naturals_to_ssm_params(*ld(*th_np))) if dtype == torch.float64 else
           dict(peer=lambda: O.naturals_to_ssm_params(*(x.astype(np.float32) for x in th_np))))
    th = tuple(tt(x, dtype) for x in th_np)
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
        _check_params(res[knob], want, tol, what=f"many chains, knob 11 = {knob}", **adj)

        def o_back(cast):
            p = O.naturals_to_ssm_params(*(cast(x) for x in th_np))
            return O.expectations_to_ssm_params(*O.ssm_to_expectations(O.SSM(p[4], p[2], p[0], p[1], p[3])))

        _check_params(back, want, tol, what=f"many chains round trip, knob 11 = {knob}",
                      **(dict(truth=lambda: o_back(lambda x: ld(x))) if dtype == torch.float64 else
                         dict(peer=lambda: o_back(lambda x: x.astype(np.float32)))))
