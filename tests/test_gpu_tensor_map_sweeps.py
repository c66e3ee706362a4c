"""The tensor-map sweep engine (csrc/sweep_tm.cuh: one ``cp.async.bulk.tensor`` per stream, tile and CTA) against
the numpy oracle and against the 1-D bulk-copy engine (tuning knob 13 = 1) on the transforms it serves:
``naturals_to_ssm_params`` (backward sweep, ssm_gaussian_transformations.py:196-262 of the reference) and
``ssm_to_expectations`` / marginals (forward sweep, :31-72; state_space_model.py:202-257).

Geometry grid: whole chains as rows (P = 1; batch sizes that leave ragged CTAs), P >= 3 segments per chain with
P below / above the rows of a CTA, ragged last segments, every tile geometry (knob 14), float64 (1e-10) and
float32 (1e-4).  ``mf_tm_launch_count`` proves which engine served the call."""
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


class knobs:
    """Sets tuning knobs for the duration of a ``with`` block."""

    def __init__(self, **kv):
        self.kv = {int(k[1:]): v for k, v in kv.items()}

    def __enter__(self):
        from markovflow_b200 import _lib

        self.lib = _lib.lib()
        # knob 13 = 2: no minimum problem size for the tensor-map engine (it is selected from 148 x 64 rows on)
        self.kv.setdefault(13, 2)
        for k, v in self.kv.items():
            self.lib.mf_set_tuning(k, v)
        return self.lib

    def __exit__(self, *exc):
        for k in self.kv:
            self.lib.mf_set_tuning(k, 0)


def tm_count():
    import ctypes

    from markovflow_b200 import _lib

    f = _lib.lib().mf_tm_launch_count
    f.restype = ctypes.c_int64
    return int(f())


def naturals_case(b, t, d, dtype, seed):
    state = np.random.get_state()
    np.random.seed(seed)
    arrays = random_ssm_arrays((b,), t - 1, d, scale_a=0.6 / np.sqrt(d))
    np.random.set_state(state)
    th_np = O.ssm_to_naturals(O.SSM(*arrays))
    if dtype == torch.float32:
        th_np = f32r(th_np)
    return arrays, th_np


def check_nat(got, th_np, dtype, what):
    want = O.naturals_to_ssm_params(*th_np)
    cache = {}

    def other(i):
        def get():
            if "v" not in cache:
                cache["v"] = (O.naturals_to_ssm_params(*ld(*th_np)) if dtype == torch.float64 else
                              O.naturals_to_ssm_params(*(x.astype(np.float32) for x in th_np)))
            return cache["v"][i]
        return get

    for i, (g, w) in enumerate(zip(got, want)):
        kw = dict(truth=other(i)) if dtype == torch.float64 else dict(peer=other(i))
        assert_parity(npy(g), w, TOL[dtype], what=f"{what} [{i}]", **kw)


# (b, t, steps per segment or 0 = whole chains): P = ceil(t / seg)
GEOMETRIES = [
    (1, 1000, 36),    # P = 28 <= 64: two chains' worth of rows per CTA would need b >= 2; one ragged CTA
    (5, 1000, 36),    # P = 28, CTAs of 56 rows = 2 chains, last CTA ragged
    (3, 640, 10),     # P = 64 = rows of a CTA
    (2, 1283, 10),    # P = 129 (prime factors 3, 43): rows per CTA = 43
    (2, 1280, 10),    # P = 128 = 2 CTAs per chain
    (4, 700, 22),     # P = 32, ragged last segment of 18 steps
    (3, 97, 8),       # P = 13: 4 chains per CTA would need 8 special rows > 4 -> 2 chains, 26 rows < 32: 1-D engine
    (2, 40, 12),      # P = 4: 1-D engine
    (37, 300, 0),     # whole chains, ragged CTA
    (200, 33, 0),
    (10000, 5, 0),    # more rows than 148 x 64
]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2])
@pytest.mark.parametrize("b,t,seg", GEOMETRIES)
def test_naturals_to_ssm_params_tensor_map_engine(b, t, seg, d, dtype):
    import markovflow_b200 as mf

    arrays, th_np = naturals_case(b, t, d, dtype, 1000 * b + t + d)
    th = tuple(tt(x, dtype) for x in th_np)
    kw = dict(k3=seg) if seg else dict(k2=1)
    with knobs(**kw):
        n0 = tm_count()
        got = mf.naturals_to_ssm_params(*th)
        used = tm_count() - n0
    with knobs(k13=1, **kw):
        n0 = tm_count()
        ref = mf.naturals_to_ssm_params(*th)
        assert tm_count() == n0
    check_nat(got, th_np, dtype, f"tensor-map engine ({used} launches) b={b} t={t} seg={seg}")
    for g, r in zip(got, ref):
        # same arithmetic in the same order on both engines
        assert torch.equal(g, r)


def test_tensor_map_engine_serves_the_config5_shapes():
    """D = 2, B x P >> 148 x 64 rows: both passes of naturals_to_ssm_params and (float64) of ssm_to_expectations run
    on the tensor-map engine, for every tile geometry."""
    import markovflow_b200 as mf

    b, t, d = 64, 2000, 2
    for dtype in (torch.float64, torch.float32):
        arrays, th_np = naturals_case(b, t, d, dtype, 5)
        th = tuple(tt(x, dtype) for x in th_np)
        for geom in (0, 1, 2, 3, 4, 8):
            with knobs(k14=geom):
                n0 = tm_count()
                got = mf.naturals_to_ssm_params(*th)
                q = mf.StateSpaceModel(got[4], got[2], got[0], got[1], got[3])
                exp = mf.ssm_to_expectations(q)
                # float32: the forward passes read `b` ([B, T-1, 2] float32: chains 8 bytes apart mod 16), which a
                # tensor map cannot describe -- they stay on the 1-D engine
                assert tm_count() - n0 == (4 if dtype == torch.float64 else 2), (dtype, geom)
            check_nat(got, th_np, dtype, f"config-5 shape, geometry {geom}")
            p = O.naturals_to_ssm_params(*th_np)
            ref = O.SSM(p[4], p[2], p[0], p[1], p[3])
            for i, (g, w) in enumerate(zip(exp, O.ssm_to_expectations(ref))):
                assert_parity(npy(g), w, 10 * TOL[dtype], what=f"expectations[{i}] geometry {geom}")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2])
@pytest.mark.parametrize("b,t,seg", GEOMETRIES)
def test_forward_moment_sweeps_tensor_map_engine(b, t, seg, d, dtype):
    """ssm_to_expectations and marginals (forward sweeps; 'incoming' transition streams start one element early)."""
    import markovflow_b200 as mf

    arrays, _ = naturals_case(b, t, d, dtype, 77 * b + t + d)
    if dtype == torch.float32:
        arrays = f32r(arrays)
    ref = O.SSM(*arrays)
    ssm = mf.StateSpaceModel(*(tt(a, dtype) for a in arrays))
    kw = dict(k3=seg) if seg else dict(k2=1)
    with knobs(**kw):
        got = mf.ssm_to_expectations(ssm)
        mu, cov = ssm.marginals
    with knobs(k13=1, **kw):
        other = mf.ssm_to_expectations(ssm)
        mu1, cov1 = ssm.marginals
    hi = O.SSM(*ld(*arrays))
    lo = O.SSM(*(a.astype(np.float32) for a in arrays))
    for i, (g, w) in enumerate(zip(got, O.ssm_to_expectations(ref))):
        kw2 = (dict(truth=lambda i=i: O.ssm_to_expectations(hi)[i]) if dtype == torch.float64 else
               dict(peer=lambda i=i: O.ssm_to_expectations(lo)[i]))
        assert_parity(npy(g), w, TOL[dtype], what=f"expectations[{i}] b={b} t={t} seg={seg}", **kw2)
    for g, r in zip(list(got) + [mu, cov], list(other) + [mu1, cov1]):
        assert torch.equal(g, r)
    assert_parity(npy(mu), O.ssm_marginal_means(ref), TOL[dtype], what="marginal means",
                  **(dict(truth=lambda: O.ssm_marginal_means(hi)) if dtype == torch.float64 else
                     dict(peer=lambda: O.ssm_marginal_means(lo))))


def test_tensor_map_engine_declines_unaligned_views():
    """A view whose chains are not 16-byte aligned cannot be described by a tensor map: the call is served by the
    1-D engine (any alignment) with the same result."""
    import markovflow_b200 as mf

    arrays, th_np = naturals_case(6, 400, 2, torch.float64, 3)
    th = tuple(tt(x) for x in th_np)

    def shifted(x):
        buf = torch.empty(x.numel() + 1, device=x.device, dtype=x.dtype)
        v = buf[1:].view(x.shape)
        v.copy_(x)
        return v

    th_odd = (shifted(th[0]), th[1], th[2])
    assert th_odd[0].data_ptr() % 16 == 8
    with knobs(k3=20):
        n0 = tm_count()
        a = mf.naturals_to_ssm_params(*th)
        n1 = tm_count()
        b = mf.naturals_to_ssm_params(*th_odd)
        n2 = tm_count()
    assert n1 - n0 == 2 and n2 == n1
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_tensor_map_engine_equals_the_1d_engine_on_random_geometries():
    """Seeded random (chains, steps, segment length) triples, D = 2, float64: whatever geometry the launcher picks or
    declines, both passes of naturals_to_ssm_params and of ssm_to_expectations return bit-identical results on the
    two engines (same arithmetic, different data movement), and the tensor-map engine serves most of them."""
    import markovflow_b200 as mf

    rng = np.random.default_rng(2026)
    served = 0
    for case_id in range(24):
        b = int(rng.integers(1, 40))
        t = int(rng.integers(40, 1500))
        seg = int(rng.integers(8, 64)) if case_id % 4 else 0
        arrays, th_np = naturals_case(b, t, 2, torch.float64, 31 * case_id + 7)
        th = tuple(tt(x) for x in th_np)
        kw = dict(k3=seg) if seg else dict(k2=1)
        with knobs(**kw):
            n0 = tm_count()
            got = mf.naturals_to_ssm_params(*th)
            q = mf.StateSpaceModel(got[4], got[2], got[0], got[1], got[3])
            exp = mf.ssm_to_expectations(q)
            served += tm_count() - n0
        with knobs(k13=1, **kw):
            ref = mf.naturals_to_ssm_params(*th)
            q1 = mf.StateSpaceModel(ref[4], ref[2], ref[0], ref[1], ref[3])
            exp1 = mf.ssm_to_expectations(q1)
        for g, r in zip(list(got) + list(exp), list(ref) + list(exp1)):
            assert torch.equal(g, r), (case_id, b, t, seg)
        check_nat(got, th_np, torch.float64, f"random geometry {case_id}: b={b} t={t} seg={seg}")
    assert served >= 24  # at least a quarter of the 96 sweeps ran on tensor maps
