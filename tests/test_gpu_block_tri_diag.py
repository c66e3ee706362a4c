"""GPU parity: ``markovflow_b200.block_tri_diag`` (CUDA kernels through the C ABI) against dense
``numpy.linalg`` -- the reference's own test identities (``tests/unit/test_block_tri_diag.py``) --
and against the numpy oracle.  float64 tolerance 1e-10 (max-abs relative to max-abs of the reference
array), float32 1e-4."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import ld as as_ld
from tests.helpers import (
    assert_parity,
    blocks_from_dense,
    max_rel_err,
    random_lower_btd,
    random_spd_btd,
    random_well_conditioned_spd_btd,
)

pytestmark = pytest.mark.gpu

INNER = [1, 3]
OUTER = [1, 4]
TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def dev():
    return torch.device("cuda:0")


def tt(x, dtype=torch.float64):
    return None if x is None else torch.as_tensor(np.ascontiguousarray(x), device=dev()).to(dtype)


def npy(x):
    return x.detach().cpu().numpy().astype(np.float64)


def _mods():
    from markovflow_b200 import LowerTriangularBlockTriDiagonal, SymmetricBlockTriDiagonal

    return LowerTriangularBlockTriDiagonal, SymmetricBlockTriDiagonal


def _skip(with_sub, t):
    return with_sub and t == 1


def test_library_loaded_is_in_tree():
    from markovflow_b200 import _lib

    assert _lib.lib().mf_version() >= 100
    assert _lib.LIB_PATH.endswith("markovflow_b200/csrc/libmarkovflow_b200.so")


@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", [2, 4])
def test_block_sub_diag(batch_shape, d, t):
    L, _ = _mods()
    diag = np.random.normal(size=batch_shape + (t, d, d))
    sub = np.random.normal(size=batch_shape + (t - 1, d, d))
    m = L(tt(diag), tt(sub))
    np.testing.assert_allclose(npy(m.block_sub_diagonal), sub)
    assert m.inner_dim == d and m.outer_dim == t and tuple(m.batch_shape) == batch_shape
    assert m.bandwidth == 2 * d - 1


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_dense_and_band(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    L, S = _mods()
    dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    m = L(tt(diag), tt(sub))
    np.testing.assert_allclose(npy(m.to_dense()), dense)
    band = npy(m.as_band)
    for r in range(band.shape[-2]):
        for j in range(t * d - r):
            np.testing.assert_array_equal(band[..., r, j], dense[..., j + r, j])
    dense_s, diag_s, sub_s = random_spd_btd(batch_shape, t, d, with_sub)
    np.testing.assert_allclose(npy(S(tt(diag_s), tt(sub_s)).to_dense()), dense_s)


@pytest.mark.parametrize("sub1", [True, False])
@pytest.mark.parametrize("sub2", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_add(batch_shape, sub1, sub2, d, t):
    if (sub1 or sub2) and t == 1:
        return
    _, S = _mods()
    d1, diag1, s1 = random_spd_btd(batch_shape, t, d, sub1)
    d2, diag2, s2 = random_spd_btd(batch_shape, t, d, sub2)
    added = (S(tt(diag1), tt(s1)) + S(tt(diag2), tt(s2))).to_dense()
    np.testing.assert_allclose(npy(added), d1 + d2, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_abs_log_det(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    L, _ = _mods()
    dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    got = npy(L(tt(diag), tt(sub)).abs_log_det())
    np.testing.assert_allclose(got, np.linalg.slogdet(dense)[1], rtol=1e-10, atol=1e-12)


def test_abs_log_det_long_chain_uses_segments():
    L, _ = _mods()
    rng = np.random.default_rng(1)
    diag = np.tril(rng.standard_normal((3, 20000, 2, 2))) + 2.0 * np.eye(2)
    got = npy(L(tt(diag)).abs_log_det())
    np.testing.assert_allclose(got, O.btd_abs_log_det(diag), rtol=1e-11)


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_cholesky_reference_generator(batch_shape, with_sub, d, t):
    """tests/unit/test_block_tri_diag.py:88-98 (rtol 1e-3 there) + oracle parity at 1e-10."""
    if _skip(with_sub, t):
        return
    _, S = _mods()
    dense, diag, sub = random_spd_btd(batch_shape, t, d, with_sub)
    chol = S(tt(diag), tt(sub)).cholesky
    np.testing.assert_allclose(npy(chol.to_dense()), np.linalg.cholesky(dense), rtol=1e-3, atol=1e-6)
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    # the reference generator is ill-conditioned; compare with the oracle relative to conditioning
    recon = npy(chol.to_dense())
    recon = recon @ np.swapaxes(recon, -1, -2)
    assert max_rel_err(recon, dense) < 1e-12
    if with_sub:
        assert chol.block_sub_diagonal is not None
    else:
        assert chol.block_sub_diagonal is None


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("b,t", [(1, 1), (5, 2), (70, 33), (3, 257)])
def test_cholesky_solve_logdet_vs_oracle(dtype, d, b, t):
    _, S = _mods()
    if t == 1:
        rng = np.random.default_rng(d)
        ld = np.tril(rng.standard_normal((b, 1, d, d)), -1) * 0.3 + 1.5 * np.eye(d)
        diag, sub = ld @ np.swapaxes(ld, -1, -2), None
    else:
        diag, sub, _, _ = random_well_conditioned_spd_btd((b,), t, d, rng=100 * d + t)
    rhs = np.random.default_rng(7).standard_normal((b, t, d))
    if dtype == torch.float32:  # oracle on the float32-rounded inputs
        diag = diag.astype(np.float32).astype(np.float64)
        sub = None if sub is None else sub.astype(np.float32).astype(np.float64)
        rhs = rhs.astype(np.float32).astype(np.float64)
    m = S(tt(diag, dtype), tt(sub, dtype))
    chol, x, logdet = m.cholesky_and_solve(tt(rhs, dtype), want_log_det=True)
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    o_x = O.btd_solve(o_ld, o_ls, rhs)
    tol = TOL[dtype]
    assert max_rel_err(npy(chol.block_diagonal), o_ld) < tol
    if sub is not None:
        assert max_rel_err(npy(chol.block_sub_diagonal), o_ls) < tol
    assert max_rel_err(npy(x), o_x) < tol
    assert max_rel_err(npy(logdet), O.btd_abs_log_det(o_ld)) < tol
    # upper triangles of the factor's diagonal blocks are exactly zero (band_to_block symmetric=False)
    assert float(torch.triu(chol.block_diagonal, 1).abs().max()) == 0.0
    # separate (unfused) calls agree with the fused sweep
    chol2 = m.cholesky
    assert torch.equal(chol2.block_diagonal, chol.block_diagonal)
    x2 = chol2.solve(tt(rhs, dtype))
    assert max_rel_err(npy(x2), o_x) < tol
    assert max_rel_err(npy(chol2.abs_log_det()), O.btd_abs_log_det(o_ld)) < tol


def test_cholesky_reads_lower_triangle_only_and_in_place_alias():
    _, S = _mods()
    diag, sub, _, _ = random_well_conditioned_spd_btd((4,), 9, 3, rng=3)
    garbage = diag + np.triu(np.random.normal(size=diag.shape), 1)
    c0 = S(tt(diag), tt(sub)).cholesky
    c1 = S(tt(garbage), tt(sub)).cholesky
    assert torch.equal(c0.block_diagonal, c1.block_diagonal)
    assert torch.equal(c0.block_sub_diagonal, c1.block_sub_diagonal)
    # in-place through the C ABI (out aliases in), as config 4 needs on one GPU
    from markovflow_b200 import _lib

    dg, sb = tt(diag).clone(), tt(sub).clone()
    info = torch.zeros(4, dtype=torch.int32, device=dev())
    st = _lib.lib().mf_btd_cholesky(
        _lib.MF_F64, _lib.ptr(dg), _lib.ptr(sb), None, _lib.ptr(dg), _lib.ptr(sb), None, None,
        _lib.ptr(info), _lib.i64(4), _lib.i64(9), _lib.i64(3), _lib.current_stream(),
    )
    assert st == 0
    torch.cuda.synchronize()
    assert torch.equal(dg, c0.block_diagonal) and torch.equal(sb, c0.block_sub_diagonal)


def test_cholesky_failure_raises_and_reports_block():
    from markovflow_b200 import CholeskyError

    _, S = _mods()
    diag, sub, _, _ = random_well_conditioned_spd_btd((3,), 6, 2, rng=5)
    diag[1, 4] = -np.eye(2)
    with pytest.raises(CholeskyError, match="chain 1, block 5"):
        S(tt(diag), tt(sub)).cholesky


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("transpose_left", [True, False])
@pytest.mark.parametrize("t", OUTER)
def test_solve(batch_shape, with_sub, d, transpose_left, t):
    if _skip(with_sub, t):
        return
    L, _ = _mods()
    dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    right = np.random.normal(size=batch_shape + (t, d))
    got = npy(L(tt(diag), tt(sub)).solve(tt(right), transpose_left=transpose_left))
    want = O.btd_solve(diag, sub, right, transpose_left=transpose_left)
    # the reference's generator (tests/unit/test_block_tri_diag.py:228-312: N(0,1) diagonal entries) is
    # ill-conditioned: where 1e-10 from the float64 oracle is out of reach, the long-double substitution
    # decides (CUDA result at least as close to it as the restated reference); residual checked tightly
    assert_parity(got, want, 1e-10, what="solve, reference generator",
                  truth=lambda: O.btd_solve(*as_ld(diag, sub, right), transpose_left=transpose_left))
    es = "...ji,...j->...i" if transpose_left else "...ij,...j->...i"
    resid = np.einsum(es, dense, got.reshape(batch_shape + (t * d,))) - right.reshape(batch_shape + (t * d,))
    assert np.max(np.abs(resid)) < 1e-9 * max(1.0, np.max(np.abs(got)))


@pytest.mark.parametrize("transpose_left", [True, False])
@pytest.mark.parametrize("d", [1, 2, 3, 5, 8])
def test_solve_vs_oracle_tight(transpose_left, d):
    L, _ = _mods()
    _, _, ld, ls = random_well_conditioned_spd_btd((6,), 41, d, rng=d)
    right = np.random.default_rng(2).standard_normal((6, 41, d))
    got = npy(L(tt(ld), tt(ls)).solve(tt(right), transpose_left=transpose_left))
    assert max_rel_err(got, O.btd_solve(ld, ls, right, transpose_left=transpose_left)) < 1e-10


def test_solve_broadcasts_leading_sample_dims_and_unit_batch():
    L, _ = _mods()
    _, _, ld, ls = random_well_conditioned_spd_btd((3,), 7, 2, rng=9)
    right = np.random.normal(size=(5, 3, 7, 2))
    got = npy(L(tt(ld), tt(ls)).solve(tt(right)))
    assert got.shape == (5, 3, 7, 2)
    assert max_rel_err(got, O.btd_solve(ld, ls, right)) < 1e-10
    # matrix batch (2,1) against right batch (2,4): the matrix's unit dim broadcasts
    _, _, ld2, ls2 = random_well_conditioned_spd_btd((2, 1), 5, 3, rng=10)
    right2 = np.random.normal(size=(2, 4, 5, 3))
    got2 = npy(L(tt(ld2), tt(ls2)).solve(tt(right2), transpose_left=True))
    assert max_rel_err(got2, O.btd_solve(ld2, ls2, right2, transpose_left=True)) < 1e-10
    with pytest.raises(ValueError):
        L(tt(ld), tt(ls)).solve(tt(np.zeros((3, 7, 3))))


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
@pytest.mark.parametrize("symmetrise", [True, False])
@pytest.mark.parametrize("transpose_left", [True, False])
def test_dense_mult(batch_shape, with_sub, d, transpose_left, symmetrise, t):
    if _skip(with_sub, t) or (transpose_left and symmetrise):
        return
    L, S = _mods()
    if symmetrise:
        dense, diag, sub = random_spd_btd(batch_shape, t, d, with_sub)
        m = S(tt(diag), tt(sub))
    else:
        dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
        m = L(tt(diag), tt(sub))
    right = np.random.normal(size=batch_shape + (t, d))
    got = npy(m.dense_mult(tt(right), transpose_left=transpose_left))
    es = "...ji,...j->...i" if transpose_left else "...ij,...j->...i"
    want = np.einsum(es, dense, right.reshape(batch_shape + (t * d,))).reshape(batch_shape + (t, d))
    assert max_rel_err(got, want) < 1e-12


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_diagonal_of_inverse(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    L, _ = _mods()
    dense, _, _ = random_spd_btd(batch_shape, t, d, with_sub)
    ld, ls = blocks_from_dense(np.linalg.cholesky(dense), d, with_sub)
    got = npy(L(tt(ld), tt(ls)).block_diagonal_of_inverse())
    want, _ = blocks_from_dense(np.linalg.inv(dense), d, False)
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-8 * np.max(np.abs(want)))  # reference's own check
    # ill-conditioned generator (M = L Lᵀ with N(1,1) diagonal entries): long double decides past 1e-10
    assert_parity(got, O.btd_inverse_subset(ld, ls)[0], 1e-10, what="block_diagonal_of_inverse, reference generator",
                  truth=lambda: O.btd_inverse_subset(*as_ld(ld, ls))[0])


@pytest.mark.parametrize("d", [1, 2, 3, 4, 6, 8])
def test_inverse_subset_with_subdiag_vs_oracle_tight(d):
    L, _ = _mods()
    _, _, ld, ls = random_well_conditioned_spd_btd((5,), 29, d, rng=20 + d)
    got_d, got_s = L(tt(ld), tt(ls))._inverse_subset(True)
    o_d, o_s = O.btd_inverse_subset(ld, ls, want_sub=True)
    assert max_rel_err(npy(got_d), o_d) < 1e-10
    assert max_rel_err(npy(got_s), o_s) < 1e-10
    # returned diagonal blocks are exactly symmetric (band_to_block symmetric=True)
    assert torch.equal(got_d, got_d.transpose(-1, -2))


@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", [3, 5])
def test_upper_diagonal_lower(batch_shape, d, t):
    _, S = _mods()
    dense, diag, sub = random_spd_btd(batch_shape, t, d, True)
    lower_m, diag_m = S(tt(diag), tt(sub)).upper_diagonal_lower()
    lower, dd = npy(lower_m.to_dense()), npy(diag_m.to_dense())
    np.testing.assert_allclose(lower, np.tril(lower))
    assert diag_m.block_sub_diagonal is None
    chol_d_u = np.swapaxes(dd, -1, -2) @ lower
    np.testing.assert_allclose(dense, np.swapaxes(chol_d_u, -1, -2) @ chol_d_u, rtol=1e-6)


@pytest.mark.parametrize("d", [1, 2, 3, 5, 8])
def test_upper_diagonal_lower_vs_oracle_tight(d):
    _, S = _mods()
    diag, sub, _, _ = random_well_conditioned_spd_btd((4,), 23, d, rng=40 + d)
    lower_m, diag_m = S(tt(diag), tt(sub)).upper_diagonal_lower()
    o_u, o_cd = O.btd_upper_diagonal_lower(diag, sub)
    assert max_rel_err(npy(lower_m.block_sub_diagonal), o_u) < 1e-10
    assert max_rel_err(npy(diag_m.block_diagonal), o_cd) < 1e-10


def test_constructor_shape_errors():
    L, S = _mods()
    with pytest.raises(ValueError):
        S(tt(np.zeros((3, 2, 3))))
    with pytest.raises(ValueError):
        S(tt(np.zeros((1, 2, 2))), tt(np.zeros((1, 2, 2))))  # sub-diagonal with outer_dim 1
    with pytest.raises(ValueError):
        L(tt(np.zeros((4, 2, 2))), tt(np.zeros((4, 2, 2))))  # wrong sub-diagonal length


def test_config2_shape_slice_matern52_posterior_precision():
    """BASELINE config 2 at reduced B,T: Matern52 D=3 posterior precision Cholesky + solve."""
    from markovflow_b200 import SymmetricBlockTriDiagonal

    rng = np.random.default_rng(71892305)
    b, t = 16, 400
    diags, subs = [], []
    for _ in range(b):
        ell, var = rng.uniform(0.5, 2.0), rng.uniform(0.5, 2.0)
        tp = np.cumsum(ell * rng.uniform(0.2, 1.0, size=t))
        k = O.Matern52(ell, var)
        ssm = k.state_space_model(tp)
        dg, sb = O.kalman_k_inv_post(ssm, k.emission_matrix(tp), np.array([[100.0]]))
        diags.append(dg)
        subs.append(sb)
    diag, sub = np.stack(diags), np.stack(subs)
    rhs = rng.standard_normal((b, t, 3))
    chol, x, _ = SymmetricBlockTriDiagonal(tt(diag), tt(sub)).cholesky_and_solve(tt(rhs))
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    assert max_rel_err(npy(chol.block_diagonal), o_ld) < 1e-10
    assert max_rel_err(npy(chol.block_sub_diagonal), o_ls) < 1e-10
    assert max_rel_err(npy(x), O.btd_solve(o_ld, o_ls, rhs)) < 1e-10


# ------------------------------------------------------------------------------------------------
# few long chains: the factorisation is evaluated parallel in time (btd_pit.cuh) -- segment elements
# of the linear-fractional map (S, r) -> (S', r'), seeds, seeded sweeps.  Must equal the oracle and
# the sequential sweep (tuning knob 2 = 1).
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
@pytest.mark.parametrize("b,t", [(1, 2000), (3, 301), (2, 130), (7, 640)])
def test_cholesky_parallel_in_time(b, t, d, dtype):
    from markovflow_b200 import _lib

    _, S = _mods()
    diag, sub, _, _ = random_well_conditioned_spd_btd((b,), t, d, rng=31 * d + t)
    rhs = np.random.default_rng(t).standard_normal((b, t, d))
    if dtype == torch.float32:
        diag, sub, rhs = (a.astype(np.float32).astype(np.float64) for a in (diag, sub, rhs))
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    o_x = O.btd_solve(o_ld, o_ls, rhs)
    tol = TOL[dtype]
    lib = _lib.lib()
    out = {}
    for knob in (0, 1):
        lib.mf_set_tuning(2, knob)
        try:
            m = S(tt(diag, dtype), tt(sub, dtype))
            chol, x, logdet = m.cholesky_and_solve(tt(rhs, dtype), want_log_det=True)
            plain = m.cholesky
        finally:
            lib.mf_set_tuning(2, 0)
        assert max_rel_err(npy(chol.block_diagonal), o_ld) < tol
        assert max_rel_err(npy(chol.block_sub_diagonal), o_ls) < tol
        assert max_rel_err(npy(x), o_x) < tol
        assert max_rel_err(npy(logdet), O.btd_abs_log_det(o_ld)) < tol
        assert float(torch.triu(chol.block_diagonal, 1).abs().max()) == 0.0
        assert max_rel_err(npy(plain.block_diagonal), o_ld) < tol
        out[knob] = (npy(chol.block_diagonal), npy(chol.block_sub_diagonal), npy(x))
    for a, b_ in zip(out[0], out[1]):
        assert max_rel_err(a, b_) < tol


def test_cholesky_parallel_in_time_segments_alias_and_failure():
    from markovflow_b200 import CholeskyError, _lib

    _, S = _mods()
    t, d, b = 400, 3, 2
    diag, sub, _, _ = random_well_conditioned_spd_btd((b,), t, d, rng=9)
    rhs = np.random.default_rng(1).standard_normal((b, t, d))
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    o_x = O.btd_solve(o_ld, o_ls, rhs)
    lib = _lib.lib()
    for seg in (2, 3, 7, 64, 133, 399):  # knob 3: steps per segment, ragged last segments included
        lib.mf_set_tuning(3, seg)
        try:
            chol, x, _ = S(tt(diag), tt(sub)).cholesky_and_solve(tt(rhs))
        finally:
            lib.mf_set_tuning(3, 0)
        assert max_rel_err(npy(chol.block_diagonal), o_ld) < 1e-10
        assert max_rel_err(npy(chol.block_sub_diagonal), o_ls) < 1e-10
        assert max_rel_err(npy(x), o_x) < 1e-10
    # in place (outputs alias inputs): the output slots cannot serve as scratch -> sequential sweep
    dg, sb = tt(diag).clone(), tt(sub).clone()
    info = torch.zeros(b, dtype=torch.int32, device=dev())
    st = lib.mf_btd_cholesky(_lib.MF_F64, _lib.ptr(dg), _lib.ptr(sb), None, _lib.ptr(dg), _lib.ptr(sb), None,
                             None, _lib.ptr(info), _lib.i64(b), _lib.i64(t), _lib.i64(d), _lib.current_stream())
    assert st == 0
    torch.cuda.synchronize()
    assert max_rel_err(npy(dg), o_ld) < 1e-10 and max_rel_err(npy(sb), o_ls) < 1e-10
    # a block that is not positive definite is reported for its chain
    bad = diag.copy()
    bad[1, 250] = -np.eye(d)
    with pytest.raises(CholeskyError, match="chain 1"):
        S(tt(bad), tt(sub)).cholesky


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("transpose_left", [False, True])
@pytest.mark.parametrize("unit", [False, True])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_solve_parallel_in_time(d, unit, transpose_left, dtype):
    """Few long chains: the triangular solve (an affine recursion) is evaluated parallel in time;
    equal to the oracle and to the sequential sweep, with leading sample dimensions broadcasting."""
    from markovflow_b200 import _lib

    L, _ = _mods()
    lib = _lib.lib()
    # (3, 700, 7): 100 segments per chain -> warp-scan fold of the elements
    for b, t, seg in ((1, 1500, 0), (2, 301, 0), (3, 400, 7), (3, 700, 7), (2, 131, 65)):
        _, _, ld, ls = random_well_conditioned_spd_btd((b,), t, d, rng=17 * d + t)
        if unit:
            ld = np.broadcast_to(np.eye(d), ld.shape).copy()
        rhs = np.random.default_rng(t).standard_normal((2, b, t, d))
        if dtype == torch.float32:
            ld, ls, rhs = (a.astype(np.float32).astype(np.float64) for a in (ld, ls, rhs))
        want = O.btd_solve(ld, ls, rhs, transpose_left=transpose_left)
        got = {}
        for knob in (0, 1):
            lib.mf_set_tuning(2, knob)
            lib.mf_set_tuning(3, seg)
            try:
                low = L(tt(ld, dtype), tt(ls, dtype), unit_diagonal=unit)
                got[knob] = npy(low.solve(tt(rhs, dtype), transpose_left=transpose_left))
            finally:
                lib.mf_set_tuning(2, 0)
                lib.mf_set_tuning(3, 0)
            assert max_rel_err(got[knob], want) < TOL[dtype]
        assert max_rel_err(got[0], got[1]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_upper_diagonal_lower_parallel_in_time(d, dtype):
    """Few long chains: U D U^T through the linear-fractional segment elements; equal to the oracle
    and to the sequential sweep; a non-positive-definite block still raises."""
    from markovflow_b200 import CholeskyError, _lib

    _, S = _mods()
    lib = _lib.lib()
    # (3, 400, 3) and (2, 129, 2): more than 64 segments per chain -> warp-scan fold of the elements
    for b, t, seg in ((1, 1500, 0), (2, 301, 0), (3, 400, 7), (3, 400, 3), (2, 131, 65), (2, 129, 2)):
        diag, sub, _, _ = random_well_conditioned_spd_btd((b,), t, d, rng=23 * d + t)
        if dtype == torch.float32:
            diag, sub = (a.astype(np.float32).astype(np.float64) for a in (diag, sub))
        o_u, o_cd = O.btd_upper_diagonal_lower(diag, sub)
        got = {}
        for knob in (0, 1):
            lib.mf_set_tuning(2, knob)
            lib.mf_set_tuning(3, seg)
            try:
                lower_m, diag_m = S(tt(diag, dtype), tt(sub, dtype)).upper_diagonal_lower()
            finally:
                lib.mf_set_tuning(2, 0)
                lib.mf_set_tuning(3, 0)
            got[knob] = (npy(lower_m.block_sub_diagonal), npy(diag_m.block_diagonal))
            assert max_rel_err(got[knob][0], o_u) < TOL[dtype]
            assert max_rel_err(got[knob][1], o_cd) < TOL[dtype]
            assert float(torch.triu(diag_m.block_diagonal, 1).abs().max()) == 0.0
        assert max_rel_err(got[0][0], got[1][0]) < TOL[dtype]
    diag, sub, _, _ = random_well_conditioned_spd_btd((2,), 500, d, rng=3)
    diag[1, 333] = -np.eye(d)
    with pytest.raises(CholeskyError, match="chain 1"):
        S(tt(diag), tt(sub)).upper_diagonal_lower()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_inverse_subset_parallel_in_time(d, dtype):
    """Few long chains: the sparse inverse subset (affine in Sigma) evaluated parallel in time; equal
    to the oracle and to the sequential sweep, with and without the sub-diagonal blocks."""
    from markovflow_b200 import _lib

    L, _ = _mods()
    lib = _lib.lib()
    for b, t, seg in ((1, 1500, 0), (2, 301, 0), (3, 400, 7), (2, 131, 65), (2, 129, 2)):
        _, _, ld, ls = random_well_conditioned_spd_btd((b,), t, d, rng=29 * d + t)
        if dtype == torch.float32:
            ld, ls = (a.astype(np.float32).astype(np.float64) for a in (ld, ls))
        o_d, o_s = O.btd_inverse_subset(ld, ls, want_sub=True)
        got = {}
        for knob in (0, 1):
            lib.mf_set_tuning(2, knob)
            lib.mf_set_tuning(3, seg)
            try:
                low = L(tt(ld, dtype), tt(ls, dtype))
                gd, gs = low._inverse_subset(True)
                only_d = low.block_diagonal_of_inverse()
            finally:
                lib.mf_set_tuning(2, 0)
                lib.mf_set_tuning(3, 0)
            got[knob] = (npy(gd), npy(gs))
            assert max_rel_err(got[knob][0], o_d) < TOL[dtype]
            assert max_rel_err(got[knob][1], o_s) < TOL[dtype]
            assert max_rel_err(npy(only_d), o_d) < TOL[dtype]
        assert max_rel_err(got[0][0], got[1][0]) < TOL[dtype]


def _block_cholesky_longdouble(diag, sub):
    """Block Cholesky in np.longdouble: a reference that separates the error of the two float64 paths
    from the conditioning of the input."""
    diag, sub = diag.astype(np.longdouble), sub.astype(np.longdouble)
    t, d = diag.shape[0], diag.shape[-1]
    ld, ls = np.zeros_like(diag), np.zeros_like(sub)
    prev = None
    for k in range(t):
        s = diag[k].copy()
        if prev is not None:
            s -= prev @ prev.T
        low = np.zeros((d, d), dtype=np.longdouble)
        for j in range(d):
            low[j, j] = np.sqrt(s[j, j] - (low[j, :j] ** 2).sum())
            for i in range(j + 1, d):
                low[i, j] = (s[i, j] - (low[i, :j] * low[j, :j]).sum()) / low[j, j]
        ld[k] = low
        if k + 1 < t:
            x = np.zeros((d, d), dtype=np.longdouble)
            for j in range(d):
                x[:, j] = (sub[k][:, j] - x[:, :j] @ low[j, :j]) / low[j, j]
            ls[k] = x
            prev = x
    return ld, ls


@pytest.mark.parametrize("ratio", [0.2, 0.02])
def test_parallel_in_time_is_as_accurate_as_the_sequential_sweep_on_ill_conditioned_input(ratio):
    """Matern52 posterior precision of ONE series with dt / lengthscale ~ ratio (block condition
    numbers 4e5 / 4e9): against a long-double factorisation the parallel-in-time path must not lose
    accuracy relative to the sequential sweep (tools/pit_conditioning.py has the wider table)."""
    from markovflow_b200 import _lib

    _, S = _mods()
    rng = np.random.default_rng(3)
    t = 1500
    tp = np.cumsum(ratio * rng.uniform(0.5, 1.5, size=t))
    k = O.Matern52(1.0, 1.0)
    diag, sub = O.kalman_k_inv_post(k.state_space_model(tp), k.emission_matrix(tp), np.array([[100.0]]))
    ref_ld, ref_ls = _block_cholesky_longdouble(diag, sub)
    lib = _lib.lib()
    err = {}
    for knob in (0, 1):
        lib.mf_set_tuning(2, knob)
        try:
            c = S(tt(diag[None]), tt(sub[None])).cholesky
        finally:
            lib.mf_set_tuning(2, 0)
        err[knob] = max(max_rel_err(npy(c.block_diagonal[0]), ref_ld.astype(np.float64)),
                        max_rel_err(npy(c.block_sub_diagonal[0]), ref_ls.astype(np.float64)))
    assert err[0] < 3.0 * err[1] + 1e-14, err
    assert err[0] < 1e-10
