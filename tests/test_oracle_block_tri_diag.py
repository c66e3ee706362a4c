"""Pin the oracle's block-tridiagonal restatement against dense ``numpy.linalg`` -- the same
identities the reference's ``tests/unit/test_block_tri_diag.py:46-225`` uses -- against the scalar
banded restatement of ``banded_matrices`` and against LAPACK's banded Cholesky (scipy)."""
import numpy as np
import pytest
import scipy.linalg

from oracle import np_oracle as O
from tests.helpers import (
    blocks_from_dense,
    dense_from_blocks,
    random_lower_btd,
    random_spd_btd,
    random_well_conditioned_spd_btd,
)

INNER = [1, 3]
OUTER = [1, 4]


def _skip(with_sub, t):
    return with_sub and t == 1


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_to_dense(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    np.testing.assert_allclose(O.btd_to_dense(diag, sub, symmetric=False), dense)
    dense_s, diag_s, sub_s = random_spd_btd(batch_shape, t, d, with_sub)
    np.testing.assert_allclose(O.btd_to_dense(diag_s, sub_s, symmetric=True), dense_s)


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
def test_abs_log_det(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    np.testing.assert_allclose(O.btd_abs_log_det(diag), np.linalg.slogdet(dense)[1])


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER + [5])
@pytest.mark.parametrize("t", OUTER + [7])
def test_cholesky(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    if d > 3 or t > 4:  # beyond the reference's sizes its generator stops being numerically PD
        if not with_sub:
            return
        diag, sub, _, _ = random_well_conditioned_spd_btd(batch_shape, t, d, rng=d * 100 + t)
        dense = O.btd_to_dense(diag, sub, symmetric=True)
    else:
        dense, diag, sub = random_spd_btd(batch_shape, t, d, with_sub)
    ld, ls = O.btd_cholesky(diag, sub)
    # the reference's generator is ill-conditioned (diagonal entries ~N(1,1) can be ~0); it
    # uses rtol=1e-3 itself (tests/unit/test_block_tri_diag.py:97)
    l_dense = O.btd_to_dense(ld, ls, symmetric=False)
    if t <= 4:
        np.testing.assert_allclose(l_dense, np.linalg.cholesky(dense), rtol=1e-3, atol=1e-6)
    # backward error is conditioning-independent: L Lᵀ reproduces M to rounding
    recon = l_dense @ np.swapaxes(l_dense, -1, -2)
    np.testing.assert_allclose(recon, dense, rtol=0, atol=1e-12 * np.max(np.abs(dense)))


def test_cholesky_reads_lower_triangle_only():
    dense, diag, sub = random_spd_btd((2,), 5, 3, True)
    garbage = diag + np.triu(np.random.normal(size=diag.shape), 1)
    ld0, ls0 = O.btd_cholesky(diag, sub)
    ld1, ls1 = O.btd_cholesky(garbage, sub)
    np.testing.assert_array_equal(ld0, ld1)
    np.testing.assert_array_equal(ls0, ls1)


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("transpose_left", [True, False])
@pytest.mark.parametrize("t", OUTER)
def test_solve(batch_shape, with_sub, d, transpose_left, t):
    if _skip(with_sub, t):
        return
    dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    right = np.random.normal(size=batch_shape + (t, d))
    got = O.btd_solve(diag, sub, right, transpose_left=transpose_left)
    inv = np.linalg.inv(dense)
    es = "...ji,...j->...i" if transpose_left else "...ij,...j->...i"
    want = np.einsum(es, inv, right.reshape(batch_shape + (t * d,))).reshape(batch_shape + (t, d))
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-8)


def test_solve_broadcasts_sample_dims():
    dense, diag, sub = random_lower_btd((3,), 4, 2, True)
    right = np.random.normal(size=(5, 3, 4, 2))
    got = O.btd_solve(diag, sub, right)
    for s in range(5):
        np.testing.assert_allclose(got[s], O.btd_solve(diag, sub, right[s]))


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER)
@pytest.mark.parametrize("symmetrise", [True, False])
@pytest.mark.parametrize("transpose_left", [True, False])
def test_dense_mult(batch_shape, with_sub, d, transpose_left, symmetrise, t):
    if _skip(with_sub, t) or (transpose_left and symmetrise):
        return
    if symmetrise:
        dense, diag, sub = random_spd_btd(batch_shape, t, d, with_sub)
    else:
        dense, diag, sub = random_lower_btd(batch_shape, t, d, with_sub)
    right = np.random.normal(size=batch_shape + (t, d))
    got = O.btd_dense_mult(diag, sub, right, transpose_left=transpose_left, symmetric=symmetrise)
    es = "...ji,...j->...i" if transpose_left else "...ij,...j->...i"
    want = np.einsum(es, dense, right.reshape(batch_shape + (t * d,))).reshape(batch_shape + (t, d))
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("with_sub", [True, False])
@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", OUTER + [6])
def test_diagonal_of_inverse(batch_shape, with_sub, d, t):
    if _skip(with_sub, t):
        return
    dense, _, _ = random_spd_btd(batch_shape, t, d, with_sub)
    ld, ls = blocks_from_dense(np.linalg.cholesky(dense), d, with_sub)
    sig, sig_sub = O.btd_inverse_subset(ld, ls, want_sub=True)
    want_d, want_s = blocks_from_dense(np.linalg.inv(dense), d, with_sub)
    scale = np.max(np.abs(want_d))
    np.testing.assert_allclose(sig, want_d, rtol=1e-3, atol=1e-8 * scale)
    if with_sub:
        np.testing.assert_allclose(sig_sub, want_s, rtol=1e-3, atol=1e-8 * scale)


@pytest.mark.parametrize("d", INNER)
@pytest.mark.parametrize("t", [3, 5])
def test_upper_diagonal_lower(batch_shape, d, t):
    dense, diag, sub = random_spd_btd(batch_shape, t, d, True)
    u_s, chol_d = O.btd_upper_diagonal_lower(diag, sub)
    eye = np.broadcast_to(np.eye(d), chol_d.shape)
    lower = O.btd_to_dense(eye, u_s, symmetric=False)
    dd = O.btd_to_dense(chol_d, None, symmetric=False)
    chol_d_u = np.swapaxes(dd, -1, -2) @ lower
    np.testing.assert_allclose(dense, np.swapaxes(chol_d_u, -1, -2) @ chol_d_u, rtol=1e-6)


# ---- scalar banded restatement (the third-party algorithm) vs the block recurrences ------------


@pytest.mark.parametrize("d,t", [(1, 5), (2, 4), (3, 6)])
def test_band_layout_is_lapack_lower_band(d, t):
    dense, diag, sub = random_spd_btd((), t, d, True)
    band = O.blocks_to_band(diag, sub)
    assert band.shape == (2 * d, t * d)
    for r in range(2 * d):
        for j in range(t * d - r):
            assert band[r, j] == dense[j + r, j]
    np.testing.assert_allclose(O.unpack_banded_matrix_to_dense(band), np.tril(dense))
    # LAPACK dpbtrf through scipy uses the identical layout
    chol_band_lapack = scipy.linalg.cholesky_banded(band, lower=True)
    np.testing.assert_allclose(O.cholesky_band(band), chol_band_lapack, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("d,t", [(1, 4), (2, 5), (3, 4)])
def test_banded_ops_equal_block_recurrences(d, t):
    dense, diag, sub = random_spd_btd((), t, d, True)
    band = O.blocks_to_band(diag, sub)
    l_band = O.cholesky_band(band)
    ld_b, ls_b = O.band_to_blocks(l_band, d)
    ld, ls = O.btd_cholesky(diag, sub)
    np.testing.assert_allclose(ld_b, ld, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ls_b, ls, rtol=1e-10, atol=1e-12)

    rhs = np.random.normal(size=(t, d))
    for tr in (False, True):
        x_band = O.solve_triang_mat(l_band, rhs.reshape(t * d, 1), transpose_left=tr)
        np.testing.assert_allclose(
            x_band.reshape(t, d), O.btd_solve(ld, ls, rhs, transpose_left=tr), rtol=1e-9, atol=1e-12
        )

    inv_band = O.inverse_from_cholesky_band(l_band)
    blk = O.band_to_block(inv_band, d)  # [2D, T*D]
    cov_blocks = blk.T.reshape(t, d, 2 * d)  # ssm_gaussian_transformations.py:447-458
    sig, sig_sub = O.btd_inverse_subset(ld, ls, want_sub=True)
    np.testing.assert_allclose(cov_blocks[..., :d], sig, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(
        np.swapaxes(cov_blocks[:-1, :, d:], -1, -2), sig_sub, rtol=1e-8, atol=1e-11
    )
    # block_diagonal_of_inverse slices the first D band rows (block_tri_diag.py:330-337)
    blk_d = O.band_to_block(inv_band[:d], d)
    np.testing.assert_allclose(blk_d.T.reshape(t, d, d), sig, rtol=1e-8, atol=1e-11)

    y = O.product_band_mat(l_band, rhs.reshape(t * d, 1)).reshape(t, d)
    np.testing.assert_allclose(y, O.btd_dense_mult(ld, ls, rhs), rtol=1e-10, atol=1e-12)
    y = O.product_band_mat(band, rhs.reshape(t * d, 1), symmetrise_left=True).reshape(t, d)
    np.testing.assert_allclose(
        y, O.btd_dense_mult(diag, sub, rhs, symmetric=True), rtol=1e-10, atol=1e-12
    )


def test_block_band_round_trip_no_subdiag():
    _, diag, _ = random_lower_btd((), 4, 3, False)
    band = O.blocks_to_band(diag, None)
    assert band.shape == (3, 12)
    d2, s2 = O.band_to_blocks(band, 3)
    assert s2 is None
    np.testing.assert_array_equal(d2, np.tril(diag))


def test_cholesky_band_failure_raises():
    band = np.array([[1.0, -1.0, 1.0], [0.0, 0.0, 0.0]])
    with pytest.raises(np.linalg.LinAlgError):
        O.cholesky_band(band)
