"""Numpy float64 restatement of Markovflow's structured linear-algebra hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the product package never
imports this file.

What is restated, and from where (paths relative to the reference checkout):

* The banded ops of the un-vendored third-party dependency ``banded-matrices==0.0.6``
  (``poetry.lock:133-144``), whose source is NOT in the reference tree.  Their
  published algorithm (LAPACK ``dpbtf2``-style scalar lower-band Cholesky, banded
  substitution, Takahashi sparse-inverse subset) is restated in the ``band_*`` /
  ``*_band`` functions below; parity is anchored on the reference's own call sites
  (``markovflow/block_tri_diag.py:158,189,233,330-331,350,436,558``,
  ``markovflow/ssm_gaussian_transformations.py:443-444,473,484``) and on the dense
  ``numpy.linalg`` identities its tests pin them with
  (``tests/unit/test_block_tri_diag.py:46-225``).
* The block-level recurrences those banded ops amount to on block-tridiagonal
  matrices (``btd_*``) -- this is the form the CUDA kernels implement.
* ``StateSpaceModel`` (``markovflow/state_space_model.py:231-609``),
  ``BaseKalmanFilter`` (``markovflow/kalman_filter.py:85-271``) and the
  natural/expectation transforms (``markovflow/ssm_gaussian_transformations.py``).
* The closed-form SDE kernels used to generate benchmark inputs
  (``markovflow/kernels/matern.py:299-324,434-501``,
  ``markovflow/kernels/periodic.py:103-187``, ``markovflow/kernels/sde_kernel.py:153-171,421-446,592-687``).

Pinning: ``tests/test_oracle_*.py`` check every function here against dense
``numpy.linalg`` (the reference tests' own oracle), against golden vectors produced
by the reference's ``tests/tools/numpy_kalman_filter.py`` run unmodified
(``tests/golden/make_golden.py``), and against ``scipy.linalg.cholesky_banded``
(LAPACK, identical band layout).  The TensorFlow reference itself cannot be
imported in this image (Python 3.12, no tensorflow / gpflow / banded_matrices).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

Array = np.ndarray

# ----------------------------------------------------------------------------------------------
# small batched dense helpers (loops over the tiny block dimension, vectorised over batch)
# ----------------------------------------------------------------------------------------------


def _t(x: Array) -> Array:
    return np.swapaxes(x, -1, -2)


def chol_lower(a: Array) -> Array:
    """Batched Cholesky reading ONLY the lower triangle (numpy and banded semantics)."""
    a = np.asarray(a)
    d = a.shape[-1]
    out = np.zeros_like(a)
    for j in range(d):
        s = a[..., j, j] - np.sum(out[..., j, :j] ** 2, axis=-1)
        ljj = np.sqrt(s)
        out[..., j, j] = ljj
        for i in range(j + 1, d):
            out[..., i, j] = (
                a[..., i, j] - np.sum(out[..., i, :j] * out[..., j, :j], axis=-1)
            ) / ljj
    return out


def tri_solve_lower(l: Array, b: Array) -> Array:
    """Solve ``L X = B`` for lower-triangular ``L[..., D, D]`` and ``B[..., D, M]`` (broadcasting)."""
    d = l.shape[-1]
    shape = np.broadcast_shapes(l.shape[:-2], b.shape[:-2]) + b.shape[-2:]
    x = np.zeros(shape, dtype=np.result_type(l, b))
    for i in range(d):
        acc = b[..., i, :] - np.einsum("...j,...jm->...m", l[..., i, :i], x[..., :i, :])
        x[..., i, :] = acc / l[..., i, i, None]
    return x


def tri_solve_lower_t(l: Array, b: Array) -> Array:
    """Solve ``Lᵀ X = B`` for lower-triangular ``L``."""
    d = l.shape[-1]
    shape = np.broadcast_shapes(l.shape[:-2], b.shape[:-2]) + b.shape[-2:]
    x = np.zeros(shape, dtype=np.result_type(l, b))
    for i in reversed(range(d)):
        acc = b[..., i, :] - np.einsum("...j,...jm->...m", l[..., i + 1 :, i], x[..., i + 1 :, :])
        x[..., i, :] = acc / l[..., i, i, None]
    return x


def cholesky_solve(l: Array, b: Array) -> Array:
    """``(L Lᵀ)⁻¹ B`` -- ``tf.linalg.cholesky_solve`` semantics."""
    return tri_solve_lower_t(l, tri_solve_lower(l, b))


def lu_solve(a: Array, b: Array) -> Array:
    """Batched ``a⁻¹ b`` by LU with partial pivoting (what ``tf.linalg.solve`` / ``np.linalg.solve`` do).
    float64 goes to LAPACK; other precisions (long double for the extended-precision adjudication of the
    parity tests, float32 for the same-arithmetic comparison) run the same elimination in loops."""
    a, b = np.asarray(a), np.asarray(b)
    dt = np.result_type(a.dtype, b.dtype)
    if dt == np.float64:
        return np.linalg.solve(a, b)
    shape = np.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    a = np.broadcast_to(a, shape + a.shape[-2:]).astype(dt, copy=True)
    x = np.broadcast_to(b, shape + b.shape[-2:]).astype(dt, copy=True)
    d = a.shape[-1]
    for j in range(d):
        piv = j + np.argmax(np.abs(a[..., j:, j]), axis=-1)
        rows_a = np.take_along_axis(a, piv[..., None, None], axis=-2)
        rows_x = np.take_along_axis(x, piv[..., None, None], axis=-2)
        np.put_along_axis(a, piv[..., None, None], a[..., j:j + 1, :], axis=-2)
        np.put_along_axis(x, piv[..., None, None], x[..., j:j + 1, :], axis=-2)
        a[..., j:j + 1, :], x[..., j:j + 1, :] = rows_a, rows_x
        for i in range(j + 1, d):
            f = a[..., i, j] / a[..., j, j]
            a[..., i, :] = a[..., i, :] - f[..., None] * a[..., j, :]
            x[..., i, :] = x[..., i, :] - f[..., None] * x[..., j, :]
    for j in range(d - 1, -1, -1):
        x[..., j, :] = (x[..., j, :] - np.sum(a[..., j, j + 1:, None] * x[..., j + 1:, :], axis=-2)) / a[..., j, j, None]
    return x


def spd_log_det(a: Array) -> Array:
    """``log det`` of symmetric positive-definite blocks, in the precision of the input."""
    a = np.asarray(a)
    if a.dtype == np.float64:
        return np.linalg.slogdet(a)[1]
    return 2.0 * np.sum(np.log(np.diagonal(chol_lower(a), axis1=-2, axis2=-1)), axis=-1)


def sym_from_lower(a: Array) -> Array:
    """Mirror the lower triangle (what ``band_to_block(symmetric=True)`` does to diagonal blocks)."""
    low = np.tril(a)
    return low + _t(np.tril(a, -1))


# ----------------------------------------------------------------------------------------------
# scalar lower-band restatement of banded_matrices (single matrix, pure loops: small cases only)
# band[r, j] = M[j + r, j]  (== LAPACK/scipy lower band storage ab[i-j, j] = a[i, j])
# ----------------------------------------------------------------------------------------------


def block_to_band(block: Array, block_size: int, symmetric: bool = True) -> Array:
    """``banded_matrices.banded.block_to_band`` as used at ``block_tri_diag.py:233``.

    ``block``: ``[nb*D, T*D]`` whose column ``j = k*D + c`` is the vertical stack
    ``[D_k[:, c]; A_k[:, c]]``.  Only entries on or below the main diagonal survive (the
    reference always calls it to obtain a *lower* band: ``left_upper_bandwidth=0`` at
    ``block_tri_diag.py:160,192``); ``symmetric`` therefore does not change the stored band.
    """
    rows, n = block.shape[-2:]
    d = block_size
    band = np.zeros(block.shape[:-2] + (rows, n), dtype=block.dtype)
    for j in range(n):
        c = j % d
        for r in range(rows):
            if c + r < rows:  # inside the stacked column (last block's padding is zero already)
                band[..., r, j] = block[..., c + r, j]
    return band


def band_to_block(band: Array, block_size: int, symmetric: bool = True) -> Array:
    """Inverse of :func:`block_to_band` (``block_tri_diag.py:330,558``).

    ``symmetric=True`` fills the strict upper triangle of each diagonal block from the lower
    one (needed for ``test_diagonal_of_inverse`` to compare against full dense blocks,
    ``tests/unit/test_block_tri_diag.py:199-202``); ``symmetric=False`` leaves it zero.
    """
    rows, n = band.shape[-2:]
    d = block_size
    block = np.zeros_like(band)
    for j in range(n):
        k, c = divmod(j, d)
        for i in range(rows):
            if i >= c:
                block[..., i, j] = band[..., i - c, j]
            elif symmetric:
                # M[kD+i, kD+c] = M[kD+c, kD+i] = band[c-i, kD+i]
                block[..., i, j] = band[..., c - i, k * d + i]
    return block


def unpack_banded_matrix_to_dense(band: Array) -> Array:
    """Lower band ``[K, N]`` -> dense lower-triangular ``[N, N]`` (``block_tri_diag.py:158``)."""
    k, n = band.shape
    dense = np.zeros((n, n), dtype=band.dtype)
    for r in range(k):
        for j in range(n - r):
            dense[j + r, j] = band[r, j]
    return dense


def pack_dense_to_band(dense: Array, lower_bandwidth: int) -> Array:
    n = dense.shape[-1]
    band = np.zeros((lower_bandwidth + 1, n), dtype=dense.dtype)
    for r in range(lower_bandwidth + 1):
        for j in range(n - r):
            band[r, j] = dense[j + r, j]
    return band


def cholesky_band(band: Array) -> Array:
    """Scalar lower-band Cholesky (``block_tri_diag.py:436``); column-by-column as in ``dpbtf2``."""
    kb = band.shape[0] - 1
    n = band.shape[1]
    l = np.array(band, dtype=np.float64, copy=True)
    for j in range(n):
        ajj = l[0, j]
        if not ajj > 0.0:
            raise np.linalg.LinAlgError("Banded Cholesky decomposition failure")
        ajj = np.sqrt(ajj)
        l[0, j] = ajj
        kn = min(kb, n - 1 - j)
        if kn > 0:
            l[1 : kn + 1, j] /= ajj
            # trailing update: A[j+a, j+b] -= L[j+a, j] L[j+b, j] for 1 <= b <= a <= kn
            for b in range(1, kn + 1):
                for a in range(b, kn + 1):
                    l[a - b, j + b] -= l[a, j] * l[b, j]
    return l


def solve_triang_mat(l_band: Array, rhs: Array, transpose_left: bool = False) -> Array:
    """Banded substitution ``L⁻¹ b`` / ``L⁻ᵀ b`` for ``rhs[N, M]`` (``block_tri_diag.py:350``)."""
    kb = l_band.shape[0] - 1
    n = l_band.shape[1]
    x = np.array(rhs, dtype=np.float64, copy=True)
    if not transpose_left:
        for i in range(n):
            for r in range(1, min(kb, i) + 1):
                x[i] -= l_band[r, i - r] * x[i - r]
            x[i] /= l_band[0, i]
    else:
        for i in reversed(range(n)):
            for r in range(1, min(kb, n - 1 - i) + 1):
                x[i] -= l_band[r, i] * x[i + r]
            x[i] /= l_band[0, i]
    return x


def inverse_from_cholesky_band(l_band: Array) -> Array:
    """Lower band of ``(L Lᵀ)⁻¹`` by the sparse-inverse-subset (Takahashi) recurrence
    (``block_tri_diag.py:331``, ``ssm_gaussian_transformations.py:444``)."""
    kb = l_band.shape[0] - 1
    n = l_band.shape[1]
    s = np.zeros((n, n))  # dense scratch; only the band is meaningful (small cases only)
    for j in reversed(range(n)):
        hi = min(n - 1, j + kb)
        ljj = l_band[0, j]
        for i in range(hi, j, -1):
            acc = 0.0
            for p in range(j + 1, hi + 1):
                sip = s[i, p] if i >= p else s[p, i]
                acc += sip * l_band[p - j, j]
            s[i, j] = -acc / ljj
        acc = 0.0
        for p in range(j + 1, hi + 1):
            acc += s[p, j] * l_band[p - j, j]
        s[j, j] = 1.0 / (ljj * ljj) - acc / ljj
    return pack_dense_to_band(s, kb)


def product_band_mat(
    band: Array, mat: Array, transpose_left: bool = False, symmetrise_left: bool = False
) -> Array:
    """``L x``, ``Lᵀ x`` or (symmetrised) ``M x`` (``block_tri_diag.py:189``)."""
    low = unpack_banded_matrix_to_dense(band)
    if symmetrise_left:
        low = low + low.T - np.diag(np.diag(low))
    elif transpose_left:
        low = low.T
    return low @ mat


def solve_triang_band(l_band: Array, r_band: Array, transpose_left: bool = False) -> Array:
    """``L⁻¹ R`` / ``L⁻ᵀ R`` truncated to R's lower band, ``R`` given by its lower band only
    (``ssm_gaussian_transformations.py:473``)."""
    low = unpack_banded_matrix_to_dense(l_band)
    r = unpack_banded_matrix_to_dense(r_band)  # tril(P): right_upper_bandwidth=0
    left = low.T if transpose_left else low
    x = lu_solve(left, r)
    return pack_dense_to_band(x, r_band.shape[0] - 1)


# ----------------------------------------------------------------------------------------------
# block layout <-> band layout (block_tri_diag.py:206-237, 549-592)
# ----------------------------------------------------------------------------------------------


def blocks_to_band(diag: Array, sub: Optional[Array]) -> Array:
    """``BlockTriDiagonal._convert_to_band`` for ONE matrix: ``[T,D,D]`` (+``[T-1,D,D]``) -> band."""
    t, d, _ = diag.shape
    if sub is None:
        concatted = _t(diag).reshape(t * d, d)
    else:
        padded = np.concatenate([sub, np.zeros((1, d, d), dtype=sub.dtype)], axis=0)
        concatted = np.concatenate([_t(diag), _t(padded)], axis=-1).reshape(t * d, 2 * d)
    return block_to_band(concatted.T, d)


def band_to_blocks(band: Array, d: int) -> Tuple[Array, Optional[Array]]:
    """``_banded_to_block_tri`` for ONE matrix (``block_tri_diag.py:549-592``)."""
    block = band_to_block(band, d, symmetric=False)
    nd = block.shape[0] // d
    t = block.shape[1] // d
    diags = _t(_t(block.reshape(nd, d, t * d)).reshape(nd, t, d, d))
    sub = diags[1, :-1] if nd == 2 else None
    return diags[0], sub


# ----------------------------------------------------------------------------------------------
# block-level recurrences (batched over leading dims): what the CUDA kernels implement
# ----------------------------------------------------------------------------------------------


def btd_to_dense(diag: Array, sub: Optional[Array], symmetric: bool) -> Array:
    """``BlockTriDiagonal.to_dense`` (``block_tri_diag.py:150-173``): upper triangles of the
    diagonal blocks are ignored; symmetric matrices are mirrored from the lower part."""
    *batch, t, d, _ = diag.shape
    dense = np.zeros(tuple(batch) + (t * d, t * d), dtype=diag.dtype)
    for k in range(t):
        dense[..., k * d : (k + 1) * d, k * d : (k + 1) * d] = np.tril(diag[..., k, :, :])
        if sub is not None and k < t - 1:
            dense[..., (k + 1) * d : (k + 2) * d, k * d : (k + 1) * d] = sub[..., k, :, :]
    if symmetric:
        dense = dense + _t(np.tril(dense, -1))
    return dense


def btd_cholesky(diag: Array, sub: Optional[Array]) -> Tuple[Array, Optional[Array]]:
    """``SymmetricBlockTriDiagonal.cholesky`` (``block_tri_diag.py:423-436``):
    ``Ld_0 = chol(D_0)``, ``Ls_k = A_k Ld_k⁻ᵀ``, ``Ld_{k+1} = chol(D_{k+1} − Ls_k Ls_kᵀ)``."""
    t = diag.shape[-3]
    ld = np.zeros_like(diag)
    ls = None if sub is None else np.zeros_like(sub)
    s = diag[..., 0, :, :]
    for k in range(t):
        ld[..., k, :, :] = chol_lower(s)
        if k < t - 1:
            if sub is not None:
                # Ls Ldᵀ = A  ->  Ld Lsᵀ = Aᵀ
                lsk = _t(tri_solve_lower(ld[..., k, :, :], _t(sub[..., k, :, :])))
                ls[..., k, :, :] = lsk
                s = diag[..., k + 1, :, :] - lsk @ _t(lsk)
            else:
                s = diag[..., k + 1, :, :]
    return ld, ls


def btd_solve(ld: Array, ls: Optional[Array], rhs: Array, transpose_left: bool = False) -> Array:
    """``LowerTriangularBlockTriDiagonal.solve`` (``block_tri_diag.py:339-351``).  Only the lower
    triangle of ``ld`` is read.  ``rhs[..., T, D]`` may carry extra leading (sample) dims."""
    t = ld.shape[-3]
    ld = np.tril(ld)
    shape = np.broadcast_shapes(ld.shape[:-3], rhs.shape[:-2]) + rhs.shape[-2:]
    x = np.zeros(shape, dtype=np.result_type(ld, rhs))
    if not transpose_left:
        for k in range(t):
            b = rhs[..., k, :]
            if k > 0 and ls is not None:
                b = b - np.einsum("...ij,...j->...i", ls[..., k - 1, :, :], x[..., k - 1, :])
            x[..., k, :] = tri_solve_lower(ld[..., k, :, :], b[..., None])[..., 0]
    else:
        for k in reversed(range(t)):
            b = rhs[..., k, :]
            if k < t - 1 and ls is not None:
                b = b - np.einsum("...ji,...j->...i", ls[..., k, :, :], x[..., k + 1, :])
            x[..., k, :] = tri_solve_lower_t(ld[..., k, :, :], b[..., None])[..., 0]
    return x


def btd_abs_log_det(ld: Array) -> Array:
    """``abs_log_det`` (``block_tri_diag.py:353-366``): ``½ Σ log(L_nn²)``."""
    dd = np.diagonal(ld, axis1=-2, axis2=-1)
    return 0.5 * np.sum(np.log(np.square(dd)), axis=(-1, -2))


def btd_inverse_subset(
    ld: Array, ls: Optional[Array], want_sub: bool = False
) -> Tuple[Array, Optional[Array]]:
    """Diagonal (and sub-diagonal) blocks of ``(L Lᵀ)⁻¹``
    (``block_diagonal_of_inverse``, ``block_tri_diag.py:318-337``; with the sub-diagonal as used by
    ``naturals_to_ssm_params``, ``ssm_gaussian_transformations.py:443-458``).

    Backward recurrence: ``Σ_{T-1,T-1} = (Ld Ldᵀ)⁻¹``; ``J_k = Ls_k Ld_k⁻¹``;
    ``Σ_{k+1,k} = −Σ_{k+1,k+1} J_k``; ``Σ_{kk} = (Ld_k Ld_kᵀ)⁻¹ − J_kᵀ Σ_{k+1,k}``.
    Returned diagonal blocks are full symmetric (``band_to_block(symmetric=True)``).
    """
    t, d = ld.shape[-3], ld.shape[-1]
    ld = np.tril(ld)
    eye = np.eye(d, dtype=ld.dtype)
    sig = np.zeros_like(ld)
    sig_sub = None
    if want_sub and ls is not None:
        sig_sub = np.zeros_like(ls)
    for k in reversed(range(t)):
        linv = tri_solve_lower(ld[..., k, :, :], np.broadcast_to(eye, ld[..., k, :, :].shape))
        local = _t(linv) @ linv
        if k < t - 1 and ls is not None:
            j = ls[..., k, :, :] @ linv
            s_sub = -sig[..., k + 1, :, :] @ j
            local = local - _t(j) @ s_sub
            if sig_sub is not None:
                sig_sub[..., k, :, :] = s_sub
        sig[..., k, :, :] = sym_from_lower(local)
    return sig, sig_sub


def btd_upper_diagonal_lower(diag: Array, sub: Array) -> Tuple[Array, Array]:
    """``SymmetricBlockTriDiagonal.upper_diagonal_lower`` (``block_tri_diag.py:438-545``).

    Returns ``(u_s, chol_d_s)``: ``Uᵀ`` has identity diagonal blocks and sub-diagonal ``u_s``;
    ``D`` is block diagonal with Cholesky factors ``chol_d_s``.
    """
    t = diag.shape[-3]
    chol_d = np.zeros_like(diag)
    u_s = np.zeros_like(sub)
    chol_d[..., t - 1, :, :] = chol_lower(diag[..., t - 1, :, :])
    for k in reversed(range(t - 1)):
        x = cholesky_solve(chol_d[..., k + 1, :, :], sub[..., k, :, :])
        u_s[..., k, :, :] = x
        dk = diag[..., k, :, :] - _t(sub[..., k, :, :]) @ x
        chol_d[..., k, :, :] = chol_lower(dk)
    return u_s, chol_d


def btd_dense_mult(
    diag: Array,
    sub: Optional[Array],
    right: Array,
    transpose_left: bool = False,
    symmetric: bool = False,
) -> Array:
    """``BlockTriDiagonal.dense_mult`` (``block_tri_diag.py:175-199``)."""
    t = diag.shape[-3]
    dl = np.tril(diag)
    dblk = sym_from_lower(diag) if symmetric else (_t(dl) if transpose_left else dl)
    y = np.einsum("...kij,...kj->...ki", dblk, right)
    if sub is not None and t > 1:
        lower_part = np.einsum("...kij,...kj->...ki", sub, right[..., :-1, :])  # rows 1..T-1
        upper_part = np.einsum("...kji,...kj->...ki", sub, right[..., 1:, :])  # rows 0..T-2
        if symmetric:
            y[..., 1:, :] += lower_part
            y[..., :-1, :] += upper_part
        elif transpose_left:
            y[..., :-1, :] += upper_part
        else:
            y[..., 1:, :] += lower_part
    return y


# ----------------------------------------------------------------------------------------------
# StateSpaceModel (state_space_model.py)
# ----------------------------------------------------------------------------------------------


class SSM:
    """Plain container with the reference's parameter layout (``state_space_model.py:77-122``)."""

    def __init__(self, mu0: Array, chol_p0: Array, a_s: Array, b_s: Array, chol_q_s: Array):
        self.mu0 = np.asarray(mu0)
        self.chol_p0 = np.asarray(chol_p0)
        self.a_s = np.asarray(a_s)
        self.b_s = np.asarray(b_s)
        self.chol_q_s = np.asarray(chol_q_s)

    @property
    def state_dim(self) -> int:
        return self.a_s.shape[-1]

    @property
    def num_transitions(self) -> int:
        return self.a_s.shape[-3]

    @property
    def batch_shape(self):
        return self.a_s.shape[:-3]

    @property
    def concatenated_state_offsets(self) -> Array:
        return np.concatenate([self.mu0[..., None, :], self.b_s], axis=-2)

    @property
    def concatenated_cholesky_process_covariance(self) -> Array:
        return np.concatenate([self.chol_p0[..., None, :, :], self.chol_q_s], axis=-3)


def ssm_build_precision(ssm: SSM) -> Tuple[Array, Array]:
    """``_build_precision`` (``state_space_model.py:431-483``): diag
    ``[P0⁻¹+A1ᵀQ1⁻¹A1, …, Qn⁻¹]``, sub ``−Q_k⁻¹A_k``."""
    d = ssm.state_dim
    inv_q_a = cholesky_solve(ssm.chol_q_s, ssm.a_s)
    aqa = _t(ssm.a_s) @ inv_q_a
    chols = ssm.concatenated_cholesky_process_covariance
    inv_q = cholesky_solve(chols, np.broadcast_to(np.eye(d, dtype=chols.dtype), chols.shape))
    diag = inv_q.copy()
    diag[..., :-1, :, :] += aqa
    return diag, -inv_q_a


def affine_recurrence(a_s: Array, c: Array) -> Array:
    """``a_inv_block.solve(c)`` (``state_space_model.py:251,278-296``):
    ``x_0 = c_0``, ``x_k = A_{k-1} x_{k-1} + c_k``; ``c`` may carry leading sample dims."""
    t = c.shape[-2]
    shape = np.broadcast_shapes(a_s.shape[:-3], c.shape[:-2]) + c.shape[-2:]
    x = np.zeros(shape, dtype=np.result_type(a_s, c))
    x[..., 0, :] = c[..., 0, :]
    for k in range(1, t):
        x[..., k, :] = np.einsum("...ij,...j->...i", a_s[..., k - 1, :, :], x[..., k - 1, :]) + c[..., k, :]
    return x


def ssm_marginal_means(ssm: SSM) -> Array:
    """``marginal_means`` (``state_space_model.py:231-251``)."""
    return affine_recurrence(ssm.a_s, ssm.concatenated_state_offsets)


def ssm_marginal_covariances(ssm: SSM) -> Array:
    """``marginal_covariances`` (``state_space_model.py:253-262``):
    ``precision.cholesky.block_diagonal_of_inverse()``."""
    ld, ls = btd_cholesky(*ssm_build_precision(ssm))
    return btd_inverse_subset(ld, ls)[0]


def ssm_subsequent_covariances(ssm: SSM, marginal_covs: Array) -> Array:
    """``subsequent_covariances`` (``state_space_model.py:326-341``): ``A_k Σ_kk``."""
    return ssm.a_s @ marginal_covs[..., :-1, :, :]


def ssm_log_det_precision(ssm: SSM) -> Array:
    """``log_det_precision`` (``state_space_model.py:343-373``)."""
    d0 = np.diagonal(ssm.chol_p0, axis1=-2, axis2=-1)
    dq = np.diagonal(ssm.chol_q_s, axis1=-2, axis2=-1)
    return -(np.sum(np.log(np.square(d0)), axis=-1) + np.sum(np.log(np.square(dq)), axis=(-1, -2)))


def _mvn_tril_log_prob(x: Array, loc: Array, scale_tril: Array) -> Array:
    d = x.shape[-1]
    z = tri_solve_lower(scale_tril, (x - loc)[..., None])[..., 0]
    half_log_det = np.sum(np.log(np.abs(np.diagonal(scale_tril, axis1=-2, axis2=-1))), axis=-1)
    return -0.5 * np.sum(z * z, axis=-1) - half_log_det - 0.5 * d * np.log(2.0 * np.pi)


def ssm_log_pdf(ssm: SSM, states: Array) -> Array:
    """``log_pdf`` (``state_space_model.py:485-526``)."""
    init = _mvn_tril_log_prob(states[..., 0, :], ssm.mu0, ssm.chol_p0)
    cond_means = np.einsum("...kij,...kj->...ki", ssm.a_s, states[..., :-1, :]) + ssm.b_s
    rest = _mvn_tril_log_prob(states[..., 1:, :], cond_means, ssm.chol_q_s)
    return init + np.sum(rest, axis=-1)


def ssm_sample_from_epsilons(ssm: SSM, epsilons: Array) -> Array:
    """``sample`` (``state_space_model.py:298-324``) with the N(0,I) draw supplied:
    ``epsilons[..., T, D]``."""
    z = np.einsum("...kij,...kj->...ki", ssm.concatenated_cholesky_process_covariance, epsilons)
    return affine_recurrence(ssm.a_s, ssm.concatenated_state_offsets + z)


def ssm_kl_divergence(q: SSM, p: SSM) -> Array:
    """``kl_divergence`` (``state_space_model.py:528-593``): ``KL(q ‖ p)``."""
    cov1 = ssm_marginal_covariances(q)
    pd, ps = ssm_build_precision(p)
    sub1 = ssm_subsequent_covariances(q, cov1)
    trace = np.sum(pd * cov1, axis=(-3, -2, -1)) + 2.0 * np.sum(ps * sub1, axis=(-3, -2, -1))
    mean_diff = ssm_marginal_means(p) - ssm_marginal_means(q)
    ld, ls = btd_cholesky(pd, ps)
    l_md = btd_dense_mult(ld, ls, mean_diff, transpose_left=True)
    mahalanobis = np.sum(l_md * l_md, axis=(-2, -1))
    dim = (q.num_transitions + 1) * q.state_dim
    return 0.5 * (trace + mahalanobis - dim - ssm_log_det_precision(p) + ssm_log_det_precision(q))


def cholesky_or_zero(cov: Array) -> Array:
    """``state_space_model_from_covariances.cholesky_or_zero`` (``state_space_model.py:634-656``)."""
    mask = np.all(cov == 0.0, axis=(-2, -1))
    eye = np.eye(cov.shape[-1])
    fixed = cov + mask[..., None, None] * eye
    chol = chol_lower(fixed)
    return np.where(mask[..., None, None], 0.0, chol)


def ssm_from_covariances(mu0, p0, a_s, b_s, q_s) -> SSM:
    return SSM(mu0, cholesky_or_zero(np.asarray(p0)), a_s, b_s, cholesky_or_zero(np.asarray(q_s)))


# ----------------------------------------------------------------------------------------------
# Kalman filter, SpInGP form (kalman_filter.py)
# ----------------------------------------------------------------------------------------------


def _r_inv_from_chol(chol_r: Array) -> Array:
    """``KalmanFilter._r_inv`` (``kalman_filter.py:341-348``)."""
    m = chol_r.shape[-1]
    return cholesky_solve(chol_r, np.eye(m, dtype=np.asarray(chol_r).dtype))


def kalman_k_inv_post(ssm: SSM, h: Array, r_inv: Array) -> Tuple[Array, Array]:
    """``_k_inv_post`` (``kalman_filter.py:85-101``): ``K⁻¹ + GᵀΣ⁻¹G``."""
    pd, ps = ssm_build_precision(ssm)
    hrh = np.einsum("...ji,...jk,...kl->...il", h, r_inv, h)
    return pd + hrh, ps


def kalman_back_project(h: Array, r_inv: Array, obs: Array) -> Array:
    """``_back_project_y_to_state`` (``kalman_filter.py:257-271``)."""
    back = np.einsum("...ij,...ki->...kj", h, r_inv)
    return np.einsum("...ij,...i->...j", back, obs)


def kalman_log_likelihood(
    ssm: SSM, h: Array, obs: Array, r_inv: Array, per_chain: bool = False
) -> Array:
    """``BaseKalmanFilter.log_likelihood`` (``kalman_filter.py:184-255``).

    ``r_inv`` is ``[m, m]`` (``KalmanFilter``) or ``[T, m, m]`` (``KalmanFilterWithSites``,
    ``kalman_filter.py:483-486``).  The reference returns the batch-summed scalar;
    ``per_chain=True`` returns the summands.
    """
    t = ssm.num_transitions + 1
    m = h.shape[-2]
    ld, ls = btd_cholesky(*kalman_k_inv_post(ssm, h, r_inv))
    marginal = np.einsum("...kij,...kj->...ki", h, ssm_marginal_means(ssm))
    disp = obs - marginal
    cst = -0.5 * np.log(2.0 * np.pi) * (m * t)
    term1 = -0.5 * np.sum(np.einsum("...op,...p,...o->...o", r_inv, disp, disp), axis=(-1, -2))
    obs_proj = kalman_back_project(h, r_inv, disp)
    term2 = 0.5 * np.sum(np.square(btd_solve(ld, ls, obs_proj)), axis=(-1, -2))
    if r_inv.ndim == 2:
        log_det_obs = t * spd_log_det(r_inv)
    else:
        log_det_obs = np.sum(spd_log_det(r_inv), axis=-1)
    term3 = 0.5 * ssm_log_det_precision(ssm) - btd_abs_log_det(ld) + 0.5 * log_det_obs
    out = cst + term1 + term2 + term3
    return out if per_chain else np.sum(out)


def kalman_posterior_ssm(ssm: SSM, h: Array, obs: Array, r_inv: Array) -> SSM:
    """``posterior_state_space_model`` (``kalman_filter.py:109-182``)."""
    d = ssm.state_dim
    post_d, post_s = kalman_k_inv_post(ssm, h, r_inv)
    u_s, chol_d = btd_upper_diagonal_lower(post_d, post_s)
    eye = np.broadcast_to(np.eye(d, dtype=chol_d.dtype), chol_d.shape)
    obs_proj = kalman_back_project(h, r_inv, obs)
    pd, ps = ssm_build_precision(ssm)
    k_inv_mu = btd_dense_mult(pd, ps, ssm_marginal_means(ssm), symmetric=True)
    tmp = btd_solve(eye, u_s, obs_proj + k_inv_mu, transpose_left=True)
    tmp = btd_solve(chol_d, None, tmp)
    m_post = btd_solve(chol_d, None, tmp, transpose_left=True)
    concatted_qs = chol_lower(cholesky_solve(np.tril(chol_d), eye))
    return SSM(
        m_post[..., 0, :],
        concatted_qs[..., 0, :, :],
        -u_s,
        m_post[..., 1:, :],
        concatted_qs[..., 1:, :, :],
    )


def sites_means_precisions(nat1: Array, nat2: Array) -> Tuple[Array, Array, Array]:
    """``UnivariateGaussianSitesNat`` (``kalman_filter.py:382-433``): means, precisions, log-dets."""
    return -0.5 * nat1 / nat2[..., 0], -2.0 * nat2, np.log(-2.0 * nat2)


def kalman_filter_time_varying(
    mu0, p0, a_s, b_s, q_s, h, r, obs
) -> Tuple[Array, Array, Array]:
    """Classical predict/update filter for ONE chain with per-step parameters: an independent
    check of :func:`kalman_log_likelihood` (same recursion as the reference's test oracle
    ``tests/tools/numpy_kalman_filter.py:66-137``, generalised to time-varying A,b,Q,H,R).

    Returns per-step log-likelihoods ``[T]``, filtered means ``[T,D]`` and covariances ``[T,D,D]``.
    """
    t = obs.shape[0]
    d = mu0.shape[-1]
    m = obs.shape[-1]
    lls = np.zeros(t)
    fm = np.zeros((t, d))
    fp = np.zeros((t, d, d))
    pm, pp = mu0, p0
    for k in range(t):
        hk = h[k]
        rk = r if r.ndim == 2 else r[k]
        v = obs[k] - hk @ pm
        s = hk @ pp @ hk.T + rk
        s_inv = np.linalg.inv(s)
        gain = pp @ hk.T @ s_inv
        fm[k] = pm + gain @ v
        fp[k] = (np.eye(d) - gain @ hk) @ pp
        lls[k] = -0.5 * (v @ s_inv @ v + m * np.log(2.0 * np.pi) + np.linalg.slogdet(s)[1])
        if k < t - 1:
            pm = a_s[k] @ fm[k] + b_s[k]
            pp = a_s[k] @ fp[k] @ a_s[k].T + q_s[k]
    return lls, fm, fp


# ----------------------------------------------------------------------------------------------
# parallel-in-time filtering elements (Särkkä & García-Fernández 2021)
# ----------------------------------------------------------------------------------------------


def pscan_elements(mu0, p0, a_s, b_s, q_s, h, r, obs):
    """Per-step scan elements ``(A, b, C, eta, J)`` for ONE chain (``[T, ...]`` arrays)."""
    t = obs.shape[0]
    d = mu0.shape[-1]
    el_a = np.zeros((t, d, d))
    el_b = np.zeros((t, d))
    el_c = np.zeros((t, d, d))
    el_eta = np.zeros((t, d))
    el_j = np.zeros((t, d, d))
    eye = np.eye(d)
    for k in range(t):
        hk = h[k]
        if k == 0:
            s = hk @ p0 @ hk.T + r
            gain = p0 @ hk.T @ np.linalg.inv(s)
            el_b[0] = mu0 + gain @ (obs[0] - hk @ mu0)
            el_c[0] = p0 - gain @ s @ gain.T
        else:
            f, u, q = a_s[k - 1], b_s[k - 1], q_s[k - 1]
            s = hk @ q @ hk.T + r
            s_inv = np.linalg.inv(s)
            gain = q @ hk.T @ s_inv
            el_a[k] = (eye - gain @ hk) @ f
            el_b[k] = u + gain @ (obs[k] - hk @ u)
            el_c[k] = (eye - gain @ hk) @ q
            el_eta[k] = f.T @ hk.T @ s_inv @ (obs[k] - hk @ u)
            el_j[k] = f.T @ hk.T @ s_inv @ hk @ f
    return el_a, el_b, el_c, el_eta, el_j


def pscan_combine(ei, ej):
    """Associative combine, ``ei`` earlier, ``ej`` later."""
    ai, bi, ci, etai, ji = ei
    aj, bj, cj, etaj, jj = ej
    d = ai.shape[-1]
    eye = np.eye(d)
    m = np.linalg.inv(eye + ci @ jj)
    n = np.linalg.inv(eye + jj @ ci)
    a = aj @ m @ ai
    b = aj @ m @ (bi + ci @ etaj) + bj
    c = aj @ m @ ci @ aj.T + cj
    eta = ai.T @ n @ (etaj - jj @ bi) + etai
    j = ai.T @ n @ jj @ ai + ji
    return a, b, c, eta, j


def pscan_log_likelihood(mu0, p0, a_s, b_s, q_s, h, r, obs, segment: int = 7) -> float:
    """Log-likelihood of ONE chain computed the parallel-in-time way: per-segment summaries,
    exclusive prefix of summaries, seeded local filtering (the structure of the CUDA scan)."""
    t = obs.shape[0]
    els = pscan_elements(mu0, p0, a_s, b_s, q_s, h, r, obs)
    starts = list(range(0, t, segment))
    summaries = []
    for s0 in starts:
        e = tuple(x[s0] for x in els)
        for k in range(s0 + 1, min(s0 + segment, t)):
            e = pscan_combine(e, tuple(x[k] for x in els))
        summaries.append(e)
    total = 0.0
    prefix = None
    m = obs.shape[-1]
    for si, s0 in enumerate(starts):
        s1 = min(s0 + segment, t)
        if prefix is None:
            pm, pp = mu0, p0
        else:
            fm_prev, fp_prev = prefix[1], prefix[2]
            pm = a_s[s0 - 1] @ fm_prev + b_s[s0 - 1]
            pp = a_s[s0 - 1] @ fp_prev @ a_s[s0 - 1].T + q_s[s0 - 1]
        for k in range(s0, s1):
            hk = h[k]
            v = obs[k] - hk @ pm
            s = hk @ pp @ hk.T + r
            s_inv = np.linalg.inv(s)
            gain = pp @ hk.T @ s_inv
            fm = pm + gain @ v
            fp = pp - gain @ s @ gain.T
            total += -0.5 * (v @ s_inv @ v + m * np.log(2.0 * np.pi) + np.linalg.slogdet(s)[1])
            if k < t - 1:
                pm = a_s[k] @ fm + b_s[k]
                pp = a_s[k] @ fp @ a_s[k].T + q_s[k]
        prefix = summaries[si] if prefix is None else pscan_combine(prefix, summaries[si])
    return total


def pscan_segment_element(a_in, b_in, q_in, h, r, obs, prior=None):
    """Range element WITH log-normaliser ``(A, b, C, eta, J, ell)`` of a run of steps of ONE chain:
    ``p(y_run | x_prev) = exp(ell - x'Jx/2 + eta'x)`` and ``x_last | x_prev, y_run ~ N(Ax+b, C)``.
    ``a_in[k], b_in[k], q_in[k]`` lead INTO step ``k``; with ``prior=(mu0, P0)`` step 0 starts from
    the prior instead and ``a_in`` has one entry fewer.  The ``ell`` of the join of all elements
    of a series is its marginal log-likelihood (extended)."""
    t = obs.shape[0]
    d = h.shape[-1]
    a, b, c = np.eye(d), np.zeros(d), np.zeros((d, d))
    eta, jm, ell = np.zeros(d), np.zeros((d, d)), 0.0
    off = 0 if prior is None else 1
    for k in range(t):
        if k == 0 and prior is not None:
            a, b, c = np.zeros((d, d)), np.asarray(prior[0], float), np.asarray(prior[1], float)
        else:
            f, u, q = a_in[k - off], b_in[k - off], q_in[k - off]
            a, b, c = f @ a, f @ b + u, f @ c @ f.T + q
        rk = r if r.ndim == 2 else r[k]
        w_mat = np.linalg.inv(np.linalg.cholesky(rk))  # whitener
        hw, yw = w_mat @ h[k], w_mat @ obs[k]
        ell += np.sum(np.log(np.diag(w_mat)))
        for i in range(hw.shape[0]):  # absorb one whitened scalar at a time
            hv, yv = hw[i], yw[i]
            g = c @ hv
            s = 1.0 + hv @ g
            gain = g / s
            w = a.T @ hv
            v = yv - hv @ b
            a, b, c = a - np.outer(gain, w), b + gain * v, c - np.outer(gain, g)
            eta, jm = eta + w * v / s, jm + np.outer(w, w) / s
            ell += -0.5 * (np.log(2.0 * np.pi * s) + v * v / s)
    return a, b, c, eta, jm, ell


def pscan_combine_ell(ei, ej):
    """Join of two range elements with log-normalisers (``ei`` earlier)."""
    a, b, c, eta, jm = pscan_combine(ei[:5], ej[:5])
    bi, ci, etaj, jj = ei[1], ei[2], ej[3], ej[4]
    d = bi.shape[-1]
    m_mat = np.eye(d) + ci @ jj
    t1 = etaj - jj @ bi
    f = 0.5 * bi @ (etaj + t1) - 0.5 * np.linalg.slogdet(m_mat)[1] + 0.5 * t1 @ np.linalg.solve(m_mat, ci @ t1)
    return a, b, c, eta, jm, ei[5] + ej[5] + f


# ----------------------------------------------------------------------------------------------
# natural / expectation parameter transforms (ssm_gaussian_transformations.py)
# ----------------------------------------------------------------------------------------------


def ssm_to_expectations(ssm: SSM) -> Tuple[Array, Array, Array]:
    """``ssm_to_expectations`` (``ssm_gaussian_transformations.py:31-89``)."""
    mu = ssm_marginal_means(ssm)[..., None]
    cov = ssm_marginal_covariances(ssm)
    eta_diag = cov + mu @ _t(mu)
    eta_sub = ssm.a_s @ cov[..., :-1, :, :] + mu[..., 1:, :, :] @ _t(mu[..., :-1, :, :])
    return mu[..., 0], eta_diag, eta_sub


def expectations_to_ssm_params(eta_linear, eta_diag, eta_subdiag):
    """``expectations_to_ssm_params`` (``ssm_gaussian_transformations.py:92-178``).
    Returns ``(As, offsets, chol_P0, chol_Qs, mu0)``."""
    eta = eta_linear[..., None]
    covs = eta_diag - eta @ _t(eta)
    covs_sub = _t(eta_subdiag) - eta[..., :-1, :, :] @ _t(eta[..., 1:, :, :])
    chols = chol_lower(covs)
    a_s = _t(cholesky_solve(chols[..., :-1, :, :], covs_sub))
    offsets = (eta[..., 1:, :, :] - a_s @ eta[..., :-1, :, :])[..., 0]
    cond = covs[..., 1:, :, :] - a_s @ (covs[..., :-1, :, :] @ _t(a_s))
    return a_s, offsets, chols[..., 0, :, :], chol_lower(cond), eta[..., 0, :, 0]


def ssm_to_naturals(ssm: SSM):
    """``ssm_to_naturals`` (``ssm_gaussian_transformations.py:181-253``)."""
    a_s = ssm.a_s
    offsets = ssm.concatenated_state_offsets[..., None]
    chols = ssm.concatenated_cholesky_process_covariance
    d = ssm.state_dim
    linv_a = tri_solve_lower(chols[..., 1:, :, :], a_s)
    theta_sub = tri_solve_lower_t(chols[..., 1:, :, :], linv_a)
    tmp = cholesky_solve(chols, offsets)
    theta_lin = np.concatenate(
        [tmp[..., :-1, :, :] - _t(a_s) @ tmp[..., 1:, :, :], tmp[..., -1:, :, :]], axis=-3
    )[..., 0]
    aqa = _t(linv_a) @ linv_a
    aqa = np.concatenate([aqa, np.zeros_like(aqa[..., :1, :, :])], axis=-3)
    prec = cholesky_solve(chols, np.broadcast_to(np.eye(d, dtype=chols.dtype), chols.shape))
    return theta_lin, -0.5 * (prec + aqa), theta_sub


def ssm_to_naturals_no_smoothing(ssm: SSM):
    """``ssm_to_naturals_no_smoothing`` (``ssm_gaussian_transformations.py:256-329``)."""
    chols = ssm.concatenated_cholesky_process_covariance
    d = ssm.state_dim
    theta_sub = cholesky_solve(chols[..., 1:, :, :], ssm.a_s)
    theta_lin = cholesky_solve(chols, ssm.concatenated_state_offsets[..., None])[..., 0]
    prec = cholesky_solve(chols, np.broadcast_to(np.eye(d, dtype=chols.dtype), chols.shape))
    return theta_lin, -0.5 * prec, theta_sub


def naturals_to_ssm_params(theta_linear, theta_diag, theta_subdiag):
    """``naturals_to_ssm_params`` (``ssm_gaussian_transformations.py:332-511``).
    Returns ``(As, offsets, chol_P0, chol_Qs, mu0)``."""
    prec_d, prec_s = -2.0 * theta_diag, -theta_subdiag
    d = prec_d.shape[-1]
    ld, ls = btd_cholesky(prec_d, prec_s)
    covs, covs_sub = btd_inverse_subset(ld, ls, want_sub=True)
    # As = (Σ_kk⁻¹ Σ_{k,k+1})ᵀ, general (LU) solve as in tf.linalg.solve (:461)
    a_s = _t(lu_solve(covs[..., :-1, :, :], _t(covs_sub)))
    # block diagonal of (A⁻ᵀ)⁻¹ tril(P): lower triangle of P_kk + A_{k+1}ᵀ P_{k+1,k}, mirrored (:473-490)
    cond_prec = prec_d.copy()
    cond_prec[..., :-1, :, :] += _t(a_s) @ prec_s
    cond_prec = sym_from_lower(cond_prec)
    chol_cond_prec = chol_lower(cond_prec)
    covariances = cholesky_solve(chol_cond_prec, np.broadcast_to(np.eye(d, dtype=cond_prec.dtype), cond_prec.shape))
    chols = chol_lower(covariances)
    eye = np.broadcast_to(np.eye(d, dtype=prec_d.dtype), prec_d.shape)
    prec_times_offsets = btd_solve(eye, -a_s, theta_linear, transpose_left=True)
    offsets = (covariances @ prec_times_offsets[..., None])[..., 0]
    return a_s, offsets[..., 1:, :], chols[..., 0, :, :], chols[..., 1:, :, :], offsets[..., 0, :]


def naturals_to_ssm_params_no_smoothing(theta_linear, theta_diag, theta_subdiag):
    """``naturals_to_ssm_params_no_smoothing`` (``ssm_gaussian_transformations.py:514-593``)."""
    d = theta_diag.shape[-1]
    chol_cp = chol_lower(-2.0 * theta_diag)
    a_s = cholesky_solve(chol_cp[..., 1:, :, :], theta_subdiag)
    offsets = cholesky_solve(chol_cp, theta_linear[..., None])[..., 0]
    cond_covs = cholesky_solve(chol_cp, np.broadcast_to(np.eye(d, dtype=chol_cp.dtype), chol_cp.shape))
    chols = chol_lower(cond_covs)
    return a_s, offsets[..., 1:, :], chols[..., 0, :, :], chols[..., 1:, :, :], offsets[..., 0, :]


# ----------------------------------------------------------------------------------------------
# closed-form SDE kernels -> SSM parameters (input generators for parity tests and the bench)
# ----------------------------------------------------------------------------------------------


class _Stationary:
    state_dim: int
    jitter: float = 0.0

    def feedback(self) -> Array:
        raise NotImplementedError

    def steady_state_covariance(self) -> Array:
        raise NotImplementedError

    def state_transitions(self, dt: Array) -> Array:
        raise NotImplementedError

    def emission_row(self) -> Array:
        h = np.zeros(self.state_dim)
        h[0] = 1.0
        return h

    def transition_statistics(self, dt: Array) -> Tuple[Array, Array]:
        """``StationaryKernel.transition_statistics`` (``kernels/sde_kernel.py:421-446``):
        ``Q_k = P∞ − A_k P∞ A_kᵀ + jitter·I``."""
        a = self.state_transitions(dt)
        pinf = self.steady_state_covariance()
        q = pinf - a @ pinf @ _t(a)
        return a, q + self.jitter * np.eye(self.state_dim)

    def state_space_model(self, time_points: Array) -> SSM:
        """``SDEKernel.state_space_model`` (``kernels/sde_kernel.py:153-171``), zero state mean."""
        t = np.asarray(time_points, dtype=np.float64)
        dt = t[..., 1:] - t[..., :-1]
        a, q = self.transition_statistics(dt)
        batch = t.shape[:-1]
        p0 = np.broadcast_to(
            self.steady_state_covariance() + self.jitter * np.eye(self.state_dim),
            batch + (self.state_dim, self.state_dim),
        )
        mu0 = np.zeros(batch + (self.state_dim,))
        b = np.zeros(dt.shape + (self.state_dim,))
        return ssm_from_covariances(mu0, p0, a, b, q)

    def emission_matrix(self, time_points: Array) -> Array:
        """``generate_emission_model`` (``kernels/sde_kernel.py:173-211``): ``H=[1,0,…]`` tiled."""
        t = np.asarray(time_points)
        return np.broadcast_to(self.emission_row(), t.shape + (1, self.state_dim)).copy()


class Matern12(_Stationary):
    """``kernels/matern.py:27-143``: ``A = exp(−Δt/ℓ)``, ``P∞ = σ²``."""

    state_dim = 1

    def __init__(self, lengthscale: float, variance: float, jitter: float = 0.0):
        self.lam = 1.0 / lengthscale
        self.variance = variance
        self.jitter = jitter

    def feedback(self):
        return np.array([[-self.lam]])

    def steady_state_covariance(self):
        return self.variance * np.eye(1)

    def state_transitions(self, dt):
        return np.exp(-self.lam * np.asarray(dt))[..., None, None]


class Matern32(_Stationary):
    """``kernels/matern.py:237-372``."""

    state_dim = 2

    def __init__(self, lengthscale: float, variance: float, jitter: float = 0.0):
        self.lam = np.sqrt(3.0) / lengthscale
        self.variance = variance
        self.jitter = jitter

    def feedback(self):
        return np.array([[0.0, 1.0], [-self.lam ** 2, -2.0 * self.lam]])

    def steady_state_covariance(self):
        return self.variance * np.array([[1.0, 0.0], [0.0, self.lam ** 2]])

    def state_transitions(self, dt):
        dt = np.asarray(dt)[..., None, None]
        eye = np.eye(2)
        return np.exp(-self.lam * dt) * (eye + (self.feedback() + self.lam * eye) * dt)


class Matern52(_Stationary):
    """``kernels/matern.py:376-501``."""

    state_dim = 3

    def __init__(self, lengthscale: float, variance: float, jitter: float = 0.0):
        self.lam = np.sqrt(5.0) / lengthscale
        self.variance = variance
        self.jitter = jitter

    def feedback(self):
        l = self.lam
        return np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-(l ** 3), -3.0 * l ** 2, -3.0 * l]])

    def steady_state_covariance(self):
        l23 = self.lam ** 2 / 3.0
        return self.variance * np.array(
            [[1.0, 0.0, -l23], [0.0, l23, 0.0], [-l23, 0.0, self.lam ** 4]]
        )

    def state_transitions(self, dt):
        dt = np.asarray(dt)[..., None, None]
        eye = np.eye(3)
        flt = (self.feedback() + self.lam * eye) * dt
        return np.exp(-self.lam * dt) * (eye + flt + flt @ flt / 2.0)


class HarmonicOscillator(_Stationary):
    """``kernels/periodic.py:27-187``."""

    state_dim = 2

    def __init__(self, variance: float, period: float, jitter: float = 0.0):
        self.lam = 2.0 * np.pi / period
        self.variance = variance
        self.jitter = jitter

    def feedback(self):
        return np.array([[0.0, -self.lam], [self.lam, 0.0]])

    def steady_state_covariance(self):
        return self.variance * np.eye(2)

    def state_transitions(self, dt):
        ang = np.asarray(dt)[..., None, None] * self.lam
        c, s = np.cos(ang), np.sin(ang)
        return np.concatenate(
            [np.concatenate([c, -s], axis=-1), np.concatenate([s, c], axis=-1)], axis=-2
        )


def stationary_ssm_extended_precision(kernel: "_Stationary", time_points: Array) -> SSM:
    """``kernel.state_space_model(time_points)`` with ``A_k`` and ``Q_k = P∞ − A_k P∞ A_kᵀ + jitter·I``
    (``kernels/sde_kernel.py:421-446``) evaluated in ``np.longdouble`` and rounded once.

    For small ``Δt`` the subtraction cancels (Matern52: ``Q_00 = O(Δt⁵)`` against ``‖P∞‖ = O(λ⁴)``), so the
    float64 evaluation of the reference formula -- in TF as much as in numpy -- carries a relative error
    of ``eps·‖P∞‖/Q_00`` that shows up at 1e-9 in a log-likelihood.  This variant separates that rounding
    of the restated reference from the error of an implementation under test.
    Matern12/32/52 only (closed form ``exp(−λΔt)·Σ_j (NΔt)ʲ/j!`` with nilpotent ``N = F + λI``)."""
    ld = np.longdouble
    t = np.asarray(time_points, dtype=np.float64)
    if t.ndim != 1:
        raise ValueError("one series at a time")
    dt = (t[1:] - t[:-1]).astype(ld)
    d = kernel.state_dim
    eye = np.eye(d, dtype=ld)
    lam = ld(kernel.lam)
    n = kernel.feedback().astype(ld) + lam * eye
    a = []
    for x in dt:
        term, acc = eye, eye.copy()
        for j in range(1, d):
            term = term @ n * x / ld(j)
            acc = acc + term
        a.append(np.exp(-lam * x) * acc)
    a = np.stack(a) if len(a) else np.zeros((0, d, d), dtype=ld)
    pinf = kernel.steady_state_covariance().astype(ld)
    q = pinf - a @ pinf @ _t(a) + ld(kernel.jitter) * eye
    p0 = (pinf + ld(kernel.jitter) * eye).astype(np.float64)
    return ssm_from_covariances(np.zeros(d), p0, a.astype(np.float64), np.zeros((len(dt), d)),
                                q.astype(np.float64))


def _block_diag(mats: Sequence[Array]) -> Array:
    batch = np.broadcast_shapes(*[m.shape[:-2] for m in mats])
    n = sum(m.shape[-1] for m in mats)
    out = np.zeros(batch + (n, n))
    o = 0
    for m in mats:
        k = m.shape[-1]
        out[..., o : o + k, o : o + k] = m
        o += k
    return out


class Sum(_Stationary):
    """``kernels/sde_kernel.py:540-687`` (``ConcatKernel`` + ``Sum``): block-diagonal ``F, P∞, A``;
    ``Q`` from the *combined* ``P∞`` with the Sum's own jitter; ``H = [H¹, H², …]``."""

    def __init__(self, kernels: Sequence[_Stationary], jitter: float = 0.0):
        self.kernels = list(kernels)
        self.state_dim = sum(k.state_dim for k in self.kernels)
        self.jitter = jitter

    def feedback(self):
        return _block_diag([k.feedback() for k in self.kernels])

    def steady_state_covariance(self):
        return _block_diag([k.steady_state_covariance() for k in self.kernels])

    def state_transitions(self, dt):
        return _block_diag([k.state_transitions(dt) for k in self.kernels])

    def emission_row(self):
        return np.concatenate([k.emission_row() for k in self.kernels])


# ------------------------------------------------------------------------------------------------
# conditionals.py
# ------------------------------------------------------------------------------------------------

def pairwise_marginals(ssm: SSM, initial_mean: Array, initial_covariance: Array) -> Tuple[Array, Array]:
    """``conditionals.pairwise_marginals`` (``conditionals.py:423-485``): joint of every pair of
    subsequent states, with the initial state prepended and appended."""
    means = ssm_marginal_means(ssm)
    covs = ssm_marginal_covariances(ssm)
    subs = ssm_subsequent_covariances(ssm, covs)
    im = np.broadcast_to(initial_mean, means.shape[:-2] + means.shape[-1:])[..., None, :]
    ic = np.broadcast_to(initial_covariance, covs.shape[:-3] + covs.shape[-2:])[..., None, :, :]
    ext_m = np.concatenate([im, means, im], axis=-2)
    joint_mean = np.concatenate([ext_m[..., :-1, :], ext_m[..., 1:, :]], axis=-1)
    ext_c = np.concatenate([ic, covs, ic], axis=-3)
    zero = np.zeros_like(ic)
    ext_s = np.concatenate([zero, subs, zero], axis=-3)
    top = np.concatenate([ext_c[..., :-1, :, :], np.swapaxes(ext_s, -1, -2)], axis=-1)
    bottom = np.concatenate([ext_s, ext_c[..., 1:, :, :]], axis=-1)
    return joint_mean, np.concatenate([top, bottom], axis=-2)


def conditional_statistics_from_transitions(a_mt: Array, q_mt: Array, a_tp: Array, q_tp: Array,
                                            return_precision: bool = False):
    """``conditionals._conditional_statistics_from_transitions`` (``conditionals.py:128-205``)."""
    a_tp_q_mt = a_tp @ q_mt
    q_mp = q_tp + a_tp @ np.swapaxes(a_tp_q_mt, -1, -2)
    chol = chol_lower(q_mp)
    v = tri_solve_lower(chol, a_tp_q_mt)
    e = np.swapaxes(tri_solve_lower_t(chol, v), -1, -2)
    d = a_mt - e @ a_tp @ a_mt
    if return_precision:
        eye = np.broadcast_to(np.eye(q_mt.shape[-1], dtype=q_mt.dtype), q_mt.shape)
        t = lu_solve(q_mt, eye) + np.swapaxes(a_tp, -1, -2) @ lu_solve(q_tp, a_tp)
    else:
        t = q_mt - np.swapaxes(v, -1, -2) @ v
    return d, e, t


def base_conditional_predict(proj: Array, tcov: Array, adjacent_states: Array,
                             pairwise_state_covariances: Optional[Array] = None):
    """``conditionals.base_conditional_predict`` (``conditionals.py:380-420``)."""
    means = (proj @ adjacent_states[..., None])[..., 0]
    covs = tcov
    if pairwise_state_covariances is not None:
        covs = covs + proj @ pairwise_state_covariances @ np.swapaxes(proj, -1, -2)
    return means, covs
