"""Random block-tridiagonal / state-space-model generators for the parity tests.

They follow the *recipes* of the reference's test generators so that the same families of inputs
are exercised: ``tests/unit/test_block_tri_diag.py:228-312`` (lower block-bidiagonal L with
``loc=1`` diagonal, ``M = L Lᵀ``) and ``tests/tools/state_space_model.py:35-81``.
"""
import json
import os
from typing import Optional, Tuple

import numpy as np


def dense_from_blocks(diag: np.ndarray, sub: Optional[np.ndarray]) -> np.ndarray:
    """Lower-triangular dense matrix from blocks (upper triangles of diagonal blocks dropped)."""
    *batch, t, d, _ = diag.shape
    out = np.zeros(tuple(batch) + (t * d, t * d))
    for k in range(t):
        out[..., k * d:(k + 1) * d, k * d:(k + 1) * d] = np.tril(diag[..., k, :, :])
        if sub is not None and k + 1 < t:
            out[..., (k + 1) * d:(k + 2) * d, k * d:(k + 1) * d] = sub[..., k, :, :]
    return out


def blocks_from_dense(dense: np.ndarray, d: int, with_sub: bool):
    n = dense.shape[-1] // d
    diag = np.stack([dense[..., k * d:(k + 1) * d, k * d:(k + 1) * d] for k in range(n)], axis=-3)
    sub = None
    if with_sub:
        sub = np.stack(
            [dense[..., (k + 1) * d:(k + 2) * d, k * d:(k + 1) * d] for k in range(n - 1)], axis=-3
        )
    return diag, sub


def random_lower_btd(batch_shape: Tuple, t: int, d: int, with_sub: bool, loc: float = 0.0):
    """Random lower block-bidiagonal matrix: returns (dense, diag, sub)."""
    diag = np.tril(np.random.normal(loc=loc, size=batch_shape + (t, d, d)))
    sub = np.random.normal(size=batch_shape + (t - 1, d, d)) if with_sub else None
    return dense_from_blocks(diag, sub), diag, sub


def random_spd_btd(batch_shape: Tuple, t: int, d: int, with_sub: bool):
    """Random SPD block-tridiagonal ``M = L Lᵀ``: returns (dense, diag, sub) with FULL diag blocks."""
    dense_l, _, _ = random_lower_btd(batch_shape, t, d, with_sub, loc=1.0)
    dense = dense_l @ np.swapaxes(dense_l, -1, -2)
    diag, sub = blocks_from_dense(dense, d, with_sub)
    return dense, diag, sub


def random_well_conditioned_spd_btd(batch_shape: Tuple, t: int, d: int, rng=None):
    """SPD block-tridiagonal with bounded condition number (for long chains / tight tolerances):
    ``L`` has diagonal blocks ``I + 0.3·tril(N)`` (positive diagonal) and sub blocks ``0.3·N``."""
    rng = np.random.default_rng(rng)
    ld = 0.3 * np.tril(rng.standard_normal(batch_shape + (t, d, d)), -1)
    ld = ld + (1.0 + 0.5 * rng.random(batch_shape + (t, d)))[..., None] * np.eye(d)
    ls = 0.3 * rng.standard_normal(batch_shape + (t - 1, d, d))
    diag = ld @ np.swapaxes(ld, -1, -2)
    diag[..., 1:, :, :] += ls @ np.swapaxes(ls, -1, -2)
    sub = ls @ np.swapaxes(ld[..., :-1, :, :], -1, -2)
    return diag, sub, ld, ls


def random_ssm_arrays(batch_shape: Tuple, num_transitions: int, d: int, scale_a: float = 0.5):
    """(mu0, chol_P0, A, b, chol_Q) with positive-diagonal Cholesky factors."""
    def chol(shape):
        l = np.tril(np.random.normal(size=shape + (d, d)), -1) * 0.3
        return l + (0.5 + np.random.uniform(size=shape + (d,)))[..., None] * np.eye(d)

    mu0 = np.random.normal(size=batch_shape + (d,))
    chol_p0 = chol(batch_shape)
    a_s = scale_a * np.random.normal(size=batch_shape + (num_transitions, d, d))
    b_s = np.random.normal(size=batch_shape + (num_transitions, d))
    chol_q = chol(batch_shape + (num_transitions,))
    return mu0, chol_p0, a_s, b_s, chol_q


def dense_ssm_mean_cov(mu0, chol_p0, a_s, b_s, chol_q):
    """Dense joint mean [.., T*D] and covariance [.., T*D, T*D] of an SSM by direct propagation
    (single chain, no batch)."""
    n, d = a_s.shape[0], a_s.shape[-1]
    t = n + 1
    means = [mu0]
    p = [chol_p0 @ chol_p0.T]
    for k in range(n):
        means.append(a_s[k] @ means[-1] + b_s[k])
        p.append(a_s[k] @ p[-1] @ a_s[k].T + chol_q[k] @ chol_q[k].T)
    cov = np.zeros((t * d, t * d))
    for i in range(t):
        cov[i * d:(i + 1) * d, i * d:(i + 1) * d] = p[i]
        c = p[i]
        for j in range(i + 1, t):
            c = a_s[j - 1] @ c
            cov[j * d:(j + 1) * d, i * d:(i + 1) * d] = c
            cov[i * d:(i + 1) * d, j * d:(j + 1) * d] = c.T
    return np.concatenate(means), cov


def max_rel_err(x, ref) -> float:
    """``max|x − ref| / max|ref|`` -- the parity measure of SURVEY.md §8(d)."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    denom = np.max(np.abs(ref)) if ref.size else 1.0
    return float(np.max(np.abs(x - ref)) / max(denom, 1e-300)) if ref.size else 0.0


# ------------------------------------------------------------------------------------------------
# The parity rule (BASELINE.json north_star; SURVEY.md §7 "Conditioning vs the 1e-10 tolerance", §8d)
# ------------------------------------------------------------------------------------------------
TOL64, TOL32 = 1e-10, 1e-4


def ld(*arrays):
    """The arrays in extended precision (numpy long double: 64-bit mantissa on x86)."""
    out = tuple(None if a is None else np.asarray(a, dtype=np.longdouble) for a in arrays)
    return out if len(out) > 1 else out[0]


def _err(x, ref) -> float:
    x, ref = np.asarray(x, dtype=np.longdouble), np.asarray(ref, dtype=np.longdouble)
    if not ref.size:
        return 0.0
    return float(np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), np.longdouble(1e-300)))


def _log_parity(record: dict) -> None:
    path = os.environ.get("MF_PARITY_LOG")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(record) + "\n")


def assert_parity(got, want, tol: float, *, truth=None, peer=None, what: str = "") -> float:
    """``max|got - want| / max|want| <= tol`` with tol = 1e-10 (float64) / 1e-4 (float32) -- nothing looser.

    Where the restated reference ITSELF cannot deliver ``tol`` on an input (it takes an ill-conditioned
    route: e.g. covariances through precision -> Cholesky -> inverse subset, ``state_space_model.py:
    253-262``), the rule of SURVEY.md §7 applies and is evaluated HERE, in the test, in extended precision:

    * ``truth`` (callable): the same quantity evaluated in long double (``want`` is then the float64
      restatement of the reference).  Pass iff the CUDA result is within ``tol`` of the truth, or at
      least as close to the truth as the restated reference is.
    * ``peer`` (callable): for float32 results, ``want`` is the float64 truth on the float32-rounded
      inputs and ``peer`` the restated reference evaluated in float32 arithmetic.  Pass iff the CUDA
      result is within ``tol`` of the truth or at least as close as the float32 reference.
    """
    e = _err(got, want)
    rec = {"what": what, "tol": tol, "err_vs_oracle": e}
    if e <= tol:
        _log_parity(rec)
        return e
    if truth is not None:
        tr = truth()
        e_got, e_ref = _err(got, tr), _err(want, tr)
        rec.update(err_vs_long_double=e_got, oracle_err_vs_long_double=e_ref)
        _log_parity(rec)
        assert e_got <= max(tol, e_ref), (
            f"{what}: {e:.2e} from the float64 oracle; against the long-double evaluation the CUDA "
            f"result is off by {e_got:.2e}, the restated reference by {e_ref:.2e} (tol {tol:.0e})")
        return e_got
    if peer is not None:
        e_ref = _err(peer(), want)
        rec.update(same_precision_oracle_err=e_ref)
        _log_parity(rec)
        assert e <= max(tol, e_ref), (
            f"{what}: {e:.2e} from the float64 truth; the reference algorithm in the same "
            f"arithmetic is off by {e_ref:.2e} (tol {tol:.0e})")
        return e
    _log_parity(rec)
    raise AssertionError(f"{what}: max rel err {e:.3e} > {tol:.0e}")
