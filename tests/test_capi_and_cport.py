"""CPU-side checks: the C-ABI library loads and exports every symbol the public header declares
(no compute calls without a GPU), bad arguments are rejected before any launch, and the C port of
the reference's banded CPU path (oracle/banded_ref.c, the CPU baseline) agrees with the numpy
oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import max_rel_err, random_well_conditioned_spd_btd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(ROOT, "markovflow_b200", "csrc", "libmarkovflow_b200.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libbanded_ref.so")):
        g.build()
    return True


def test_library_exports_every_declared_symbol(built):
    from markovflow_b200 import _lib

    names = _lib.declared_symbols()
    assert len(names) >= 8 and "mf_btd_cholesky" in names
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/markovflow_b200.h but not exported"
    assert handle.mf_version() >= 100


def test_header_is_plain_c(built, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "markovflow_b200.h"\nint main(void){return MF_OK;}\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_bad_arguments_rejected_without_launch(built):
    from markovflow_b200 import _lib

    h = _lib.lib()
    null = ctypes.c_void_p(0)
    i64 = ctypes.c_int64
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert h.mf_btd_cholesky(_lib.MF_F64, null, null, null, null, null, null, null, null, i64(1), i64(1), i64(1), null) == _lib.MF_ERR_BAD_ARG
    assert h.mf_btd_cholesky(_lib.MF_F64, one, null, null, one, null, null, null, null, i64(1), i64(0), i64(1), null) == _lib.MF_ERR_BAD_ARG
    assert h.mf_btd_cholesky(7, one, null, null, one, null, null, null, null, i64(1), i64(1), i64(2), null) == _lib.MF_ERR_BAD_ARG
    # sub given without out_sub
    assert h.mf_btd_cholesky(_lib.MF_F64, one, one, null, one, null, null, null, null, i64(1), i64(3), i64(2), null) == _lib.MF_ERR_BAD_ARG
    # empty batch is a no-op
    assert h.mf_btd_cholesky(_lib.MF_F64, null, null, null, null, null, null, null, null, i64(0), i64(3), i64(2), null) == _lib.MF_OK
    assert h.mf_btd_solve(_lib.MF_F64, null, null, null, null, i64(1), i64(1), i64(1), i64(1), 0, null) == _lib.MF_ERR_BAD_ARG
    assert h.mf_set_tuning(99, 0) == _lib.MF_ERR_BAD_ARG


def test_package_has_no_oracle_dependency():
    """The product must not import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "markovflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "np_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("d,t", [(1, 7), (2, 12), (3, 33), (5, 9)])
def test_c_port_matches_numpy_oracle(built, d, t):
    from oracle import c_ref

    diag, sub, _, _ = random_well_conditioned_spd_btd((5,), t, d, rng=d * 10 + t)
    rhs = np.random.default_rng(0).standard_normal((5, t, d))
    ld, ls, x, info = c_ref.chol_solve_batch(diag, sub, rhs)
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    assert max_rel_err(ld, o_ld) < 1e-12 and max_rel_err(ls, o_ls) < 1e-12
    assert max_rel_err(x, O.btd_solve(o_ld, o_ls, rhs)) < 1e-12
    assert not info.any()


def test_c_port_no_subdiag_and_failure_index(built):
    from oracle import c_ref

    diag, _, _, _ = random_well_conditioned_spd_btd((3,), 4, 2, rng=3)
    ld, ls, x, info = c_ref.chol_solve_batch(diag, None, None)
    assert ls is None and x is None
    assert max_rel_err(ld, O.btd_cholesky(diag, None)[0]) < 1e-12
    diag, sub, _, _ = random_well_conditioned_spd_btd((3,), 6, 2, rng=5)
    diag[1, 4] = -np.eye(2)
    _, _, _, info = c_ref.chol_solve_batch(diag, sub, None)
    assert list(info) == [0, 5, 0]


@pytest.mark.parametrize("d,m,t,hb", [(1, 1, 9, 1), (2, 1, 40, 1), (3, 2, 17, 3), (5, 3, 6, 3)])
def test_c_port_kalman_log_likelihood_matches_numpy_oracle(built, d, m, t, hb):
    """oracle/ssm_ref.c::ref_kalman_loglik_batch walks kalman_filter.py:184-255 like the numpy restatement."""
    from oracle import c_ref
    from tests.helpers import random_ssm_arrays

    np.random.seed(d * 100 + t)
    b = 3
    mu0, l0, a, bb, lq = random_ssm_arrays((b,), t - 1, d)
    rng = np.random.default_rng(d + t)
    h = rng.standard_normal((t, m, d) if hb == 1 else (b, t, m, d))
    y = rng.standard_normal((b, t, m))
    lr = np.tril(rng.standard_normal((m, m))) * 0.2 + 0.5 * np.eye(m)
    got = c_ref.kalman_loglik_batch(mu0, l0, a, bb, lq, h, y, lr)
    want = O.kalman_log_likelihood(O.SSM(mu0, l0, a, bb, lq), h, y, O._r_inv_from_chol(lr), per_chain=True)
    assert max_rel_err(got, want) < 1e-12


@pytest.mark.parametrize("d,t", [(1, 8), (2, 33), (3, 12), (4, 7)])
def test_c_port_transforms_match_numpy_oracle(built, d, t):
    """ref_nat_to_ssm_batch / ref_ssm_to_expectations_batch against the numpy restatements of
    ssm_gaussian_transformations.py:332-511 and :31-89."""
    from oracle import c_ref
    from tests.helpers import random_ssm_arrays

    np.random.seed(d * 10 + t)
    arrays = random_ssm_arrays((4,), t - 1, d)
    ssm = O.SSM(*arrays)
    for g, w in zip(c_ref.ssm_to_expectations_batch(*arrays), O.ssm_to_expectations(ssm)):
        assert max_rel_err(g, w) < 1e-12
    th = O.ssm_to_naturals(ssm)
    for g, w in zip(c_ref.nat_to_ssm_batch(*th), O.naturals_to_ssm_params(*th)):
        assert max_rel_err(g, w) < 1e-11
