"""The TensorFlow leg of the drop-in boundary (reference seam: ``markovflow/block_tri_diag.py:22-31``
imports the banded TF ops; ``kalman_filter.py:234-255`` runs TF ops on what the operators return).

TensorFlow cannot be installed in this image, so a stand-in ``tensorflow`` module
(``tests/tf_stub``) provides the eager-tensor type and ``tf.experimental.dlpack``.  With it the leg runs
for real: TF tensor in -> zero-copy torch view -> (kernels) -> zero-copy TF tensor out, for tensors AND
for the operator objects a call returns.  CPU part: accessors / elementwise paths (no kernels);
GPU part: the CUDA operators on TF-owned device memory."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import max_rel_err, random_ssm_arrays, random_well_conditioned_spd_btd

STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_stub")


@pytest.fixture()
def tf(monkeypatch):
    monkeypatch.syspath_prepend(STUB)
    for name in [n for n in sys.modules if n == "tensorflow" or n.startswith("tensorflow.")]:
        monkeypatch.delitem(sys.modules, name)
    import tensorflow

    assert tensorflow.__file__.startswith(STUB)
    yield tensorflow
    for name in [n for n in sys.modules if n == "tensorflow" or n.startswith("tensorflow.")]:
        sys.modules.pop(name, None)


def test_tf_tensors_cross_without_a_copy_and_come_back_as_tf(tf):
    from markovflow_b200 import interop

    x = tf.constant(np.arange(12.0).reshape(3, 4))
    assert interop.is_tf_tensor(x) and interop.framework_of(x) == "tf"
    view = interop.as_torch(x)
    assert isinstance(view, torch.Tensor) and view.data_ptr() == x.data_ptr()  # zero copy in
    back = interop.like_input(view, x)
    assert interop.is_tf_tensor(back) and back.data_ptr() == x.data_ptr()  # zero copy out
    assert interop.like_input(view, view) is view  # torch in -> torch out


def test_operator_objects_keep_the_callers_framework_cpu(tf):
    """Accessors, ``__add__``, ``to_dense`` (no kernels involved): TF in -> TF out, through the objects
    the calls return; torch in -> torch out."""
    from markovflow_b200 import StateSpaceModel, SymmetricBlockTriDiagonal, interop

    diag, sub, _, _ = random_well_conditioned_spd_btd((2,), 4, 3, rng=1)
    m_tf = SymmetricBlockTriDiagonal(tf.constant(diag), tf.constant(sub))
    assert interop.is_tf_tensor(m_tf.block_diagonal) and interop.is_tf_tensor(m_tf.block_sub_diagonal)
    added = m_tf + m_tf
    assert interop.is_tf_tensor(added.block_diagonal)  # the returned object inherited the framework
    np.testing.assert_allclose(added.block_diagonal.numpy(), 2 * diag)
    assert interop.is_tf_tensor(m_tf.to_dense()) and interop.is_tf_tensor(m_tf.as_band)
    m_t = SymmetricBlockTriDiagonal(torch.as_tensor(diag), torch.as_tensor(sub))
    assert isinstance(m_t.block_diagonal, torch.Tensor) and isinstance((m_t + m_t).block_diagonal, torch.Tensor)

    arrays = random_ssm_arrays((2,), 5, 2)
    ssm = StateSpaceModel(*(tf.constant(a) for a in arrays))
    for got, want in ((ssm.state_transitions, arrays[2]), (ssm.initial_mean, arrays[0]),
                      (ssm.concatenated_state_offsets, np.concatenate([arrays[0][:, None], arrays[3]], axis=1))):
        assert interop.is_tf_tensor(got)
        np.testing.assert_allclose(got.numpy(), want)
    a_inv = ssm.a_inv_block
    assert interop.is_tf_tensor(a_inv.block_sub_diagonal)
    np.testing.assert_allclose(a_inv.block_sub_diagonal.numpy(), -arrays[2])


@pytest.mark.gpu
def test_cuda_operators_on_tf_owned_device_memory(tf):
    """Cholesky + solve, Kalman log-likelihood / posterior and a transform with TF GPU tensors as inputs:
    every result (and every result of the returned objects) is a TF tensor, values equal the oracle."""
    import markovflow_b200 as mf
    from markovflow_b200 import interop

    dev = torch.device("cuda:0")
    tfc = lambda x: tf.constant(np.ascontiguousarray(x), device=dev)
    diag, sub, _, _ = random_well_conditioned_spd_btd((3,), 9, 3, rng=2)
    rhs = np.random.default_rng(0).standard_normal((3, 9, 3))
    m = mf.SymmetricBlockTriDiagonal(tfc(diag), tfc(sub))
    chol = m.cholesky
    o_ld, o_ls = O.btd_cholesky(diag, sub)
    assert interop.is_tf_tensor(chol.block_diagonal)
    assert max_rel_err(chol.block_diagonal.numpy(), o_ld) < 1e-10
    x = chol.solve(tfc(rhs))
    assert interop.is_tf_tensor(x) and max_rel_err(x.numpy(), O.btd_solve(o_ld, o_ls, rhs)) < 1e-10
    assert interop.is_tf_tensor(chol.abs_log_det()) and interop.is_tf_tensor(m.dense_mult(tfc(rhs)))

    np.random.seed(3)
    arrays = random_ssm_arrays((2,), 11, 2)
    ref = O.SSM(*arrays)
    rng = np.random.default_rng(4)
    h, y, lr = rng.standard_normal((12, 1, 2)), rng.standard_normal((2, 12, 1)), np.array([[0.3]])
    ssm = mf.StateSpaceModel(*(tfc(a) for a in arrays))
    mean, cov = ssm.marginals
    assert interop.is_tf_tensor(mean) and interop.is_tf_tensor(cov)
    assert max_rel_err(mean.numpy(), O.ssm_marginal_means(ref)) < 1e-10
    kf = mf.KalmanFilter(ssm, mf.EmissionModel(tfc(h)), tfc(y), tfc(lr))
    ll = kf.log_likelihood()
    assert interop.is_tf_tensor(ll)
    want = O.kalman_log_likelihood(ref, h, y, O._r_inv_from_chol(lr))
    assert max_rel_err(ll.numpy(), want) < 1e-10
    post = kf.posterior_state_space_model()
    assert interop.is_tf_tensor(post.state_transitions) and interop.is_tf_tensor(post.marginal_means)
    ref_post = O.kalman_posterior_ssm(ref, h, y, O._r_inv_from_chol(lr))
    assert max_rel_err(post.state_transitions.numpy(), ref_post.a_s) < 1e-10
    th = mf.ssm_to_naturals(ssm)
    assert all(interop.is_tf_tensor(t) for t in th)
    back = mf.naturals_to_ssm_params(*th)
    assert all(interop.is_tf_tensor(t) for t in back)
    assert max_rel_err(back[0].numpy(), arrays[2]) < 1e-10
    # torch in -> torch out is unchanged
    ssm_t = mf.StateSpaceModel(*(torch.as_tensor(a, device=dev) for a in arrays))
    assert isinstance(ssm_t.marginal_means, torch.Tensor)
