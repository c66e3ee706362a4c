"""Stand-in ``tensorflow`` module for the interop tests (TensorFlow is not installable in this image).

It implements exactly the surface ``markovflow_b200.interop`` touches -- an eager tensor type whose
module name starts with ``tensorflow`` and ``tf.experimental.dlpack.{to_dlpack, from_dlpack}`` -- on top
of torch storage, so the TF leg of the boundary (TF tensor in -> zero-copy view -> kernels -> zero-copy
TF tensor out) executes for real.  TEST INFRASTRUCTURE ONLY."""
import torch as _torch

from . import experimental  # noqa: F401


class Tensor:
    """Minimal eager tensor: shares memory with a torch tensor."""

    def __init__(self, storage: "_torch.Tensor"):
        self._t = storage

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return self._t.dtype

    def numpy(self):
        return self._t.detach().cpu().numpy()

    def data_ptr(self):
        return self._t.data_ptr()

    def __getitem__(self, idx):
        return Tensor(self._t[idx])

    def __neg__(self):
        return Tensor(-self._t)

    def __add__(self, other):
        return Tensor(self._t + (other._t if isinstance(other, Tensor) else other))


def constant(value, dtype=None, device=None):
    return Tensor(_torch.as_tensor(value, dtype=dtype, device=device))
