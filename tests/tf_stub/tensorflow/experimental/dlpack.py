"""``tf.experimental.dlpack`` of the stand-in module: DLPack capsules over torch storage."""
import torch as _torch


def to_dlpack(tensor):
    return _torch.utils.dlpack.to_dlpack(tensor._t)


def from_dlpack(capsule):
    from tensorflow import Tensor

    return Tensor(_torch.utils.dlpack.from_dlpack(capsule))
