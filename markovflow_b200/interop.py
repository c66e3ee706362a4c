"""Zero-copy tensor interchange at the Python boundary.

The kernels consume raw device pointers.  torch is used here only as the holder of device memory
and streams (``data_ptr()``, ``torch.cuda.current_stream()``); TensorFlow tensors cross through
DLPack (``tf.experimental.dlpack``) without a copy, so the reference's TF code can call these
classes on GPU tensors with no custom-op rebuild.  TensorFlow is not installed in the build image,
so the TF leg is exercised only where TF exists.
"""
from __future__ import annotations

from typing import Any

import torch


def is_tf_tensor(x: Any) -> bool:
    mod = type(x).__module__ or ""
    return mod.startswith("tensorflow")


def as_torch(x: Any, device: torch.device | None = None) -> torch.Tensor:
    """View ``x`` as a torch tensor without copying when it already lives on a CUDA device."""
    if isinstance(x, torch.Tensor):
        return x
    if is_tf_tensor(x):  # pragma: no cover - TF absent in this image
        import tensorflow as tf

        return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(x))
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.as_tensor(x, device=dev)


def like_input(t: torch.Tensor, like: Any) -> Any:
    """Return ``t`` in the framework of ``like`` (TF in -> TF out, zero-copy); torch otherwise."""
    if is_tf_tensor(like):  # pragma: no cover - TF absent in this image
        import tensorflow as tf

        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))
    return t


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: markovflow_b200 runs on the GPU only (no CPU fallback)"
        )
