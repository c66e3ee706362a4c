"""Zero-copy tensor interchange at the Python boundary.

The kernels consume raw device pointers.  torch is used here only as the holder of device memory
and streams (``data_ptr()``, ``torch.cuda.current_stream()``); TensorFlow tensors cross through
DLPack (``tf.experimental.dlpack``) without a copy IN BOTH DIRECTIONS, so the reference's TF code can
call these classes on GPU tensors -- and run TF ops on what they return -- with no custom-op rebuild.
TensorFlow is not installed in the build image: the leg is exercised by ``tests/test_tf_interop.py``
with a stand-in ``tensorflow`` module that implements ``tf.experimental.dlpack`` on torch storage.
"""
from __future__ import annotations

import functools
import threading
from typing import Any

import torch


def is_tf_tensor(x: Any) -> bool:
    mod = type(x).__module__ or ""
    return mod.startswith("tensorflow")


def as_torch(x: Any, device: torch.device | None = None) -> torch.Tensor:
    """View ``x`` as a torch tensor without copying when it already lives on a CUDA device."""
    if isinstance(x, torch.Tensor):
        return x
    if is_tf_tensor(x):
        import tensorflow as tf

        return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(x))
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.as_tensor(x, device=dev)


def to_tf(t: torch.Tensor) -> Any:
    """A torch tensor as a TensorFlow tensor sharing its memory (DLPack, no copy)."""
    import tensorflow as tf

    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.detach().contiguous()))


def like_input(t: torch.Tensor, like: Any) -> Any:
    """Return ``t`` in the framework of ``like`` (TF in -> TF out, zero-copy); torch otherwise."""
    return to_tf(t) if framework_of(like) == "tf" else t


# ---- the framework of a call: TensorFlow in -> TensorFlow out ---------------------------------------
# Every public entry point of the package is wrapped in :func:`boundary`.  The OUTERMOST call decides the
# framework of its results from its arguments (a TF tensor, or an object of this package that was built
# from TF tensors, anywhere among them); nested calls between the package's own classes see torch
# tensors only.  Objects of this package that are returned (``.cholesky``, ``.precision``,
# ``posterior_state_space_model()`` ...) inherit the tag, so the whole chain of results stays in the
# caller's framework -- the reference's callers run TF ops on them (``kalman_filter.py:234-255``).
_state = threading.local()


def framework_of(*objs: Any) -> str:
    for o in objs:
        if is_tf_tensor(o) or getattr(o, "_fw", None) == "tf":
            return "tf"
        if isinstance(o, (tuple, list)) and framework_of(*o) == "tf":
            return "tf"
    return "torch"


def to_framework(obj: Any, fw: str) -> Any:
    if fw != "tf" or obj is None:
        return obj
    if isinstance(obj, torch.Tensor):
        return to_tf(obj)
    if isinstance(obj, (tuple, list)) and not isinstance(obj, torch.Size):
        return type(obj)(to_framework(o, fw) for o in obj)
    if hasattr(obj, "_fw"):
        obj._fw = fw
    return obj


def _cuda_device_of(objs) -> "torch.device | None":
    """Device of the first CUDA operand (a tensor, or an object of this package: ``_dev``)."""
    for o in objs:
        if isinstance(o, torch.Tensor):
            if o.is_cuda:
                return o.device
        else:
            d = getattr(o, "_dev", None)
            if isinstance(d, torch.device) and d.type == "cuda":
                return d
    return None


def boundary(fn):
    """Mark ``fn`` (function or method) as a public entry point: see the note above.  The outermost
    call also makes the operands' device current for its duration (the C ABI launches on the current
    device, on torch's current stream of that device)."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        depth = getattr(_state, "depth", 0)
        if depth:
            return fn(*args, **kwargs)
        _state.depth = 1
        try:
            dev = _cuda_device_of(args)
            if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    out = fn(*args, **kwargs)
            else:
                out = fn(*args, **kwargs)
        finally:
            _state.depth = 0
        return to_framework(out, framework_of(*args, *kwargs.values()))

    return wrapper


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: markovflow_b200 runs on the GPU only (no CPU fallback)"
        )
