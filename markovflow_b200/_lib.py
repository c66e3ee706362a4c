"""ctypes binding of ``libmarkovflow_b200.so`` (C ABI in ``include/markovflow_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception
is raised.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C markovflow_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmarkovflow_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "markovflow_b200.h")

MF_OK, MF_ERR_BAD_ARG, MF_ERR_UNSUPPORTED, MF_ERR_CUDA = 0, 1, 2, 3
MF_F32, MF_F64 = 0, 1

_STATUS = {
    MF_ERR_BAD_ARG: "bad argument (null pointer, non-positive dimension or unknown dtype)",
    MF_ERR_UNSUPPORTED: "unsupported dimension for this build",
    MF_ERR_CUDA: "CUDA error",
}


class MarkovflowB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


class CholeskyError(ArithmeticError):
    """Non-positive pivot: the analogue of TF's 'Banded Cholesky decomposition failure'."""


def declared_symbols(header_path: str = HEADER_PATH) -> List[str]:
    """Names of every function the public header declares."""
    with open(header_path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", text)))


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C markovflow_b200/csrc`. markovflow_b200 has no CPU fallback."
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mf_last_cuda_error.restype = ctypes.c_char_p
        _lib.mf_version.restype = ctypes.c_int
    return _lib


def check(status: int, what: str) -> None:
    if status != MF_OK:
        detail = _STATUS.get(status, f"status {status}")
        if status == MF_ERR_CUDA:
            detail += ": " + lib().mf_last_cuda_error().decode()
        raise MarkovflowB200Error(f"{what}: {detail}")


def ptr(t) -> ctypes.c_void_p:
    """Raw device pointer of a (contiguous) torch tensor, or NULL for None."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def i64(v: int) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def current_stream() -> ctypes.c_void_p:
    """The raw ``cudaStream_t`` torch is launching on (current device).  ``torch.cuda.current_stream()`` costs
    ~15 us per call (device-index and availability look-ups behind it); the private accessor it ends in costs ~1 us
    -- these calls sit on the host path of every operator, which a failure check's synchronisation exposes."""
    import torch

    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return ctypes.c_void_p(raw(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dtype) -> int:
    import torch

    if dtype == torch.float64:
        return MF_F64
    if dtype == torch.float32:
        return MF_F32
    raise TypeError(f"markovflow_b200 supports float32/float64 tensors, got {dtype}")
