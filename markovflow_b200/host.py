"""Host-buffer entry point: block-tridiagonal Cholesky + solve for data that lives in HOST memory.

This is what a host-resident caller (the reference runs TensorFlow on CPU tensors) would use: the
batch is cut into chunks of chains, and host->device copies, the fused CUDA sweep and device->host
copies of successive chunks overlap on three streams (PCIe is full duplex).  Device staging slots
are cached between calls.  Inputs/outputs should be pinned for the copies to be asynchronous.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, dtype_code, i64, ptr
from .block_tri_diag import _raise_if_failed

_SLOTS: Dict[tuple, dict] = {}
NSLOT = 3


def _slots(chunk: int, t: int, d: int, dtype, device, with_rhs: bool) -> dict:
    key = (chunk, t, d, dtype, str(device), with_rhs)
    if key not in _SLOTS:
        def buf(*shape):
            return [torch.empty(shape, dtype=dtype, device=device) for _ in range(NSLOT)]

        _SLOTS[key] = {
            "diag": buf(chunk, t, d, d), "sub": buf(chunk, t - 1, d, d),
            "rhs": buf(chunk, t, d) if with_rhs else None,
            "ld": buf(chunk, t, d, d), "ls": buf(chunk, t - 1, d, d),
            "x": buf(chunk, t, d) if with_rhs else None,
            "info": [torch.empty(chunk, dtype=torch.int32, device=device) for _ in range(NSLOT)],
            "streams": [torch.cuda.Stream(device=device) for _ in range(3)],
        }
    return _SLOTS[key]


def cholesky_solve_host(
    diag: torch.Tensor,
    sub: torch.Tensor,
    rhs: Optional[torch.Tensor] = None,
    out: Optional[Tuple[torch.Tensor, ...]] = None,
    chunk: int = 128,
    device: Optional[torch.device] = None,
):
    """``SymmetricBlockTriDiagonal(diag, sub).cholesky`` (+ ``.solve(rhs)``) on host tensors.

    ``diag [B,T,D,D]``, ``sub [B,T-1,D,D]``, ``rhs [B,T,D]`` are CPU tensors; returns CPU tensors
    ``(Ld, Ls, x_or_None, info)`` (written into ``out`` when given).  Returns after all copies have
    landed.  Counts of bytes moved: ``(h2d_bytes, d2h_bytes)`` are attached as attributes of the
    function (``cholesky_solve_host.last_bytes``).
    """
    assert not diag.is_cuda and not sub.is_cuda, "host entry point takes CPU tensors"
    device = device or torch.device("cuda", torch.cuda.current_device())
    b, t, d, _ = diag.shape
    dtype = diag.dtype
    chunk = min(chunk, b)
    s = _slots(chunk, t, d, dtype, device, rhs is not None)
    h2d, comp, d2h = s["streams"]
    if out is None:
        pin = diag.is_pinned()
        ld_h = torch.empty_like(diag, pin_memory=pin)
        ls_h = torch.empty_like(sub, pin_memory=pin)
        x_h = torch.empty_like(rhs, pin_memory=pin) if rhs is not None else None
        info_h = torch.empty(b, dtype=torch.int32, pin_memory=pin)
    else:
        ld_h, ls_h, x_h, info_h = out
    lib = _lib.lib()
    code = dtype_code(dtype)
    start = torch.cuda.current_stream(device)
    ev_start = torch.cuda.Event()
    ev_start.record(start)
    h2d.wait_event(ev_start)
    slot_free = [None] * NSLOT
    nchunks = (b + chunk - 1) // chunk
    h2d_bytes = d2h_bytes = 0
    for c in range(nchunks):
        b0, b1 = c * chunk, min(b, (c + 1) * chunk)
        nb = b1 - b0
        k = c % NSLOT
        with torch.cuda.stream(h2d):
            if slot_free[k] is not None:
                h2d.wait_event(slot_free[k])
            s["diag"][k][:nb].copy_(diag[b0:b1], non_blocking=True)
            s["sub"][k][:nb].copy_(sub[b0:b1], non_blocking=True)
            h2d_bytes += diag[b0:b1].numel() * diag.element_size() + sub[b0:b1].numel() * sub.element_size()
            if rhs is not None:
                s["rhs"][k][:nb].copy_(rhs[b0:b1], non_blocking=True)
                h2d_bytes += rhs[b0:b1].numel() * rhs.element_size()
            ev_in = torch.cuda.Event()
            ev_in.record(h2d)
        with torch.cuda.stream(comp):
            comp.wait_event(ev_in)
            check(
                lib.mf_btd_cholesky(
                    code, ptr(s["diag"][k]), ptr(s["sub"][k]),
                    ptr(s["rhs"][k]) if rhs is not None else None,
                    ptr(s["ld"][k]), ptr(s["ls"][k]),
                    ptr(s["x"][k]) if rhs is not None else None, None, ptr(s["info"][k]),
                    i64(nb), i64(t), i64(d), _lib.ctypes.c_void_p(comp.cuda_stream),
                ),
                "mf_btd_cholesky",
            )
            ev_done = torch.cuda.Event()
            ev_done.record(comp)
        with torch.cuda.stream(d2h):
            d2h.wait_event(ev_done)
            ld_h[b0:b1].copy_(s["ld"][k][:nb], non_blocking=True)
            ls_h[b0:b1].copy_(s["ls"][k][:nb], non_blocking=True)
            d2h_bytes += ld_h[b0:b1].numel() * ld_h.element_size() + ls_h[b0:b1].numel() * ls_h.element_size()
            if rhs is not None:
                x_h[b0:b1].copy_(s["x"][k][:nb], non_blocking=True)
                d2h_bytes += x_h[b0:b1].numel() * x_h.element_size()
            info_h[b0:b1].copy_(s["info"][k][:nb], non_blocking=True)
            d2h_bytes += nb * 4
            ev_out = torch.cuda.Event()
            ev_out.record(d2h)
            slot_free[k] = ev_out
    start.wait_stream(d2h)
    d2h.synchronize()
    cholesky_solve_host.last_bytes = (h2d_bytes, d2h_bytes)
    cholesky_solve_host.last_launches = nchunks
    _raise_if_failed(info_h, "cholesky_solve_host")
    return ld_h, ls_h, x_h, info_h
