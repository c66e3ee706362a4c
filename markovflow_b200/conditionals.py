"""Conditionals on top of the Gauss-Markov marginals: drop-in for the kernel-independent part of the
reference's ``markovflow/conditionals.py``.

* :func:`pairwise_marginals` (reference ``conditionals.py:423-485``): ONE fused moment sweep
  (means, covariances and lag-one blocks in a single pass, parallel in time for few long chains)
  and one assembly kernel, instead of ``marginals`` + ``covariance_blocks`` (which repeats the
  covariance recursion) + six concatenations.
* :func:`conditional_statistics_from_transitions` (``:128-205``), :func:`base_conditional_predict`
  (``:380-420``) and :func:`conditional_predict_from_transitions` (``:29-83`` with the SDE kernel's
  ``transition_statistics`` supplied by the caller: kernels are outside the hot-path scope).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr
from .interop import as_torch, boundary, require_cuda


def _prod(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


@boundary
def pairwise_marginals(dist, initial_mean, initial_covariance) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mean ``batch + [T+1, 2D]`` and covariance ``batch + [T+1, 2D, 2D]`` of every pair of
    subsequent states, starting from and reverting to ``N(initial_mean, initial_covariance)``."""
    mean, cov, sub = dist._marginals(True, True, True)
    bs = tuple(dist.batch_shape)
    t, d = mean.shape[-2], mean.shape[-1]
    b = _prod(bs)
    im = as_torch(initial_mean, mean.device).to(mean.dtype)
    ic = as_torch(initial_covariance, mean.device).to(mean.dtype)
    if tuple(im.shape) == (d,):
        init_batch = 1
    else:
        im = im.expand(bs + (d,))
        ic = ic.expand(bs + (d, d))
        init_batch = b
    im = im.reshape(init_batch, d).contiguous()
    ic = ic.reshape(init_batch, d, d).contiguous()
    o_mean = torch.empty(b, t + 1, 2 * d, dtype=mean.dtype, device=mean.device)
    o_cov = torch.empty(b, t + 1, 2 * d, 2 * d, dtype=mean.dtype, device=mean.device)
    check(
        _lib.lib().mf_pairwise_marginals(
            dtype_code(mean.dtype), ptr(mean.reshape(b, t, d)), ptr(cov.reshape(b, t, d, d)),
            ptr(None if sub is None else sub.reshape(b, t - 1, d, d)), ptr(im), ptr(ic), i64(init_batch),
            ptr(o_mean), ptr(o_cov), i64(b), i64(t), i64(d), current_stream()),
        "mf_pairwise_marginals",
    )
    return o_mean.reshape(bs + (t + 1, 2 * d)), o_cov.reshape(bs + (t + 1, 2 * d, 2 * d))


@boundary
def conditional_statistics_from_transitions(
    state_transitions_to_t, process_covariances_to_t, state_transitions_from_t,
    process_covariances_from_t, return_precision: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(D_t, E_t, T_t)`` of ``p(x_t | x_-, x_+) = N(D_t x_- + E_t x_+, T_t)`` (``T_t^{-1}`` when
    ``return_precision``), all ``batch + [N, D, D]``."""
    a_mt = as_torch(state_transitions_to_t)
    require_cuda(a_mt, "state_transitions_to_t")
    lead = tuple(a_mt.shape[:-2])
    d = a_mt.shape[-1]
    n = _prod(lead)
    args = [as_torch(x, a_mt.device).to(a_mt.dtype).expand(lead + (d, d)).reshape(n, d, d).contiguous()
            for x in (a_mt, process_covariances_to_t, state_transitions_from_t, process_covariances_from_t)]
    o_p = torch.empty(n, d, 2 * d, dtype=a_mt.dtype, device=a_mt.device)
    o_t = torch.empty(n, d, d, dtype=a_mt.dtype, device=a_mt.device)
    check(
        _lib.lib().mf_conditional_statistics(
            dtype_code(a_mt.dtype), ptr(args[0]), ptr(args[1]), ptr(args[2]), ptr(args[3]), ptr(o_p),
            ptr(o_t), None, int(bool(return_precision)), i64(n), i64(d), current_stream()),
        "mf_conditional_statistics",
    )
    o_p = o_p.reshape(lead + (d, 2 * d))
    return o_p[..., :d], o_p[..., d:], o_t.reshape(lead + (d, d))


def _predict(proj, tcov, pair_means, pair_covs, indices):
    proj = as_torch(proj)
    require_cuda(proj, "conditional_projections")
    bs = tuple(proj.shape[:-3])
    n, d = proj.shape[-3], proj.shape[-2]
    b = _prod(bs)
    dt, dev = proj.dtype, proj.device
    pm = as_torch(pair_means, dev).to(dt)
    m = pm.shape[-2]
    pc = None if pair_covs is None else as_torch(pair_covs, dev).to(dt).reshape(b, m, 2 * d, 2 * d).contiguous()
    idx = None if indices is None else as_torch(indices, dev).to(torch.int64).reshape(b, n).contiguous()
    o_mean = torch.empty(b, n, d, dtype=dt, device=dev)
    o_cov = torch.empty(b, n, d, d, dtype=dt, device=dev)
    check(
        _lib.lib().mf_conditional_predict(
            dtype_code(dt), ptr(proj.reshape(b, n, d, 2 * d).contiguous()),
            ptr(as_torch(tcov, dev).to(dt).reshape(b, n, d, d).contiguous()),
            ptr(pm.reshape(b, m, 2 * d).contiguous()), ptr(pc), ptr(idx), ptr(o_mean), ptr(o_cov),
            i64(b), i64(n), i64(m), i64(d), current_stream()),
        "mf_conditional_predict",
    )
    return o_mean.reshape(bs + (n, d)), o_cov.reshape(bs + (n, d, d))


@boundary
def base_conditional_predict(conditional_projections, conditional_covariances, adjacent_states,
                             pairwise_state_covariances=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``N(P_t m_t, T_t + P_t S_t P_t^T)`` (``T_t`` alone when no pairwise covariance is given)."""
    return _predict(conditional_projections, conditional_covariances, adjacent_states,
                    pairwise_state_covariances, None)


@boundary
def insertion_indices(new_time_points, training_time_points) -> torch.Tensor:
    """Index of the pair of training states around every new time point (``tf.searchsorted``,
    reference ``conditionals.py:243``); both inputs sorted along the last axis."""
    return torch.searchsorted(as_torch(training_time_points).contiguous(), as_torch(new_time_points).contiguous())


@boundary
def conditional_predict_from_transitions(
    indices, state_transitions_to_t, process_covariances_to_t, state_transitions_from_t,
    process_covariances_from_t, training_pairwise_means, training_pairwise_covariances=None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """:func:`conditional_predict` of the reference with the kernel's ``transition_statistics`` of
    the two gaps around every new time point supplied by the caller (``batch + [N, D, D]`` each) and
    ``indices = insertion_indices(new_time_points, training_time_points)``."""
    d_t, e_t, t_t = conditional_statistics_from_transitions(
        state_transitions_to_t, process_covariances_to_t, state_transitions_from_t, process_covariances_from_t)
    proj = torch.cat([d_t, e_t], dim=-1)
    return _predict(proj, t_t, training_pairwise_means, training_pairwise_covariances, indices)
