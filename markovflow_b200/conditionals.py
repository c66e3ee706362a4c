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


APPROX_INF = 1e10  # reference markovflow/base.py:46: the "time" of the phantom states before / after the data


def _conditional_statistics(new_time_points, training_time_points, kernel):
    """``(P_t, T_t, indices)`` for every new time point (reference ``conditionals.py:207-254``): the two
    gaps around it go through ``kernel.transition_statistics`` (closed forms, per point), the conditional
    statistics through ``mf_conditional_statistics``."""
    new = as_torch(new_time_points)
    train = as_torch(training_time_points, new.device).to(new.dtype)
    idx = torch.searchsorted(train.contiguous(), new.contiguous())
    inf = APPROX_INF * torch.ones_like(train[..., -1:])
    aug = torch.cat([-inf, train, inf], dim=-1)
    minus = torch.gather(aug, -1, idx)
    plus = torch.gather(aug, -1, idx + 1)
    a_mt, q_mt = kernel.transition_statistics(minus, new - minus)
    a_tp, q_tp = kernel.transition_statistics(new, plus - new)
    d_t, e_t, t_t = conditional_statistics_from_transitions(a_mt, q_mt, a_tp, q_tp)
    return torch.cat([d_t, e_t], dim=-1), t_t, idx


@boundary
def conditional_statistics(new_time_points, training_time_points, kernel) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(P_t, T_t)`` of ``p(x_t | x_-, x_+) = N(P_t [x_-, x_+], T_t)`` (reference ``conditionals.py:86-125``)."""
    p, t, _ = _conditional_statistics(new_time_points, training_time_points, kernel)
    return p, t


@boundary
def conditional_predict(new_time_points, training_time_points, kernel, training_pairwise_means,
                        training_pairwise_covariances=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's ``conditional_predict`` (``conditionals.py:29-83``), same signature: marginals
    ``N(P_t m_t, T_t + P_t S_t P_t^T)`` at the (sorted) ``new_time_points`` given the pairwise marginals of the
    states at the (sorted) ``training_time_points``; without covariances, the conditional density.  The gather
    of the pairs by insertion index is fused into the prediction kernel."""
    proj, tcov, idx = _conditional_statistics(new_time_points, training_time_points, kernel)
    return _predict(proj, tcov, training_pairwise_means, training_pairwise_covariances, idx)
