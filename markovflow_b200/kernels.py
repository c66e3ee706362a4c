"""Stationary Matern priors whose Kalman log-likelihood is evaluated WITHOUT materialising the SSM
(SURVEY.md §8f-2).

Reference path (``GaussianProcessRegression.log_likelihood``, ``models/gaussian_process_regression.py``):
``kernel.state_space_model(time_points)`` (``kernels/sde_kernel.py:153-171``) writes ``A_k``,
``chol Q_k``, ``b_k`` for every step, ``kernel.generate_emission_model`` (``:173-211``) writes ``H``,
and ``KalmanFilter.log_likelihood`` (``kalman_filter.py:184-255``) reads them back.  Here
``mf_kalman_matern_log_likelihood`` builds ``A_k, Q_k`` in registers from ``Δt_k`` (closed forms of
``kernels/matern.py:80-86,299-324,434-460`` and ``Q_k = P∞ − A_k P∞ A_kᵀ + jitter·I``,
``sde_kernel.py:421-446``): a step reads two values instead of ``2D²+2D+1``.

The kernels are mirrored as far as the operator API consumes them (the ``SDEKernel`` protocol below:
``state_space_model``, ``transition_statistics``, ``generate_emission_model``, initial moments) for
``Matern12/32/52``, ``HarmonicOscillator`` and ``Sum`` -- per-step closed forms, no recursion.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr
from .interop import as_torch, boundary, require_cuda
from .kalman_filter import _workspace


@boundary
def matern_kalman_log_likelihood(state_dim: int, lengthscale, variance, observations,
                                 chol_obs_covariance, time_points=None, time_deltas=None,
                                 jitter: float = 0.0, first_is_initial: bool = True,
                                 return_element: bool = False):
    """Per-series marginal log-likelihood ``batch_shape`` of ``y_k = x_k[0] + N(0, R)`` under a
    Matern-(``state_dim`` − ½) prior.

    ``lengthscale``, ``variance``: scalars or ``batch_shape`` tensors; ``observations``:
    ``batch_shape + [T]`` (or ``[..., T, 1]``); ``chol_obs_covariance``: scalar / ``[1,1]``;
    exactly one of ``time_points`` ``batch_shape + [T]`` and ``time_deltas`` ``batch_shape + [T-1]``
    (``[T]`` with ``first_is_initial=False``: a later time segment of a longer series, in which case
    only the scan element is meaningful).  ``return_element=True`` also returns the
    ``[..., 3D²+2D+1]`` scan element of the whole segment (time-sharded evaluation).
    """
    if state_dim not in (1, 2, 3):
        raise ValueError("state_dim must be 1 (Matern12), 2 (Matern32) or 3 (Matern52)")
    y = as_torch(observations)
    require_cuda(y, "observations")
    if (time_points is None) == (time_deltas is None):
        raise ValueError("pass exactly one of time_points and time_deltas")
    times = as_torch(time_points if time_points is not None else time_deltas, y.device)
    if y.dim() == times.dim() + 1 and y.shape[-1] == 1:  # [..., T, 1] as the reference stores it
        y = y[..., 0]
    batch = tuple(y.shape[:-1])
    t = int(y.shape[-1])
    if t < 1:
        raise ValueError("need at least one observation")
    dtype, dev = y.dtype, y.device
    if time_points is not None:
        tp = as_torch(time_points, dev).to(dtype)
        if tuple(tp.shape) != batch + (t,):
            raise ValueError(f"time_points must be {batch + (t,)}, got {tuple(tp.shape)}")
        if not first_is_initial:
            raise ValueError("a later time segment needs time_deltas (the delta INTO its first step)")
        dts = tp[..., 1:] - tp[..., :-1]
    else:
        dts = as_torch(time_deltas, dev).to(dtype)
    nt = t - (1 if first_is_initial else 0)
    if tuple(dts.shape) != batch + (nt,):
        raise ValueError(f"time_deltas must be {batch + (nt,)}, got {tuple(dts.shape)}")
    bsz = 1
    for s in batch:
        bsz *= int(s)

    def per_series(v, name):
        v = torch.as_tensor(v, dtype=dtype, device=dev)
        if v.dim() == 0:
            v = v.expand(batch) if batch else v
        if tuple(v.shape) != batch:
            raise ValueError(f"{name} must be a scalar or {batch}, got {tuple(v.shape)}")
        return v.reshape(bsz).contiguous()

    ls = per_series(lengthscale, "lengthscale")
    var = per_series(variance, "variance")
    lr = torch.as_tensor(chol_obs_covariance, dtype=dtype, device=dev).reshape(-1)
    if lr.numel() != 1:
        raise ValueError("chol_obs_covariance must hold one value (output_dim is 1)")
    lr = lr.contiguous()
    y2 = y.reshape(bsz, t).contiguous()
    dt2 = dts.reshape(bsz, nt).contiguous()
    d = state_dim
    n = 3 * d * d + 2 * d + 1
    out = torch.empty(bsz, dtype=dtype, device=dev)
    elem = torch.empty(bsz, n, dtype=dtype, device=dev) if (return_element or not first_is_initial) else None
    lib = _lib.lib()
    lib.mf_kalman_matern_workspace_bytes.restype = _lib.ctypes.c_size_t
    nbytes = int(lib.mf_kalman_matern_workspace_bytes(dtype_code(dtype), i64(bsz), i64(t), i64(d)))
    ws = _workspace(nbytes, dev)
    check(
        lib.mf_kalman_matern_log_likelihood(
            dtype_code(dtype), ptr(ls), ptr(var), _lib.ctypes.c_double(float(jitter)), ptr(dt2),
            ptr(y2), ptr(lr), ptr(out), ptr(elem), i64(bsz), i64(t), i64(d),
            _lib.ctypes.c_int(1 if first_is_initial else 0), ptr(ws),
            _lib.ctypes.c_size_t(nbytes), current_stream()),
        "mf_kalman_matern_log_likelihood",
    )
    ll = out.reshape(batch)
    if return_element or not first_is_initial:
        return ll, elem.reshape(batch + (n,))
    return ll


class SDEKernel:
    """The part of ``markovflow.kernels.SDEKernel`` / ``StationaryKernel`` (``kernels/sde_kernel.py:38-520``)
    that the operator API consumes: the kernel as a linear SDE ``dx = F x dt + L dB`` observed through
    ``H = [1, 0, ...]``, i.e. a recipe for state-space models on arbitrary time points.  Closed forms
    are per-step elementwise maps (no recursion) and stay in torch; everything downstream
    (``state_space_model(...)``, ``ConditionalProcess``, Kalman filters) runs on the CUDA operators.

    Subclasses provide ``state_dim``, ``steady_state_covariance`` ``[D,D]`` and
    ``state_transitions(transition_times, time_deltas)``; hyper-parameters may be tensors that
    require gradients (the closed forms are differentiable)."""

    output_dim: int = 1
    jitter: float = 0.0

    # -- to be provided ------------------------------------------------------------------------
    state_dim: int = 0

    def _like(self, ref=None):
        for v in getattr(self, "_hyper", ()):
            if isinstance(v, torch.Tensor):
                return dict(dtype=v.dtype, device=v.device)
        if isinstance(ref, torch.Tensor):
            return dict(dtype=ref.dtype if ref.is_floating_point() else torch.float64, device=ref.device)
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        return dict(dtype=torch.float64, device=dev)

    def _pinf(self, ref=None) -> torch.Tensor:
        """``steady_state_covariance`` on the dtype / device of ``ref`` (python-float hyper-parameters)."""
        raise NotImplementedError

    @property
    def steady_state_covariance(self) -> torch.Tensor:
        return self._pinf(None)

    def state_transitions(self, transition_times, time_deltas) -> torch.Tensor:
        raise NotImplementedError

    # -- reference API (sde_kernel.py:153-520) ---------------------------------------------------
    def _jitter(self, ref) -> torch.Tensor:
        return self.jitter * torch.eye(self.state_dim, dtype=ref.dtype, device=ref.device)

    @property
    def jitter_matrix(self) -> torch.Tensor:
        return self._jitter(self.steady_state_covariance)

    def initial_mean(self, batch_shape, ref=None) -> torch.Tensor:
        kw = self._like(ref)
        return torch.zeros(tuple(batch_shape) + (self.state_dim,), **kw)

    def initial_covariance(self, initial_time_point) -> torch.Tensor:
        t0 = as_torch(initial_time_point)
        pinf = self._pinf(t0)
        pinf = pinf + self._jitter(pinf)
        return pinf.expand(tuple(t0.shape[:-1]) + (self.state_dim, self.state_dim))

    def state_offsets(self, transition_times, time_deltas) -> torch.Tensor:
        dt = as_torch(time_deltas)
        return torch.zeros(tuple(dt.shape) + (self.state_dim,), dtype=dt.dtype, device=dt.device)

    def transition_statistics(self, transition_times, time_deltas):
        """``(A_k, Q_k)`` with ``Q_k = Pinf - A_k Pinf A_k^T + jitter I`` (``sde_kernel.py:421-446``)."""
        a = self.state_transitions(transition_times, time_deltas)
        pinf = self._pinf(a).to(a.dtype)
        q = pinf - a @ pinf @ a.transpose(-1, -2)
        return a, q + self._jitter(a)

    def transition_statistics_from_time_points(self, time_points):
        tp = as_torch(time_points)
        return self.transition_statistics(tp[..., :-1], tp[..., 1:] - tp[..., :-1])

    def state_space_model(self, time_points):
        """``sde_kernel.py:153-171``: the prior on ``time_points`` as a :class:`StateSpaceModel`."""
        from .state_space_model import state_space_model_from_covariances

        tp = as_torch(time_points)
        a, q = self.transition_statistics_from_time_points(tp)
        batch = tuple(tp.shape[:-1])
        return state_space_model_from_covariances(
            initial_mean=self.initial_mean(batch, a).to(a.dtype),
            initial_covariance=self.initial_covariance(tp[..., :1]).to(a.dtype).contiguous(),
            state_transitions=a,
            state_offsets=self.state_offsets(tp[..., :-1], tp[..., 1:] - tp[..., :-1]).to(a.dtype),
            process_covariances=q)

    def build_finite_distribution(self, time_points):
        return self.state_space_model(time_points)

    def generate_emission_model(self, time_points):
        """``H = [1, 0, ...]`` tiled over the time points (``sde_kernel.py:173-211``)."""
        from .emission_model import EmissionModel

        tp = as_torch(time_points)
        row = self._emission_row().to(tp.dtype if tp.is_floating_point() else torch.float64).to(tp.device)
        return EmissionModel(row.expand(tuple(tp.shape) + (self.output_dim, self.state_dim)).contiguous())

    def _emission_row(self) -> torch.Tensor:
        h = torch.zeros(self.output_dim, self.state_dim, dtype=torch.float64)
        h[:, 0] = 1.0
        return h


class _Matern(SDEKernel):
    """Common part of the mirrored Matern kernels (reference ``kernels/matern.py``): hyper-parameters,
    the closed-form SDE statistics and the fused Kalman log-likelihood."""

    state_dim: int = 0

    def __init__(self, lengthscale, variance, output_dim: int = 1, jitter: float = 0.0) -> None:
        if output_dim != 1:
            raise ValueError("the fused path covers output_dim = 1")
        self.lengthscale = lengthscale
        self.variance = variance
        self.jitter = float(jitter)
        self.output_dim = output_dim
        self._hyper = (lengthscale, variance)

    def _lam(self, ref=None) -> torch.Tensor:
        kw = self._like(ref)
        ls = torch.as_tensor(self.lengthscale, **kw) if not isinstance(self.lengthscale, torch.Tensor) else self.lengthscale
        return math.sqrt(2.0 * self.state_dim - 1.0) / ls

    def _var(self, ref=None) -> torch.Tensor:
        kw = self._like(ref)
        return torch.as_tensor(self.variance, **kw) if not isinstance(self.variance, torch.Tensor) else self.variance

    @property
    def feedback_matrix(self) -> torch.Tensor:
        return self._feedback(None)

    def _feedback(self, ref) -> torch.Tensor:
        """Companion matrix of ``(s + lam)^D`` (``kernels/matern.py:60-78,271-297,407-432``)."""
        lam, d = self._lam(ref), self.state_dim
        f = torch.zeros(d, d, dtype=lam.dtype, device=lam.device)
        for i in range(d - 1):
            f[i, i + 1] = 1.0
        coef = [math.comb(d, i) for i in range(d)]
        row = torch.stack([-coef[i] * lam ** (d - i) for i in range(d)])
        return torch.cat([f[:-1], row[None]], dim=0)

    def _pinf(self, ref=None) -> torch.Tensor:
        lam, var = self._lam(ref), self._var(ref)
        one, zero = torch.ones_like(lam), torch.zeros_like(lam)
        if self.state_dim == 1:
            rows = [[one]]
        elif self.state_dim == 2:
            rows = [[one, zero], [zero, lam * lam]]
        else:
            l2 = lam * lam
            rows = [[one, zero, -l2 / 3.0], [zero, l2 / 3.0, zero], [-l2 / 3.0, zero, l2 * l2]]
        return var * torch.stack([torch.stack(r) for r in rows])

    def state_transitions(self, transition_times, time_deltas) -> torch.Tensor:
        """``A = exp(-lam dt) (I + N dt + N^2 dt^2 / 2)`` with ``N = F + lam I`` nilpotent
        (``kernels/matern.py:80-86,299-324,434-460``)."""
        dt = as_torch(time_deltas)
        lam = self._lam(dt).to(dt.dtype)
        d = self.state_dim
        eye = torch.eye(d, dtype=dt.dtype, device=dt.device)
        n_mat = self._feedback(dt).to(dt.dtype) + lam * eye
        x = n_mat * dt[..., None, None]
        poly = eye + x
        if d == 3:
            poly = poly + x @ x / 2.0
        return torch.exp(-lam * dt)[..., None, None] * poly

    @boundary
    def kalman_log_likelihood_per_chain(self, time_points, observations, chol_obs_covariance):
        return matern_kalman_log_likelihood(self.state_dim, self.lengthscale, self.variance,
                                            observations, chol_obs_covariance,
                                            time_points=time_points, jitter=self.jitter)

    @boundary
    def kalman_log_likelihood(self, time_points, observations, chol_obs_covariance) -> torch.Tensor:
        """``KalmanFilter(kernel.state_space_model(t), kernel.generate_emission_model(t), R, y)
        .log_likelihood()`` (reference ``kalman_filter.py:184-255``: summed over the batch)."""
        return torch.sum(self.kalman_log_likelihood_per_chain(time_points, observations,
                                                              chol_obs_covariance))


class Matern12(_Matern):
    """``kernels/matern.py:27-143``."""
    state_dim = 1


class Matern32(_Matern):
    """``kernels/matern.py:237-372``."""
    state_dim = 2


class Matern52(_Matern):
    """``kernels/matern.py:376-501``."""
    state_dim = 3


class HarmonicOscillator(SDEKernel):
    """``kernels/periodic.py:27-187``: ``k(r) = variance cos(2 pi r / period)``; the state rotates,
    ``Pinf = variance I`` and ``Q_k = jitter I``."""

    state_dim = 2

    def __init__(self, variance, period, output_dim: int = 1, jitter: float = 0.0) -> None:
        self.variance, self.period = variance, period
        self.jitter, self.output_dim = float(jitter), output_dim
        self._hyper = (variance, period)

    def _pinf(self, ref=None) -> torch.Tensor:
        kw = self._like(ref)
        var = self.variance if isinstance(self.variance, torch.Tensor) else torch.as_tensor(self.variance, **kw)
        return var * torch.eye(2, dtype=var.dtype, device=var.device)

    def state_transitions(self, transition_times, time_deltas) -> torch.Tensor:
        dt = as_torch(time_deltas)
        per = self.period if isinstance(self.period, torch.Tensor) else torch.as_tensor(self.period, dtype=dt.dtype, device=dt.device)
        ang = dt * (2.0 * math.pi / per)
        c, s = torch.cos(ang), torch.sin(ang)
        return torch.stack([torch.stack([c, -s], dim=-1), torch.stack([s, c], dim=-1)], dim=-2)


class Sum(SDEKernel):
    """``kernels/sde_kernel.py:540-687``: independent components, block-diagonal state, summed outputs."""

    def __init__(self, kernels, jitter: float = 0.0) -> None:
        self.kernels = list(kernels)
        self.jitter = float(jitter)
        self.output_dim = self.kernels[0].output_dim
        self.state_dim = sum(k.state_dim for k in self.kernels)
        self._hyper = tuple(h for k in self.kernels for h in getattr(k, "_hyper", ()))

    def _pinf(self, ref=None) -> torch.Tensor:
        return torch.block_diag(*(k._pinf(ref) for k in self.kernels))

    def state_transitions(self, transition_times, time_deltas) -> torch.Tensor:
        dt = as_torch(time_deltas)
        blocks = [k.state_transitions(transition_times, dt) for k in self.kernels]
        out = torch.zeros(tuple(dt.shape) + (self.state_dim, self.state_dim), dtype=blocks[0].dtype, device=dt.device)
        o = 0
        for k, blk in zip(self.kernels, blocks):
            out[..., o:o + k.state_dim, o:o + k.state_dim] = blk
            o += k.state_dim
        return out

    def transition_statistics(self, transition_times, time_deltas):
        """Block-diagonal of the components' statistics plus this kernel's jitter (``:592-640``)."""
        dt = as_torch(time_deltas)
        parts = [k.transition_statistics(transition_times, dt) for k in self.kernels]
        a = torch.zeros(tuple(dt.shape) + (self.state_dim, self.state_dim), dtype=parts[0][0].dtype, device=dt.device)
        q = torch.zeros_like(a)
        o = 0
        for k, (ak, qk) in zip(self.kernels, parts):
            a[..., o:o + k.state_dim, o:o + k.state_dim] = ak
            q[..., o:o + k.state_dim, o:o + k.state_dim] = qk
            o += k.state_dim
        return a, q + self._jitter(a)

    def _emission_row(self) -> torch.Tensor:
        h = torch.zeros(self.output_dim, self.state_dim, dtype=torch.float64)
        o = 0
        for k in self.kernels:
            h[:, o] = 1.0
            o += k.state_dim
        return h
