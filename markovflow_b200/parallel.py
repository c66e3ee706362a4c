"""Multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed``).

* Batches of independent chains (BASELINE configs 2, 4, 5) are split over ranks with NO data-path
  collective (``batch_shape`` dims are pure broadcasting, ``block_tri_diag.py:110-115``):
  :func:`shard_bounds` / :func:`shard_batch`.
* ONE long series (config 3) is split in time.  Each rank reduces its segment -- ONE pass over its
  data -- to one scan element ``(A, b, C, eta, J, ell)`` (``mf_kalman_segment_summary``), the
  elements are all-gathered (NCCL over NVLink; 8 x 136 B at D = 2) and every rank joins them in time
  order (``mf_kalman_fold_elements``): the ``ell`` of the join is the log-likelihood of the whole
  series.  No second sweep and no further collective.  (``mf_kalman_log_likelihood_seeded`` -- a
  local filter seeded with the prefix of the earlier ranks -- remains available for callers that
  want per-segment shares.)

The compute engine is pluggable so that the host logic (slicing, gather order, fold range) is
testable on CPU with ``gloo``; the default engine is the CUDA library and there is no other engine
in this package.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import check, current_stream, dtype_code, i64, ptr


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[lo, hi)`` of ``n`` units for ``rank`` (first ``n % world`` ranks get
    one extra)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int, dim: int = 0) -> torch.Tensor:
    """This rank's slice of a batch of independent chains."""
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


@dataclass
class TimeSegment:
    """One rank's contiguous time segment of a (batch of) long series, in the kernels' "incoming
    transition" convention: ``first`` segments carry ``(mu0, chol_p0)`` and ``T-1`` transitions,
    later segments carry ``T`` transitions (the first one leads into local step 0)."""
    first: bool
    mu0: Optional[torch.Tensor]
    chol_p0: Optional[torch.Tensor]
    a: torch.Tensor
    b: torch.Tensor
    chol_q: torch.Tensor
    h: torch.Tensor       # [Bh,Tl,m,D]
    obs: torch.Tensor     # [B,Tl,m]
    chol_r: torch.Tensor  # [1 or Tl,m,m]

    @property
    def num_steps(self) -> int:
        return int(self.obs.shape[-2])


def time_segment(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, rank: int, world: int) -> TimeSegment:
    """Slice rank ``rank``'s segment out of full-series arrays ``a [B,T-1,D,D]`` ... ``obs [B,T,m]``."""
    t = int(obs.shape[-2])
    lo, hi = shard_bounds(t, rank, world)
    if hi - lo < 1:
        raise ValueError("every rank needs at least one time step")
    first = lo == 0
    tlo = 0 if first else lo - 1  # transition k leads into step k+1
    thi = hi - 1
    r = chol_r if chol_r.shape[0] == 1 else chol_r[lo:hi]
    return TimeSegment(
        first, mu0 if first else None, chol_p0 if first else None,
        a[:, tlo:thi].contiguous(), b[:, tlo:thi].contiguous(), chol_q[:, tlo:thi].contiguous(),
        h[:, lo:hi].contiguous(), obs[:, lo:hi].contiguous(), r.contiguous())


class CudaKalmanEngine:
    """The three C-ABI calls of the time-sharded protocol (``include/markovflow_b200.h``)."""

    def __init__(self) -> None:
        self._ws = None
        self._ws_bytes = 0

    @staticmethod
    def elem_size(d: int) -> int:
        return 3 * d * d + 2 * d + 1

    def _dims(self, seg: TimeSegment):
        bsz, tl, m = seg.obs.shape
        d = int(seg.a.shape[-1])
        return int(bsz), int(tl), int(m), d, int(seg.h.shape[0]), int(seg.chol_r.shape[0])

    def _workspace(self, seg: TimeSegment):
        bsz, tl, m, d, _, _ = self._dims(seg)
        lib = _lib.lib()
        lib.mf_kalman_workspace_bytes.restype = _lib.ctypes.c_size_t
        n = int(lib.mf_kalman_workspace_bytes(dtype_code(seg.obs.dtype), i64(bsz), i64(tl), i64(d)))
        if self._ws is None or self._ws_bytes < n or self._ws.device != seg.obs.device:
            self._ws = torch.empty(max(n, 1), dtype=torch.uint8, device=seg.obs.device)
            self._ws_bytes = n
        return self._ws, n

    def segment_summary(self, seg: TimeSegment) -> torch.Tensor:
        bsz, tl, m, d, hb, rs = self._dims(seg)
        ws, n = self._workspace(seg)
        out = torch.empty(bsz, self.elem_size(d), dtype=seg.obs.dtype, device=seg.obs.device)
        check(
            _lib.lib().mf_kalman_segment_summary(
                dtype_code(seg.obs.dtype), ptr(seg.mu0), ptr(seg.chol_p0), ptr(seg.a), ptr(seg.b),
                ptr(seg.chol_q), ptr(seg.h), ptr(seg.obs), ptr(seg.chol_r), ptr(out), i64(bsz),
                i64(tl), i64(d), i64(m), i64(hb), i64(rs), int(seg.first), ptr(ws),
                _lib.ctypes.c_size_t(n), current_stream()),
            "mf_kalman_segment_summary",
        )
        return out

    def fold(self, elems: torch.Tensor, d: int) -> torch.Tensor:
        """``elems [n,B,N]`` (time order) -> their join ``[B,N]``."""
        n, bsz, _ = elems.shape
        elems = elems.contiguous()
        out = torch.empty(bsz, elems.shape[-1], dtype=elems.dtype, device=elems.device)
        check(
            _lib.lib().mf_kalman_fold_elements(
                dtype_code(elems.dtype), ptr(elems), ptr(out), i64(n), i64(bsz), i64(d),
                current_stream()),
            "mf_kalman_fold_elements",
        )
        return out

    def seeded_log_likelihood(self, seg: TimeSegment, prefix: Optional[torch.Tensor],
                              summaries_valid: bool) -> torch.Tensor:
        bsz, tl, m, d, hb, rs = self._dims(seg)
        ws, n = self._workspace(seg)
        out = torch.empty(bsz, dtype=seg.obs.dtype, device=seg.obs.device)
        check(
            _lib.lib().mf_kalman_log_likelihood_seeded(
                dtype_code(seg.obs.dtype), ptr(seg.mu0), ptr(seg.chol_p0), ptr(seg.a), ptr(seg.b),
                ptr(seg.chol_q), ptr(seg.h), ptr(seg.obs), ptr(seg.chol_r), ptr(prefix), ptr(out),
                i64(bsz), i64(tl), i64(d), i64(m), i64(hb), i64(rs), int(seg.first),
                int(bool(summaries_valid)), ptr(ws), _lib.ctypes.c_size_t(n), current_stream()),
            "mf_kalman_log_likelihood_seeded",
        )
        return out


class PeerRing:
    """The peer-mapped exchange regions of a time-sharded evaluation (``mf_peer_*``, ``include/markovflow_b200.h``).

    Every rank allocates one region (plain ``cudaMalloc``), exports its CUDA IPC handle, and maps the regions of
    all other ranks of the node (NVLink / NVSwitch peer-to-peer).  The handles travel once through
    ``all_gather_object``; afterwards a time-sharded log-likelihood is ONE library call per rank
    (:meth:`kalman` / :meth:`matern`): the kernel that reduces the rank's segment writes its element into every
    rank's region, raises flags, waits for the others' and joins the elements in rank order -- no collective."""

    def __init__(self, dtype: torch.dtype, batch: int, state_dim: int, group=None, regions=None, rank=None,
                 world=None) -> None:
        import torch.distributed as dist

        self.lib = _lib.lib()
        self.dtype, self.batch, self.state_dim = dtype, int(batch), int(state_dim)
        self.lib.mf_kalman_peer_region_bytes.restype = _lib.ctypes.c_size_t
        self._own = self._opened = None
        if regions is not None:  # regions supplied by the caller (tests: virtual ranks on one device)
            self.rank, self.world = int(rank), int(world)
            self._ptrs = [int(p) for p in regions]
        else:
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            nbytes = self.region_bytes(dtype, batch, state_dim, self.world)
            own = _lib.ctypes.c_void_p()
            handle = (_lib.ctypes.c_ubyte * 64)()
            check(self.lib.mf_peer_alloc(_lib.ctypes.c_size_t(nbytes), _lib.ctypes.byref(own), handle), "mf_peer_alloc")
            self._own = own
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._ptrs, self._opened = [], []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    self._ptrs.append(int(own.value))
                    continue
                p = _lib.ctypes.c_void_p()
                check(self.lib.mf_peer_open((_lib.ctypes.c_ubyte * 64).from_buffer_copy(hb), _lib.ctypes.byref(p)),
                      "mf_peer_open")
                self._opened.append(p)
                self._ptrs.append(int(p.value))
            dist.barrier(group=group)
        self._arr = (_lib.ctypes.c_void_p * self.world)(*self._ptrs)
        self.epoch = 0

    @staticmethod
    def region_bytes(dtype, batch: int, state_dim: int, world: int) -> int:
        lib = _lib.lib()
        lib.mf_kalman_peer_region_bytes.restype = _lib.ctypes.c_size_t
        return int(lib.mf_kalman_peer_region_bytes(dtype_code(dtype), i64(batch), i64(state_dim), int(world)))

    def close(self) -> None:
        for p in self._opened or []:
            self.lib.mf_peer_close(p)
        if self._own is not None:
            self.lib.mf_peer_free(self._own)
        self._opened, self._own = None, None

    def _next(self) -> int:
        self.epoch += 1
        return self.epoch

    def kalman(self, seg: TimeSegment, engine: Optional["CudaKalmanEngine"] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(log-likelihood [B] of the whole series, its joined scan element [B,N])`` -- collective: every rank
        calls it with its own segment (rank order == time order)."""
        engine = engine or CudaKalmanEngine()
        bsz, tl, m, d, hb, rs = engine._dims(seg)
        ws, n = engine._workspace(seg)
        out = torch.empty(bsz, dtype=seg.obs.dtype, device=seg.obs.device)
        elem = torch.empty(bsz, engine.elem_size(d), dtype=seg.obs.dtype, device=seg.obs.device)
        check(
            self.lib.mf_kalman_time_sharded_log_likelihood(
                dtype_code(seg.obs.dtype), ptr(seg.mu0), ptr(seg.chol_p0), ptr(seg.a), ptr(seg.b), ptr(seg.chol_q),
                ptr(seg.h), ptr(seg.obs), ptr(seg.chol_r), ptr(out), ptr(elem), i64(bsz), i64(tl), i64(d), i64(m),
                i64(hb), i64(rs), int(seg.first), self._arr, int(self.rank), int(self.world),
                _lib.ctypes.c_uint64(self._next()), ptr(ws), _lib.ctypes.c_size_t(n), current_stream()),
            "mf_kalman_time_sharded_log_likelihood",
        )
        return out, elem

    def matern(self, state_dim: int, lengthscale, variance, seg_deltas, seg_obs, chol_obs_covariance, first: bool,
               jitter: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same for a Matern prior with the state-space model built in the kernel from the time deltas."""
        y = seg_obs.contiguous()
        bsz, t = y.shape
        dt = seg_deltas.contiguous()
        dtype, dev = y.dtype, y.device
        ls = torch.as_tensor(lengthscale, dtype=dtype, device=dev).expand(bsz).contiguous()
        var = torch.as_tensor(variance, dtype=dtype, device=dev).expand(bsz).contiguous()
        lr = torch.as_tensor(chol_obs_covariance, dtype=dtype, device=dev).reshape(-1)[:1].contiguous()
        self.lib.mf_kalman_matern_workspace_bytes.restype = _lib.ctypes.c_size_t
        nbytes = int(self.lib.mf_kalman_matern_workspace_bytes(dtype_code(dtype), i64(bsz), i64(t), i64(state_dim)))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        out = torch.empty(bsz, dtype=dtype, device=dev)
        elem = torch.empty(bsz, 3 * state_dim * state_dim + 2 * state_dim + 1, dtype=dtype, device=dev)
        check(
            self.lib.mf_kalman_matern_time_sharded_log_likelihood(
                dtype_code(dtype), ptr(ls), ptr(var), _lib.ctypes.c_double(float(jitter)), ptr(dt), ptr(y), ptr(lr),
                ptr(out), ptr(elem), i64(bsz), i64(t), i64(state_dim), int(bool(first)), self._arr, int(self.rank),
                int(self.world), _lib.ctypes.c_uint64(self._next()), ptr(ws), _lib.ctypes.c_size_t(nbytes),
                current_stream()),
            "mf_kalman_matern_time_sharded_log_likelihood",
        )
        return out, elem


def time_sharded_log_likelihood(seg: TimeSegment, group=None, engine=None, ring: Optional[PeerRing] = None
                                ) -> torch.Tensor:
    """Per-chain log-likelihood ``[B]`` of the whole series, computed collectively: every rank
    passes its own :class:`TimeSegment` (rank order == time order) and receives the same value.
    With a :class:`PeerRing` the exchange happens inside the reduction kernel over peer memory (one library
    call, no collective); without one: NCCL all-gather of the elements + a fold launch."""
    import torch.distributed as dist

    engine = engine or CudaKalmanEngine()
    if ring is not None:
        if seg.first != (ring.rank == 0):
            raise ValueError("rank 0 (and only rank 0) must hold the segment that starts at the prior")
        return ring.kalman(seg, engine)[0]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if seg.first != (rank == 0):
        raise ValueError("rank 0 (and only rank 0) must hold the segment that starts at the prior")
    d = int(seg.a.shape[-1])
    elem = engine.segment_summary(seg)
    if world > 1:
        gathered = [torch.empty_like(elem) for _ in range(world)]
        dist.all_gather(gathered, elem, group=group)
        elem = engine.fold(torch.stack(gathered), d)
    return elem[:, -1].clone()


def time_sharded_segments(ssm, emission_matrix, observations, chol_obs_covariance,
                          world: int) -> List[TimeSegment]:
    """All ``world`` segments of a (batch of) series held on one device."""
    mu0, l0, a, b, lq, bsz, t, d = ssm._flat()
    m = int(emission_matrix.shape[-2])
    h = emission_matrix.reshape(-1, t, m, d)
    y = observations.reshape(bsz, t, m)
    lr = chol_obs_covariance.reshape(-1, m, m)
    return [time_segment(mu0, l0, a, b, lq, h, y, lr, r, world) for r in range(world)]


def time_sharded_log_likelihood_local(ssm, emission_matrix, observations, chol_obs_covariance,
                                      world: int, engine=None, seeded: bool = False) -> torch.Tensor:
    """The same protocol executed for ``world`` virtual ranks on ONE device (validation, and the
    single-GPU leg of the scaling bench): returns the per-chain log-likelihood ``[batch]``.
    ``seeded=True`` exercises the alternative protocol (prefix fold + seeded local filters)."""
    engine = engine or CudaKalmanEngine()
    d = ssm.state_dim
    segs = time_sharded_segments(ssm, emission_matrix, observations, chol_obs_covariance, world)
    engines = [engine.__class__() for _ in segs]  # one workspace per virtual rank
    elems = [e.segment_summary(s) for e, s in zip(engines, segs)]
    if not seeded:
        return engine.fold(torch.stack(elems), d)[:, -1].reshape(tuple(ssm.batch_shape))
    total = None
    for r, (e, s) in enumerate(zip(engines, segs)):
        prefix = e.fold(torch.stack(elems[:r]), d) if r > 0 else None
        share = e.seeded_log_likelihood(s, prefix, summaries_valid=True)
        total = share if total is None else total + share
    return total.reshape(tuple(ssm.batch_shape))


# ---- Matern prior with the SSM built in the kernel -------------------------------

def matern_time_segment(time_deltas: torch.Tensor, observations: torch.Tensor, rank: int, world: int):
    """Rank ``rank``'s slice of ``time_deltas [B,T-1]`` / ``observations [B,T]``:
    ``(first, deltas, obs)`` in the "incoming delta" convention of ``mf_kalman_matern_log_likelihood``."""
    t = int(observations.shape[-1])
    lo, hi = shard_bounds(t, rank, world)
    if hi - lo < 1:
        raise ValueError("every rank needs at least one time step")
    first = lo == 0
    tlo = 0 if first else lo - 1
    return first, time_deltas[:, tlo:hi - 1].contiguous(), observations[:, lo:hi].contiguous()


def time_sharded_matern_log_likelihood(state_dim: int, lengthscale, variance, seg_deltas, seg_obs,
                                       chol_obs_covariance, first: bool, jitter: float = 0.0,
                                       group=None, engine=None) -> torch.Tensor:
    """Collective per-series log-likelihood ``[B]`` of a long series split in time over the ranks
    (rank order == time order): local scan element -> all-gather -> ordered fold, as
    :func:`time_sharded_log_likelihood`, with the SSM built inside the kernel from the deltas."""
    import torch.distributed as dist

    from .kernels import matern_kalman_log_likelihood

    engine = engine or CudaKalmanEngine()
    world = dist.get_world_size(group)
    if first != (dist.get_rank(group) == 0):
        raise ValueError("rank 0 (and only rank 0) must hold the segment that starts at the prior")
    _, elem = matern_kalman_log_likelihood(state_dim, lengthscale, variance, seg_obs,
                                           chol_obs_covariance, time_deltas=seg_deltas, jitter=jitter,
                                           first_is_initial=first, return_element=True)
    if world > 1:
        gathered = [torch.empty_like(elem) for _ in range(world)]
        dist.all_gather(gathered, elem, group=group)
        elem = engine.fold(torch.stack(gathered), state_dim)
    return elem[:, -1].clone()
