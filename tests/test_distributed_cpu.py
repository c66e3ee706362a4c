"""Host-side multi-GPU logic on CPU: ``gloo`` backend, world_size 2 (and 3).

The CUDA kernels cannot run here, so the three compute calls of the time-sharded protocol are
served by an engine built on the numpy oracle (test infrastructure); everything else -- slicing of
the series into per-rank segments, the all-gather, which gathered elements a rank folds, the final
all-reduce, batch sharding bounds -- is the product code of ``markovflow_b200/parallel.py``."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle as O


class OracleKalmanEngine:
    """numpy stand-in for ``CudaKalmanEngine`` (single chain, m = 1)."""

    @staticmethod
    def _np(seg):
        a, b, lq = (x[0].numpy() for x in (seg.a, seg.b, seg.chol_q))
        q = lq @ np.swapaxes(lq, -1, -2)
        h, y = seg.h[0].numpy(), seg.obs[0].numpy()
        r = seg.chol_r[0].numpy() @ seg.chol_r[0].numpy().T
        return a, b, q, h, y, r

    def segment_summary(self, seg):
        a, b, q, h, y, r = self._np(seg)
        prior = None
        if seg.first:
            l0 = seg.chol_p0[0].numpy()
            prior = (seg.mu0[0].numpy(), l0 @ l0.T)
        e = O.pscan_segment_element(a, b, q, h, r, y, prior=prior)
        return self._pack(e)

    @staticmethod
    def _pack(e):
        return torch.from_numpy(np.concatenate([np.reshape(x, -1) for x in e]))[None]

    @staticmethod
    def _unpack(v, d):
        v = v.numpy()
        dd = d * d
        return (v[:dd].reshape(d, d), v[dd:dd + d], v[dd + d:2 * dd + d].reshape(d, d),
                v[2 * dd + d:2 * dd + 2 * d], v[2 * dd + 2 * d:3 * dd + 2 * d].reshape(d, d),
                float(v[3 * dd + 2 * d]))

    def fold(self, elems, d):
        e = self._unpack(elems[0, 0], d)
        for i in range(1, elems.shape[0]):
            e = O.pscan_combine_ell(e, self._unpack(elems[i, 0], d))
        return self._pack(e)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(t=61, seed=3):
    rng = np.random.default_rng(seed)
    k = O.Matern32(1.1, 0.9)
    tp = np.cumsum(rng.uniform(0.05, 0.3, size=t))
    ssm = k.state_space_model(tp)
    h = k.emission_matrix(tp)
    y = rng.standard_normal((t, 1))
    lr = np.array([[0.3]])
    return ssm, h, y, lr


def _worker(rank, world, port, t, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from markovflow_b200.parallel import time_segment, time_sharded_log_likelihood

        ssm, h, y, lr = _case(t)
        f = lambda x: torch.from_numpy(np.ascontiguousarray(x))[None]
        seg = time_segment(f(ssm.mu0), f(ssm.chol_p0), f(ssm.a_s), f(ssm.b_s), f(ssm.chol_q_s),
                           f(h), f(y), f(lr), rank, world)
        ll = time_sharded_log_likelihood(seg, engine=OracleKalmanEngine())
        out[rank] = float(ll[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,t", [(2, 61), (3, 50)])
def test_time_sharded_log_likelihood_gloo(world, t):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, t, out), nprocs=world, join=True)
        got = [out[r] for r in range(world)]
    ssm, h, y, lr = _case(t)
    want = float(O.kalman_log_likelihood(ssm, h, y, O._r_inv_from_chol(lr)))
    for g in got:  # every rank folds the same gathered elements
        assert abs(g - want) < 1e-9 * abs(want)


def test_shard_bounds_cover_and_balance():
    from markovflow_b200.parallel import shard_batch, shard_bounds

    for n in (1, 7, 8, 4096, 10 ** 7):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    x = torch.arange(10)
    assert torch.equal(torch.cat([shard_batch(x, r, 3) for r in range(3)]), x)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_time_segments_tile_the_series():
    from markovflow_b200.parallel import time_segment

    t, d = 23, 2
    a = torch.arange((t - 1) * d * d, dtype=torch.float64).reshape(1, t - 1, d, d)
    b = torch.arange((t - 1) * d, dtype=torch.float64).reshape(1, t - 1, d)
    h = torch.zeros(1, t, 1, d, dtype=torch.float64)
    y = torch.arange(t, dtype=torch.float64).reshape(1, t, 1)
    lr = torch.ones(1, 1, 1, dtype=torch.float64)
    segs = [time_segment(torch.zeros(1, d), torch.eye(d)[None], a, b, a, h, y, lr, r, 4) for r in range(4)]
    assert segs[0].first and not any(s.first for s in segs[1:])
    assert torch.equal(torch.cat([s.obs for s in segs], dim=1), y)
    # first segment: T-1 transitions; later ones: T transitions, the first leading into their step 0
    assert segs[0].a.shape[1] == segs[0].num_steps - 1
    for s in segs[1:]:
        assert s.a.shape[1] == s.num_steps
    assert torch.equal(torch.cat([s.a for s in segs], dim=1), a)


def test_matern_time_segments_tile_the_series():
    """Host-side slicing of the in-kernel-SSM path: segments cover every step once; a later segment
    carries the delta leading INTO its first step."""
    import torch

    from markovflow_b200.parallel import matern_time_segment

    t = 23
    dts = torch.arange(1, t, dtype=torch.float64)[None]  # delta k leads into step k+1
    obs = torch.arange(t, dtype=torch.float64)[None]
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            first, d, y = matern_time_segment(dts, obs, r, world)
            assert first == (r == 0)
            assert d.shape[-1] == y.shape[-1] - (1 if first else 0)
            into = d[0] if first else d[0, 1:]
            assert torch.equal(into, y[0, 1:])  # delta into step k is k by construction
            if not first:
                assert d[0, 0] == y[0, 0]
            seen.append(y[0])
        assert torch.equal(torch.cat(seen), obs[0])
