#!/usr/bin/env python
"""Benchmark of BASELINE.json's metric: "Cholesky+solve & Kalman log-lik state-steps/s; % HBM peak".

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W  (CPU reference arm)

Headline line (``value``, ``roofline``, ``e2e``): config 2 -- Matern52 (D=3) posterior precision,
B=4096 independent series per GPU x T=10,000 states, float64; one *step* = one fused Cholesky +
forward-solve sweep over the whole batch.  Chains are independent, so N GPUs hold N x 4096 series (weak
scaling, no collective on the data path).  ``e2e`` = the same work through the host-buffer C ABI
(``mf_host_btd_cholesky``, pinned HOST arrays, copies inside the timed region).

The other half of the metric and the other configurations travel in the driver-kept objects of the same
line: ``roofline.by_config`` (config 3 Kalman log-likelihood, config 4 D=17, config 5 transforms, config 1:
state-steps/s, ms, fraction of the HBM roofline, parity against the oracle / C port AT THE NAMED SIZE),
``cpu_baseline.by_config`` (the C port of the reference's CPU path for each of them, core count stated)
and ``e2e.kalman`` (config 3 through ``mf_host_kalman_log_likelihood``).  ``extra`` repeats them with
details.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_JSON_OUT = sys.stdout

B_PER_GPU, T, D = 4096, 10000, 3
ALGO_BYTES_PER_STEP = (4 * D * D + 2 * D) * 8  # read diag+sub+rhs, write Ld+Ls+x  
METRIC = "block_tridiag_cholesky_solve_state_steps_per_s"
UNIT = "state-steps/s"
WORKLOAD = "config2: Matern52 D=3 posterior precision, B=4096 series/GPU x T=10000, float64 Cholesky+solve"


def config_dict(world: int) -> dict:
    """``config`` of the JSON line -- the SAME object in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "T": T, "D": D,
            "parallelism": f"batch-sharded x{world}, no collective",
            "l2": "inputs (6.9 GB) and outputs (6.9 GB) per step exceed the 126 MB L2"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 6:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "samples": len(self.rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------------
# CPU legs: the C port of the reference's CPU path (oracle/banded_ref.c, oracle/ssm_ref.c) on host cores
# ------------------------------------------------------------------------------------------------------
def host_cores() -> int:
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _time_cpu(fn, reps: int, warm: int = 1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)


def cpu_reference_run(steps: int, warmup: int, sample_chains: int):
    """Config 2: time the C port of the reference path on a bounded sample of the workload."""
    import numpy as np

    from oracle import c_ref, np_oracle as O

    rng = np.random.default_rng(71892305)
    diags, subs = [], []
    proto = 16  # distinct chains built through the oracle's Matern52 restatement, then tiled
    for _ in range(proto):
        ell, var = rng.uniform(0.5, 2.0), rng.uniform(0.5, 2.0)
        tp = np.cumsum(ell * rng.uniform(0.2, 1.0, size=T))
        k = O.Matern52(ell, var)
        dg, sb = O.kalman_k_inv_post(k.state_space_model(tp), k.emission_matrix(tp), np.array([[100.0]]))
        diags.append(dg)
        subs.append(sb)
    reps = (sample_chains + proto - 1) // proto
    diag = np.tile(np.stack(diags), (reps, 1, 1, 1))[:sample_chains]
    sub = np.tile(np.stack(subs), (reps, 1, 1, 1))[:sample_chains]
    rhs = rng.standard_normal((sample_chains, T, D))
    cores = host_cores()
    info = []
    sec = _time_cpu(lambda: info.append(c_ref.chol_solve_batch(diag, sub, rhs, nthreads=cores)[3]), steps,
                    max(1, warmup))
    assert int(info[-1].max()) == 0
    return sample_chains * T / sec, sec, cores


def cpu_other_configs() -> dict:
    """Bounded CPU legs for configs 1, 3, 4, 5 (state-steps/s of the C port, cores used, what was run).
    A single series cannot be spread over cores by the reference's algorithm (one banded factorisation),
    so configs 1 and 3 run on ONE core; configs 4 and 5 thread over chains like config 2."""
    import numpy as np

    from oracle import c_ref, np_oracle as O

    cores = host_cores()
    rng = np.random.default_rng(71892305)
    out = {}
    lr = np.array([[0.1]])

    def matern32_series(t):
        k = O.Matern32(1.0, 1.0)
        tp = np.cumsum(rng.uniform(0.05, 0.15, size=t))
        ssm = k.state_space_model(tp[None])
        y = np.sin(tp)[None, :, None] + 0.1 * rng.standard_normal((1, t, 1))
        return ssm, k.emission_matrix(tp), y

    for tag, t, reps in (("config1_kalman_loglik_T1000", 1000, 200), ("config3_kalman_loglik", 1_000_000, 2)):
        ssm, h, y = matern32_series(t)
        sec = _time_cpu(lambda: c_ref.kalman_loglik_batch(ssm.mu0, ssm.chol_p0, ssm.a_s, ssm.b_s, ssm.chol_q_s,
                                                          h, y, lr, nthreads=1), reps)
        out[tag] = {"value": t / sec, "unit": UNIT, "cores": 1, "kind": "port", "ms": sec * 1e3,
                    "sample": f"ONE Matern32 series of T={t} (named: "
                              f"{'1e3' if t == 1000 else '1e7, per state-step'}), oracle/ssm_ref.c "
                              "ref_kalman_loglik_batch (SpInGP route of kalman_filter.py:184-255)"}
    # config 4: D = 17 sum kernel, one chain per core, T = 2000
    k4 = O.Sum([O.Matern52(1.0, 1.0)] + [O.HarmonicOscillator(0.5 ** j, 1.0 / j) for j in range(1, 8)], jitter=1e-6)
    t4, b4 = 2000, max(cores, 4)
    tp = np.cumsum(rng.uniform(0.05, 0.15, size=t4))
    pd, ps = O.kalman_k_inv_post(k4.state_space_model(tp), k4.emission_matrix(tp), np.array([[100.0]]))
    diag, sub = np.tile(pd, (b4, 1, 1, 1)), np.tile(ps, (b4, 1, 1, 1))
    rhs = rng.standard_normal((b4, t4, 17))
    sec = _time_cpu(lambda: c_ref.chol_solve_batch(diag, sub, rhs, nthreads=cores), 3)
    out["config4_cholesky_solve_d17"] = {
        "value": b4 * t4 / sec, "unit": UNIT, "cores": cores, "kind": "port", "ms": sec * 1e3,
        "sample": f"{b4} chains x T={t4} of the D=17 sum-kernel posterior precision (named: 256 x 1e5, per "
                  "state-step), oracle/banded_ref.c, OpenMP over chains"}
    # config 5: CVI naturals -> SSM -> expectations, D = 2
    t5, b5 = 10_000, 4 * cores
    k5 = O.Matern32(1.0, 1.0)
    tp = np.linspace(0.0, 0.1 * (t5 - 1), t5)
    ssm5 = k5.state_space_model(tp)
    h5 = k5.emission_matrix(tp)
    pd, ps = O.ssm_build_precision(ssm5)
    prec = rng.uniform(0.5, 2.0, size=(t5, 1, 1))
    th = (np.tile(np.einsum("tmd,tm->td", h5, rng.standard_normal((t5, 1))), (b5, 1, 1)),
          np.tile(-0.5 * (pd + np.einsum("tmd,tmn,tne->tde", h5, prec, h5)), (b5, 1, 1, 1)),
          np.tile(-ps, (b5, 1, 1, 1)))
    res = []
    sec = _time_cpu(lambda: res.append(c_ref.nat_to_ssm_batch(*th, nthreads=cores)), 3)
    a_s, offs, l0, lq, mu0 = res[-1]
    out["config5_naturals_to_ssm_params_f64"] = {
        "value": b5 * t5 / sec, "unit": UNIT, "cores": cores, "kind": "port", "ms": sec * 1e3,
        "sample": f"{b5} of 1024 chains x M={t5}, D=2, oracle/ssm_ref.c ref_nat_to_ssm_batch "
                  "(ssm_gaussian_transformations.py:332-511), OpenMP over chains"}
    sec = _time_cpu(lambda: c_ref.ssm_to_expectations_batch(mu0, l0, a_s, offs, lq, nthreads=cores), 3)
    out["config5_ssm_to_expectations_f64"] = {
        "value": b5 * t5 / sec, "unit": UNIT, "cores": cores, "kind": "port", "ms": sec * 1e3,
        "sample": f"{b5} of 1024 chains x M={t5}, D=2, ref_ssm_to_expectations_batch (:31-89)"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1024
    value, sec, cores = cpu_reference_run(args.steps, args.warmup, sample)
    by_config = {}
    try:
        by_config = cpu_other_configs()
    except Exception as exc:  # noqa: BLE001 -- the headline line must survive
        by_config = {"error": f"{type(exc).__name__}: {exc}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(int(os.environ.get("WORLD_SIZE", str(args.gpus)))),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of 4096 chains x T={T} per step (block->band, banded "
                                   "Cholesky, band->block, re-band, banded solve), OpenMP over chains",
                         "note": "TensorFlow reference not installable (py3.12, no TF/gpflow/banded_matrices): "
                                 "this arm times the C restatement of its CPU path",
                         "by_config": by_config},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def _timed(fn, warm=3, reps=10):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        z.record()
        evs.append((a, z))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(z) for a, z in evs)
    return ts[len(ts) // 2]


def _rel(a, ref) -> float:
    import numpy as np

    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))


def other_configs(dev, rank, world, dist, peak, args):
    """Configs 3, 1, 5, 4 on the GPU: device-resident timing (CUDA events, median of 10), fraction of the HBM
    roofline on SURVEY.md §8d's algorithmic bytes, and parity against the C port / numpy oracle on the SAME
    inputs at the named size (whole series for config 3, strided chains for configs 4 and 5)."""
    import numpy as np
    import torch

    import bench_inputs
    import markovflow_b200 as mf
    from markovflow_b200 import _lib
    from markovflow_b200.host import kalman_log_likelihood_host
    from markovflow_b200.parallel import CudaKalmanEngine, time_segment
    from oracle import c_ref, np_oracle as O

    out, kalman_e2e, oracle_ll3 = {}, None, None
    npy = lambda x: x.detach().cpu().numpy()

    def entry(steps, bytes_per_step, ms, **kw):
        gbs = steps * bytes_per_step / (ms * 1e-3) / 1e9
        # a series split over the ranks ("strong") streams from all their memories: the roofline is N peaks
        npeak = world if kw.get("scaling", "").startswith("strong") else 1
        return {"value": steps / (ms * 1e-3), "unit": UNIT, "ms": ms, "bytes_per_state_step": bytes_per_step,
                "achieved_GBps": gbs, "frac": gbs / (peak * npeak), **kw}

    # ---- config 3: one Matern32 series, T = 1e7, float64 Kalman log-likelihood ------------------------
    t3 = 10_000_000
    ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t3, dev)
    mu0, l0, a, b, lq, _, _, d = ssm._flat()
    ring = None
    if world == 1:
        ms = _timed(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
        ll = float(mf.kalman_log_likelihood(ssm, h, y, lr))
        e3 = entry(t3, 104, ms, workload="Matern32 D=2, ONE series T=1e7, f64 (parallel in time, one pass)",
                   loglik=ll, scaling="single GPU")
    else:
        from markovflow_b200.parallel import PeerRing, time_sharded_log_likelihood

        seg = time_segment(mu0, l0, a, b, lq, h.reshape(1, t3, 1, d), y.reshape(1, t3, 1),
                           lr.reshape(1, 1, 1), rank, world)
        eng = CudaKalmanEngine()
        ring = PeerRing(torch.float64, 1, d)  # peer-mapped regions (cudaIpc over NVLink), set up once

        def via_nccl():
            elem = eng.segment_summary(seg)
            gathered = [torch.empty_like(elem) for _ in range(world)]
            dist.all_gather(gathered, elem)
            return eng.fold(torch.stack(gathered), d)[:, -1]

        def via_peers():
            return time_sharded_log_likelihood(seg, engine=eng, ring=ring)

        def timed_max(fn):
            dist.barrier()
            ms_ = _timed(fn)
            t_ = torch.tensor([ms_], dtype=torch.float64, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item())

        ms_nccl = timed_max(via_nccl)
        ms = timed_max(via_peers)
        ll = float(via_peers()[0])
        e3 = entry(t3, 104, ms, workload=f"Matern32 D=2, ONE series T=1e7, f64, time-sharded over {world} GPUs: "
                   "segment element + exchange over peer memory + ordered join in ONE reduction kernel "
                   "(mf_kalman_time_sharded_log_likelihood)", loglik=ll, scaling="strong",
                   ms_nccl_all_gather_plus_fold=ms_nccl, rel_diff_peers_vs_nccl=abs(ll - float(via_nccl()[0])) / abs(ll),
                   frac_note="of n_gpus HBM peaks")
    if rank == 0:
        # parity at the named size: the C port of the reference's SpInGP route over the WHOLE series, and
        # end to end from host memory (mf_host_kalman_log_likelihood: 104 B per step in, one value out)
        try:
            host = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x) for x in
                    (mu0, l0, a, b, lq, h.reshape(t3, 1, d), y.reshape(1, t3, 1), lr)]
            torch.cuda.synchronize()
            if not args.no_e2e:
                ll_h = kalman_log_likelihood_host(*host, device=dev)
                t0 = time.perf_counter()
                n_e2e = 3
                for _ in range(n_e2e):
                    ll_h = kalman_log_likelihood_host(*host, device=dev)
                sec = (time.perf_counter() - t0) / n_e2e
                h2d_b, d2h_b = kalman_log_likelihood_host.last_bytes
                kalman_e2e = {"value": t3 / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "h2d_bytes_per_step": h2d_b,
                              "d2h_bytes_per_step": d2h_b, "rel_diff_vs_device_resident": abs(float(ll_h[0]) - ll) / abs(ll),
                              "api": "mf_host_kalman_log_likelihood (pinned host arrays, time chunks -> scan "
                                     "elements on the device, one value back)"}
            t0 = time.perf_counter()
            want = c_ref.kalman_loglik_batch(*(x.numpy() for x in host), nthreads=1)
            oracle_ll3 = float(want[0])
            e3["parity_max_rel_err_vs_oracle"] = abs(ll - oracle_ll3) / abs(oracle_ll3)
            e3["parity_note"] = f"C port of kalman_filter.py:184-255 over all 1e7 steps ({time.perf_counter() - t0:.1f} s, 1 core)"
            del host, want
        except Exception as exc:  # noqa: BLE001
            e3["parity_error"] = f"{type(exc).__name__}: {exc}"
    out["config3_kalman_loglik"] = e3

    # ---- config 3 again with the SSM built INSIDE the kernel from the time deltas ------
    from markovflow_b200.parallel import matern_time_segment, time_sharded_matern_log_likelihood
    dts = bench_inputs.matern32_time_deltas(1, t3, dev)
    y2 = y.reshape(1, t3).contiguous()
    one = torch.ones(1, dtype=torch.float64, device=dev)
    if world == 1:
        fused = lambda: mf.matern_kalman_log_likelihood(2, one, one, y2, lr, time_deltas=dts)
        ms = _timed(fused)
        kw = dict(scaling="single GPU", ms_cuda_graph_replay=_timed(mf.Graphed(fused), warm=2, reps=10))
    else:
        first, seg_dt, seg_y = matern_time_segment(dts, y2, rank, world)
        fused = lambda: ring.matern(2, one, one, seg_dt, seg_y, lr, first)[0]
        dist.barrier()
        ms = _timed(fused)
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        kw = dict(scaling="strong")
    ll2 = float(fused()[0])
    e = entry(t3, 16, ms, workload="same series, A_k/Q_k built in registers from dt_k "
              "(mf_kalman_matern_log_likelihood): 16 B per step cross HBM, arithmetic-bound",
              loglik=ll2, rel_diff_vs_materialised=abs(ll2 - ll) / abs(ll), **kw)
    e["equivalent_GBps_of_materialised_ssm"] = t3 * 104 / (ms * 1e-3) / 1e9
    if oracle_ll3 is not None:
        e["parity_max_rel_err_vs_oracle"] = abs(ll2 - oracle_ll3) / abs(oracle_ll3)
        e["parity_note"] = "same series: the C port's value for the materialised SSM"
    out["config3_kalman_loglik_from_time_deltas"] = e
    del dts, y2, ssm, h, y, mu0, l0, a, b, lq
    torch.cuda.empty_cache()
    # ---- the same job ten times longer (T = 1e8): per-GPU work that outweighs the exchange latency --------
    try:
        t8 = 100_000_000
        ssm8, h8, y8, lr8 = bench_inputs.kalman_inputs_config3(t8, dev)
        if world == 1:
            ms8 = _timed(lambda: mf.kalman_log_likelihood(ssm8, h8, y8, lr8), warm=2, reps=5)
            ll8 = float(mf.kalman_log_likelihood(ssm8, h8, y8, lr8))
            out["config3_kalman_loglik_T1e8"] = entry(t8, 104, ms8, loglik=ll8, scaling="single GPU",
                                                      workload="Matern32 D=2, ONE series T=1e8, f64")
        else:
            f8 = ssm8._flat()
            seg8 = time_segment(*f8[:5], h8.reshape(1, t8, 1, 2), y8.reshape(1, t8, 1), lr8.reshape(1, 1, 1), rank, world)
            del ssm8, h8, y8, f8
            torch.cuda.empty_cache()
            eng8 = CudaKalmanEngine()
            dist.barrier()
            ms8 = _timed(lambda: time_sharded_log_likelihood(seg8, engine=eng8, ring=ring), warm=2, reps=5)
            t_ = torch.tensor([ms8], dtype=torch.float64, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ll8 = float(time_sharded_log_likelihood(seg8, engine=eng8, ring=ring)[0])
            out["config3_kalman_loglik_T1e8"] = entry(
                t8, 104, float(t_.item()), loglik=ll8, scaling="strong",
                workload=f"Matern32 D=2, ONE series T=1e8, f64, time-sharded over {world} GPUs (peer-memory exchange)")
            del seg8
        torch.cuda.empty_cache()
    except Exception as exc:  # noqa: BLE001
        out["config3_kalman_loglik_T1e8"] = {"skipped": f"{type(exc).__name__}: {exc}"}
    if ring is not None:
        dist.barrier()
        ring.close()
    if rank != 0:
        return out, kalman_e2e

    # ---- config 1: ONE Matern32 series T = 1000: log-likelihood; + posterior SSM + marginals -----------
    ssm1, h1, y1, lr1 = bench_inputs.kalman_inputs_config3(1000, dev)
    kf = mf.KalmanFilter(ssm1, mf.EmissionModel(h1), y1, lr1)

    def gpr_job():
        post = kf.posterior_state_space_model()
        return kf.log_likelihood(), post.marginals

    ms_ll = _timed(lambda: kf.log_likelihood(), warm=3, reps=20)
    ms_job = _timed(gpr_job, warm=2, reps=5)
    f1 = ssm1._flat()
    want = c_ref.kalman_loglik_batch(*(npy(x) for x in f1[:5]), npy(h1.reshape(1000, 1, 2)), npy(y1.reshape(1, 1000, 1)),
                                     npy(lr1), nthreads=1)
    ref_post = O.kalman_posterior_ssm(O.SSM(*(npy(x) for x in f1[:5])), npy(h1.reshape(1, 1000, 1, 2)),
                                      npy(y1.reshape(1, 1000, 1)), O._r_inv_from_chol(npy(lr1)))
    ll1, (m1, c1) = gpr_job()
    out["config1_kalman_loglik_T1000"] = {
        "value": 1000 / (ms_ll * 1e-3), "unit": UNIT, "ms": ms_ll, "ms_loglik_plus_posterior_marginals": ms_job,
        "ms_job_cuda_graph_replay": _timed(mf.Graphed(gpr_job), warm=2, reps=10),
        "parity_max_rel_err_vs_oracle": max(abs(float(ll1) - float(want[0])) / abs(float(want[0])),
                                            _rel(npy(m1), O.ssm_marginal_means(ref_post)),
                                            _rel(npy(c1), O.ssm_marginal_covariances(ref_post))),
        "workload": "Matern32 D=2, ONE series T=1000, f64: KalmanFilter.log_likelihood (value / ms); "
                    "+ posterior_state_space_model + posterior marginals (ms_loglik_plus_...); launch-bound"}
    del ssm1, h1, y1, kf
    torch.cuda.empty_cache()

    # ---- config 5: CVI site update, B = 1024 chains x M = 1e4 inducing states, D = 2 ---------------
    b5, t5 = 1024, 10_000
    th64 = bench_inputs.cvi_naturals_config5(b5, t5, dev, dtype=torch.float64)
    pick = torch.arange(0, b5, b5 // 8, device=dev)
    o_nat = c_ref.nat_to_ssm_batch(*(npy(x[pick]) for x in th64))
    o_exp = c_ref.ssm_to_expectations_batch(o_nat[4], o_nat[2], o_nat[0], o_nat[1], o_nat[3])
    for dtype, es, tag in ((torch.float64, 8, "f64"), (torch.float32, 4, "f32")):
        th = tuple(x.to(dtype) for x in th64)
        ms = _timed(lambda: mf.naturals_to_ssm_params(*th))
        got = mf.naturals_to_ssm_params(*th)
        err = max(_rel(npy(g[pick]), w) for g, w in zip(got, o_nat))
        # the API raises on a non-positive pivot, which costs one reduction + a device->host sync per call; with
        # the check off (markovflow_b200.set_check_numerics(False)) the calls queue back to back: kernel time only
        mf.set_check_numerics(False)
        try:
            ms_nocheck = _timed(lambda: mf.naturals_to_ssm_params(*th))
        finally:
            mf.set_check_numerics(True)
        out[f"config5_naturals_to_ssm_params_{tag}"] = entry(
            b5 * t5, 20 * es, ms, workload="Matern32 prior + sites, B=1024 x M=1e4, D=2",
            parity_max_rel_err_vs_oracle=err, parity_note="8 strided chains vs the float64 C port",
            ms_without_failure_check=ms_nocheck,
            frac_without_failure_check=b5 * t5 * 20 * es / (ms_nocheck * 1e-3) / 1e9 / peak)
        q = mf.StateSpaceModel(*(g.contiguous() for g in (got[4], got[2], got[0], got[1], got[3])))
        ms = _timed(lambda: mf.ssm_to_expectations(q))
        err = max(_rel(npy(g[pick]), w) for g, w in zip(mf.ssm_to_expectations(q), o_exp))
        out[f"config5_ssm_to_expectations_{tag}"] = entry(
            b5 * t5, 20 * es, ms, parity_max_rel_err_vs_oracle=err, parity_note="8 strided chains vs the float64 C port")
    del th64, th, got, q
    torch.cuda.empty_cache()

    # ---- config 4 at its named size (B=256 x T=1e5: 118 GB of blocks, factored in place) ------------
    try:
        from tools.config4_full import run as config4_full

        free_gb = torch.cuda.mem_get_info(dev)[0] / 1e9
        if free_gb < 150:
            raise RuntimeError(f"only {free_gb:.0f} GB of device memory free (needs ~145 GB)")
        e = config4_full(256, 100_000, dev)
        e.update(value=e["state_steps_per_s"], unit=UNIT, frac=e["achieved_GBps"] / peak)
        out["config4_cholesky_solve_d17"] = e
    except Exception as exc:  # noqa: BLE001 -- e.g. a smaller-memory device
        out["config4_cholesky_solve_d17"] = {"skipped": f"{type(exc).__name__}: {exc}"}
    torch.cuda.empty_cache()
    return out, kalman_e2e


def run_gpu(args):
    import torch

    import bench_inputs
    from markovflow_b200 import _lib
    from markovflow_b200.host import cholesky_solve_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    b = B_PER_GPU
    diag, sub, rhs = bench_inputs.matern52_posterior_precision(b, T, dev, seed=bench_inputs.SEED + rank)
    od, os_, ox = torch.empty_like(diag), torch.empty_like(sub), torch.empty_like(rhs)
    info = torch.empty(b, dtype=torch.int32, device=dev)
    lib = _lib.lib()

    def step():
        st = lib.mf_btd_cholesky(_lib.MF_F64, _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(rhs), _lib.ptr(od),
                                 _lib.ptr(os_), _lib.ptr(ox), None, _lib.ptr(info), _lib.i64(b),
                                 _lib.i64(T), _lib.i64(D), _lib.current_stream())
        _lib.check(st, "mf_btd_cholesky")

    # ---- device-resident timing ----------------------------------------------------------------
    # cudaProfilerStart/Stop bracket the warm-up + timed launches: `ncu --profile-from-start off`
    # then lists exactly the launches of the timed region (no effect without a profiler)
    torch.cuda.profiler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for a, z in evs:
            a.record()
            step()
            z.record()
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
        # keep the sampled window long enough (~1.5 s) for several nvidia-smi samples under load
        for _ in range(max(0, 450 - args.steps)):
            step()
        torch.cuda.synchronize()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    kernel_ms = statistics.mean(a.elapsed_time(z) for a, z in evs)
    assert int(info.max()) == 0, "synthetic precision was not positive definite"
    ms_per_step = total_ms / args.steps
    value = world * b * T / (ms_per_step * 1e-3)

    # ---- parity spot check against the oracle on a strided subset of chains (not timed) --------
    parity = None
    if rank == 0:
        import numpy as np

        from oracle import np_oracle as O

        idx = torch.arange(0, b, b // 8, device=dev)
        dg, sb, rh = diag[idx].cpu().numpy(), sub[idx].cpu().numpy(), rhs[idx].cpu().numpy()
        o_ld, o_ls = O.btd_cholesky(dg, sb)
        o_x = O.btd_solve(o_ld, o_ls, rh)
        parity = max(_rel(od[idx].cpu().numpy(), o_ld), _rel(os_[idx].cpu().numpy(), o_ls),
                     _rel(ox[idx].cpu().numpy(), o_x))
        assert parity < 1e-10, f"parity vs oracle failed: {parity:.3e}"

    # ---- end-to-end through the host-buffer C ABI (pinned host memory, copies timed) -------------
    e2e = None
    if not args.no_e2e:
        hd = torch.empty(diag.shape, dtype=diag.dtype, pin_memory=True)
        hs = torch.empty(sub.shape, dtype=sub.dtype, pin_memory=True)
        hr = torch.empty(rhs.shape, dtype=rhs.dtype, pin_memory=True)
        hd.copy_(diag); hs.copy_(sub); hr.copy_(rhs)
        out = (torch.empty(diag.shape, dtype=diag.dtype, pin_memory=True),
               torch.empty(sub.shape, dtype=sub.dtype, pin_memory=True),
               torch.empty(rhs.shape, dtype=rhs.dtype, pin_memory=True),
               torch.empty(b, dtype=torch.int32, pin_memory=True))
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(2):
            cholesky_solve_host(hd, hs, hr, out=out, device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            cholesky_solve_host(hd, hs, hr, out=out, device=dev)
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
        h2d_b, d2h_b = cholesky_solve_host.last_bytes
        e2e = {"value": world * b * T / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_b,
               "d2h_bytes_per_step": d2h_b, "ms_per_step": e2e_s * 1e3, "steps": n_e2e,
               "api": "mf_host_btd_cholesky (C ABI; pinned host arrays, 128-chain chunks, 3 streams)"}
        if rank == 0:
            # chunks of <= 1024 chains take the parallel-in-time path: same factor to rounding
            ref_ld = od[::512].cpu()
            e2e_err = float((out[0][::512] - ref_ld).abs().max() / ref_ld.abs().max())
            assert e2e_err < 1e-11, f"e2e result differs from device-resident result: {e2e_err:.2e}"
        del hd, hs, hr, out
        lib.mf_host_release(-1)

    peak, peak_src = measured_peak()
    others, kalman_e2e = None, None
    if not args.no_extras:
        del diag, sub, rhs, od, os_, ox
        torch.cuda.empty_cache()
        try:
            others, kalman_e2e = other_configs(dev, rank, world, dist, peak, args)
        except Exception as exc:  # noqa: BLE001 -- the headline line must survive a failing extra
            import traceback

            traceback.print_exc(file=sys.stderr)
            others = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    achieved = ALGO_BYTES_PER_STEP * b * T / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("btd_chol_tma_kernel_dram_bytes_per_launch")
        traffic_src = "static: " + tj.get("source", "profiles/traffic.json (ncu --set full capture of this kernel)")
    cpu, cpu_by = None, {}
    if not args.no_cpu:
        v, sec, cores = cpu_reference_run(steps=3, warmup=1, sample_chains=1024)
        try:
            cpu_by = cpu_other_configs()
        except Exception as exc:  # noqa: BLE001
            cpu_by = {"error": f"{type(exc).__name__}: {exc}"}
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1024 of 4096 chains x T={T}, 3 timed passes of oracle/banded_ref.c "
                         f"({sec:.2f} s per pass), OpenMP over chains", "by_config": cpu_by}
    # compact per-config lines inside the objects the driver keeps; details (and config 3 LAST, so that it
    # survives a truncated tail) in `extra`
    compact_keys = ("value", "ms", "frac", "bytes_per_state_step", "parity_max_rel_err_vs_oracle", "scaling",
                    "ms_without_failure_check", "frac_without_failure_check")
    by_config = {}
    if others and "error" not in others:
        by_config = {k: {q: v[q] for q in compact_keys if q in v} for k, v in others.items()}
        for k, v in by_config.items():
            if k in cpu_by and "value" in cpu_by[k] and "value" in v:
                v["cpu_value"], v["cpu_cores"] = cpu_by[k]["value"], cpu_by[k]["cores"]
    extra = None
    if others:
        order = [k for k in others if not k.startswith("config3")] + [k for k in others if k.startswith("config3")][::-1]
        extra = {k: others[k] for k in order}
    if e2e is not None and kalman_e2e is not None:
        e2e["kalman"] = kalman_e2e
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
        "e2e": e2e, "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel": "btd_chol_tma_kernel<double,3,rhs>",
                     "kernel_ms": kernel_ms, "algorithmic_bytes_per_state_step": ALGO_BYTES_PER_STEP,
                     "parity_max_rel_err_vs_oracle": parity, "by_config": by_config},
        "cpu_baseline": cpu, "clocks": clocks.summary(), "extra": extra,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 1/3/4/5 measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: everything else that libraries write to file descriptor 1
    # (e.g. NCCL's version banner) is sent to stderr; the JSON line goes to the saved descriptor.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
