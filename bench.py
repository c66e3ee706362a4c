#!/usr/bin/env python
"""Headline benchmark: block-tridiagonal Cholesky + solve, BASELINE.json config 2.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W  (CPU reference arm)

Workload (``config.workload``): Matern52 (D=3) posterior precision, B=4096 independent series per
GPU x T=10,000 states, float64; one *step* = one fused Cholesky + forward-solve sweep over the whole
batch.  Chains are independent, so N GPUs hold N x 4096 series (weak scaling, no collective on the
data path).  ``value`` = state-steps/s with inputs resident in HBM; ``e2e`` = the same work through
``markovflow_b200.host.cholesky_solve_host`` with pinned HOST buffers (copies inside the timed
region); ``roofline`` = algorithmic bytes / kernel time against the measured HBM copy bandwidth;
``cpu_baseline`` = the C restatement of the reference's CPU path (oracle/banded_ref.c) on the host
cores of the same box.  Prints ONE JSON line on rank 0.
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


def cpu_reference_run(steps: int, warmup: int, sample_chains: int):
    """Time the C port of the reference path on a bounded sample of the workload (host cores)."""
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
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _ in range(max(1, warmup)):
        c_ref.chol_solve_batch(diag, sub, rhs, nthreads=cores)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        _, _, _, info = c_ref.chol_solve_batch(diag, sub, rhs, nthreads=cores)
        times.append(time.perf_counter() - t0)
    assert int(info.max()) == 0
    sec = sum(times) / len(times)
    return sample_chains * T / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1024
    value, sec, cores = cpu_reference_run(args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "TensorFlow reference not installable (py3.12, no TF/"
                   "gpflow/banded_matrices); this arm times the C restatement of its CPU path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of 4096 chains x T={T} per step (block->band, banded "
                                   "Cholesky, band->block, re-band, banded solve), OpenMP over chains"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


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


def extra_measurements(dev, rank, world, dist, peak):
    """The other BASELINE configurations (not the headline line): Kalman log-likelihood of ONE long
    series (config 3; time-sharded over the ranks when world > 1), the CVI parameter transforms
    (config 5, f64 and f32) and the large-block factorisation (config 4 at reduced T).  Device-
    resident inputs, CUDA-event timing, median of 10; roofline on SURVEY.md §8d's algorithmic bytes."""
    import torch

    import bench_inputs
    import markovflow_b200 as mf
    from markovflow_b200 import _lib
    from markovflow_b200.parallel import CudaKalmanEngine, time_segment

    out = {}

    def entry(steps, bytes_per_step, ms, **kw):
        gbs = steps * bytes_per_step / (ms * 1e-3) / 1e9
        return {"state_steps_per_s": steps / (ms * 1e-3), "ms": ms, "bytes_per_state_step": bytes_per_step,
                "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak, **kw}

    # ---- config 3: one Matern32 series, T = 1e7, float64 Kalman log-likelihood -------------------
    t3 = 10_000_000
    ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t3, dev)
    if world == 1:
        ms = _timed(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
        ll = float(mf.kalman_log_likelihood(ssm, h, y, lr))
        out["config3_kalman_loglik"] = entry(
            t3, 104, ms, workload="Matern32 D=2, single series T=1e7, f64, parallel-in-time "
            "(one pass: per-segment scan elements + ordered reduction)", loglik=ll, scaling="single GPU")
    else:
        mu0, l0, a, b, lq, bsz, t, d = ssm._flat()
        seg = time_segment(mu0, l0, a, b, lq, h.reshape(1, t, 1, d), y.reshape(1, t, 1),
                           lr.reshape(1, 1, 1), rank, world)
        eng = CudaKalmanEngine()

        def sharded():
            elem = eng.segment_summary(seg)
            gathered = [torch.empty_like(elem) for _ in range(world)]
            dist.all_gather(gathered, elem)
            return eng.fold(torch.stack(gathered), d)[:, -1]

        ms = _timed(sharded)
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ll = float(sharded()[0])
        ref = float(mf.kalman_log_likelihood(ssm, h, y, lr)) if rank == 0 else None
        out["config3_kalman_loglik"] = entry(
            t3, 104, float(tms.item()), workload=f"Matern32 D=2, single series T=1e7, f64, time-sharded "
            f"over {world} GPUs: local segment element, NCCL all-gather of {world} x 136 B, ordered fold",
            loglik=ll, loglik_single_gpu=ref, scaling="strong",
            frac_note="fraction of the AGGREGATE peak = frac_of_hbm_peak / n_gpus")
    # ---- config 3 again with the SSM built INSIDE the kernel from the time deltas ------
    # same series (same deltas, same observations): a step reads (dt_k, y_k) = 16 B instead of 104 B
    from markovflow_b200.parallel import matern_time_segment, time_sharded_matern_log_likelihood
    dts = bench_inputs.matern32_time_deltas(1, t3, dev)
    y2 = y.reshape(1, t3).contiguous()
    one = torch.ones(1, dtype=torch.float64, device=dev)
    ll_mat = out["config3_kalman_loglik"].get("loglik_single_gpu") or out["config3_kalman_loglik"]["loglik"]
    if world == 1:
        fused = lambda: mf.matern_kalman_log_likelihood(2, one, one, y2, lr, time_deltas=dts)
        ms = _timed(fused)
        ll2 = float(fused()[0])
        extra_kw = dict(scaling="single GPU", ms_cuda_graph_replay=_timed(mf.Graphed(fused), warm=2, reps=10))
    else:
        first, seg_dt, seg_y = matern_time_segment(dts, y2, rank, world)
        fused = lambda: time_sharded_matern_log_likelihood(2, one, one, seg_dt, seg_y, lr, first)
        ms = _timed(fused)
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        ll2 = float(fused()[0])
        extra_kw = dict(scaling="strong", frac_note="fraction of the AGGREGATE peak = frac_of_hbm_peak / n_gpus")
    e = entry(t3, 16, ms, workload="same series as config3_kalman_loglik, Matern32 A_k/Q_k built in "
              "registers from dt_k (mf_kalman_matern_log_likelihood): arithmetic-bound, 16 B per step "
              "cross HBM; speed-up over the materialised-SSM kernel = ratio of the two ms",
              loglik=ll2, rel_diff_vs_materialised=abs(ll2 - ll_mat) / abs(ll_mat), **extra_kw)
    e["equivalent_GBps_of_materialised_ssm"] = t3 * 104 / (ms * 1e-3) / 1e9
    out["config3_kalman_loglik_from_time_deltas"] = e
    del dts, y2
    del ssm, h, y
    torch.cuda.empty_cache()
    if rank != 0:
        return out

    # ---- config 1: ONE Matern32 series, log-likelihood + posterior SSM + posterior marginals --------
    for t1 in (1_000, 1_000_000):
        ssm1, h1, y1, lr1 = bench_inputs.kalman_inputs_config3(t1, dev)
        kf = mf.KalmanFilter(ssm1, mf.EmissionModel(h1), y1, lr1)

        def gpr_job():
            post = kf.posterior_state_space_model()
            return kf.log_likelihood(), post.marginals

        ms = _timed(gpr_job, warm=2, reps=5)
        ms_graph = _timed(mf.Graphed(gpr_job), warm=2, reps=10) if t1 <= 100_000 else None
        out[f"config1_gpr_single_series_T{t1}"] = {
            "ms": ms, "ms_cuda_graph_replay": ms_graph, "state_steps_per_s": t1 / (ms * 1e-3),
            "workload": f"Matern32 D=2, ONE series T={t1}, f64: KalmanFilter.log_likelihood + "
                        "posterior_state_space_model + posterior marginals; every sweep parallel in time"}
        del ssm1, h1, y1, kf
    torch.cuda.empty_cache()

    # ---- config 5: CVI site update, B = 1024 chains x M = 1e4 inducing states, D = 2 ---------------
    b5, t5 = 1024, 10_000
    th64 = bench_inputs.cvi_naturals_config5(b5, t5, dev, dtype=torch.float64)
    ref = mf.naturals_to_ssm_params(*th64)
    for dtype, es, tag in ((torch.float64, 8, "f64"), (torch.float32, 4, "f32")):
        th = tuple(x.to(dtype) for x in th64)
        ms = _timed(lambda: mf.naturals_to_ssm_params(*th))
        got = mf.naturals_to_ssm_params(*th)
        err = max(float((g.double() - r).abs().max() / r.abs().max()) for g, r in zip(got, ref))
        out[f"config5_naturals_to_ssm_params_{tag}"] = entry(
            b5 * t5, 20 * es, ms, workload="Matern32 prior + sites, B=1024 x M=1e4, D=2",
            max_rel_err_vs_f64=err)
        # dense parameter arrays (the transform returns slices of its concatenated outputs)
        q = mf.StateSpaceModel(*(g.contiguous() for g in (got[4], got[2], got[0], got[1], got[3])))
        ms = _timed(lambda: mf.ssm_to_expectations(q))
        out[f"config5_ssm_to_expectations_{tag}"] = entry(b5 * t5, 20 * es, ms)
    del th64, ref, th, got, q
    torch.cuda.empty_cache()

    # ---- config 4 at reduced T: D = 17 sum kernel, B = 256, in-place Cholesky + solve -------------
    b4, t4 = 256, 4000
    diag, sub, rhs = bench_inputs.sum_kernel_posterior_precision(b4, t4, dev)
    d0, s0 = diag.clone(), sub.clone()
    x = torch.empty_like(rhs)
    info = torch.empty(b4, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    ts = []
    for _ in range(4):
        diag.copy_(d0)
        sub.copy_(s0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mf_btd_cholesky(_lib.MF_F64, _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(rhs),
                                       _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(x), None, _lib.ptr(info),
                                       _lib.i64(b4), _lib.i64(t4), _lib.i64(17), _lib.current_stream()),
                   "mf_btd_cholesky")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    assert int(info.abs().max()) == 0
    out["config4_cholesky_solve_d17"] = entry(
        b4 * t4, 9520, sorted(ts[1:])[1], workload=f"Matern52 + 7 harmonics (D=17), B=256 x T={t4} "
        "(median of 3 launches; the named size follows), f64, in place, one warp per chain")
    del diag, sub, rhs, d0, s0, x
    torch.cuda.empty_cache()

    # ---- config 4 at its named size (B=256 x T=1e5: 118 GB of blocks, factored in place) ------------
    try:
        from tools.config4_full import run as config4_full

        free_gb = torch.cuda.mem_get_info(dev)[0] / 1e9
        if free_gb < 150:
            raise RuntimeError(f"only {free_gb:.0f} GB of device memory free (needs ~145 GB)")
        e = config4_full(256, 100_000, dev)
        e["frac_of_hbm_peak"] = e["achieved_GBps"] / peak
        out["config4_cholesky_solve_d17_named_size"] = e
    except Exception as exc:  # noqa: BLE001 -- e.g. a smaller-memory device
        out["config4_cholesky_solve_d17_named_size"] = {"skipped": f"{type(exc).__name__}: {exc}"}
    torch.cuda.empty_cache()
    return out


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

        def rel(a, ref):
            return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))

        parity = max(rel(od[idx].cpu().numpy(), o_ld), rel(os_[idx].cpu().numpy(), o_ls),
                     rel(ox[idx].cpu().numpy(), o_x))
        assert parity < 1e-10, f"parity vs oracle failed: {parity:.3e}"

    # ---- end-to-end through the host-buffer API (pinned host memory, copies timed) -------------
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
               "api": "markovflow_b200.host.cholesky_solve_host (pinned host tensors, 128-chain chunks, 3 streams)"}
        if rank == 0:
            # chunks of <= 1024 chains take the parallel-in-time path: same factor to rounding
            ref_ld = od[::512].cpu()
            e2e_err = float((out[0][::512] - ref_ld).abs().max() / ref_ld.abs().max())
            assert e2e_err < 1e-11, f"e2e result differs from device-resident result: {e2e_err:.2e}"

    peak, peak_src = measured_peak()
    extras = None
    if not args.no_extras:
        del diag, sub, rhs, od, os_, ox
        torch.cuda.empty_cache()
        try:
            extras = extra_measurements(dev, rank, world, dist, peak)
        except Exception as exc:  # noqa: BLE001 -- the headline line must survive a failing extra
            import traceback

            traceback.print_exc(file=sys.stderr)
            extras = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    achieved = ALGO_BYTES_PER_STEP * b * T / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("btd_chol_tma_kernel_dram_bytes_per_launch")
    cpu = None
    if not args.no_cpu:
        v, sec, cores = cpu_reference_run(steps=3, warmup=1, sample_chains=1024)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1024 of 4096 chains x T={T}, 3 timed passes of oracle/banded_ref.c "
                         f"({sec:.2f} s per pass), OpenMP over chains"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": b, "T": T, "D": D, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "inputs (6.9 GB) and outputs (6.9 GB) per step exceed the 126 MB L2",
                   "parity_max_rel_err_vs_oracle": parity},
        "e2e": e2e, "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": "btd_chol_tma_kernel<double,3,rhs>", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_state_step": ALGO_BYTES_PER_STEP},
        "cpu_baseline": cpu, "clocks": clocks.summary(), "extra": extras,
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
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3/4/5 measurements")
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
