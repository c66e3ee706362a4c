"""Times the non-headline kernels on their BASELINE configurations (device-resident, CUDA events):
    python tools/bench_paths.py [kalman3] [cvi5] [kalman_batch] [ssm]
Prints one JSON line per measurement with the HBM-roofline fraction (algorithmic bytes per
state-step from SURVEY.md §8d)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench_inputs  # noqa: E402
import markovflow_b200 as mf  # noqa: E402
from markovflow_b200 import _lib  # noqa: E402

DEV = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def report(name, steps, bytes_per_step, ms, **extra):
    gbs = steps * bytes_per_step / (ms * 1e-3) / 1e9
    print(json.dumps({"name": name, "ms": round(ms, 4), "state_steps_per_s": steps / (ms * 1e-3),
                      "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 4),
                      "bytes_per_step": bytes_per_step, **extra}), flush=True)


def kalman3(t=10_000_000):
    ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t, DEV)
    lib = _lib.lib()
    for mode, label in ((2, "parallel-in-time"),):
        lib.mf_set_tuning(2, mode)
        ms = timeit(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
        ll = float(mf.kalman_log_likelihood(ssm, h, y, lr))
        report(f"kalman_loglik config3 T={t} D=2 f64 [{label}]", t, (2 * 4 + 2 + 2 + 1) * 8, ms, loglik=ll)
    lib.mf_set_tuning(2, 0)


def kalman_batch(b=4096, t=10_000):
    mu0, l0, a, off, lq, h = bench_inputs.matern32_ssm(b, t, DEV, jitter_hyper=True)
    ssm = mf.StateSpaceModel(mu0, l0, a, off, lq)
    y = torch.randn(b, t, 1, dtype=torch.float64, device=DEV)
    lr = torch.tensor([[0.1]], dtype=torch.float64, device=DEV)
    ms = timeit(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
    report(f"kalman_loglik B={b} T={t} D=2 f64 [thread per chain]", b * t, (2 * 4 + 2 + 1) * 8, ms,
           note="H shared over the batch")


def cvi5(b=1024, t=10_000):
    for dtype, s in ((torch.float64, 8), (torch.float32, 4)):
        th = bench_inputs.cvi_naturals_config5(b, t, DEV, dtype=dtype)
        ms = timeit(lambda: mf.naturals_to_ssm_params(*th))
        report(f"naturals_to_ssm_params config5 B={b} T={t} D=2 {dtype}", b * t, (4 * 4 + 2 * 2) * s, ms)
        p = mf.naturals_to_ssm_params(*th)
        q = mf.StateSpaceModel(p[4], p[2], p[0], p[1], p[3])
        ms = timeit(lambda: mf.ssm_to_expectations(q))
        report(f"ssm_to_expectations config5 B={b} T={t} D=2 {dtype}", b * t, (4 * 4 + 2 * 2) * s, ms)
        ms = timeit(lambda: mf.ssm_to_naturals(q))
        report(f"ssm_to_naturals config5 B={b} T={t} D=2 {dtype}", b * t, (4 * 4 + 2 * 2) * s, ms)


def ssm(b=4096, t=10_000):
    mu0, l0, a, off, lq, h = bench_inputs.matern32_ssm(b, t, DEV, jitter_hyper=True)
    m = mf.StateSpaceModel(mu0, l0, a, off, lq)
    report("ssm.marginals B=4096 T=1e4 D=2", b * t, (3 * 4 + 2 * 2) * 8, timeit(lambda: m.marginals))
    report("ssm.precision B=4096 T=1e4 D=2", b * t, (4 * 4) * 8, timeit(lambda: m.precision))
    x = m.sample(())
    report("ssm.sample B=4096 T=1e4 D=2 (incl. randn)", b * t, (2 * 4 + 3 * 2) * 8, timeit(lambda: m.sample(())))
    report("ssm.log_pdf B=4096 T=1e4 D=2", b * t, (2 * 4 + 2 * 2) * 8, timeit(lambda: m.log_pdf(x)))
    report("ssm.kl_divergence B=4096 T=1e4 D=2", b * t, (4 * 4 + 2 * 2) * 8, timeit(lambda: m.kl_divergence(m)))


def btd(b=4096, t=10_000):
    """solve / inverse subset / U D U^T on config-2 shapes (D = 3), sweep vs direct-load kernels."""
    diag, sub, rhs = bench_inputs.matern52_posterior_precision(b, t, DEV)
    m = mf.SymmetricBlockTriDiagonal(diag, sub)
    chol = m.cholesky
    lib = _lib.lib()
    for knob, label in ((0, "TMA sweep"), (1, "direct loads")):
        lib.mf_set_tuning(4, knob)
        report(f"btd.solve fwd B={b} T={t} D=3 [{label}]", b * t, (2 * 9 + 2 * 3) * 8,
               timeit(lambda: chol.solve(rhs)))
        report(f"btd.solve bwd B={b} T={t} D=3 [{label}]", b * t, (2 * 9 + 2 * 3) * 8,
               timeit(lambda: chol.solve(rhs, transpose_left=True)))
        report(f"btd.block_diagonal_of_inverse B={b} T={t} D=3 [{label}]", b * t, (3 * 9) * 8,
               timeit(lambda: chol.block_diagonal_of_inverse()))
        report(f"btd.upper_diagonal_lower B={b} T={t} D=3 [{label}]", b * t, (4 * 9) * 8,
               timeit(lambda: m.upper_diagonal_lower()))
    lib.mf_set_tuning(4, 0)


def fewchains():
    """Cholesky + solve of FEW long chains: parallel in time (btd_pit.cuh) vs the sequential sweep."""
    lib = _lib.lib()
    for b, t in ((1, 10_000_000), (1, 1_000_000), (8, 100_000), (64, 10_000), (256, 10_000), (1024, 10_000)):
        diag, sub, rhs = bench_inputs.matern52_posterior_precision(b, t, DEV, chunk=min(b, 64))
        m = mf.SymmetricBlockTriDiagonal(diag, sub)
        for knob, label in ((0, "parallel in time"), (1, "sequential sweep")):
            if knob == 1 and b * t >= 10_000_000 and b == 1:
                continue  # 3 s per call
            lib.mf_set_tuning(2, knob)
            ms = timeit(lambda: m.cholesky_and_solve(rhs), warm=2, reps=5)
            report(f"cholesky+solve B={b} T={t} D=3 f64 [{label}]", b * t, 336, ms)
        lib.mf_set_tuning(2, 0)
        del diag, sub, rhs, m
        torch.cuda.empty_cache()


def gpr1():
    """BASELINE config 1 scaled up: ONE Matern32 series, KalmanFilter log-likelihood + posterior SSM +
    posterior marginals (kalman_filter.py:109-255); parallel in time vs sequential sweeps."""
    lib = _lib.lib()
    for t in (1_000, 100_000, 1_000_000, 10_000_000):
        ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t, DEV)
        kf = mf.KalmanFilter(ssm, mf.EmissionModel(h), y, lr)

        def job():
            ll = kf.log_likelihood()
            post = kf.posterior_state_space_model()
            return ll, post.marginals

        if t <= 100_000:
            g = mf.Graphed(job)
            ms = timeit(g, warm=2, reps=10)
            report(f"GPR single series T={t} D=2 f64: log-lik + posterior SSM + marginals [CUDA graph replay]", t,
                   (2 * 4 + 2 + 2 + 1 + 3 * 4 + 2 * 2) * 8, ms)
            ll_e, (m_e, c_e) = job()
            ll_g, (m_g, c_g) = g()
            assert torch.allclose(ll_e, ll_g) and torch.allclose(m_e, m_g) and torch.allclose(c_e, c_g)
        for knob, label in ((0, "parallel in time"), (1, "sequential sweeps")):
            if knob == 1 and t > 1_000_000:
                continue  # ~10 s per call
            lib.mf_set_tuning(2, knob)
            ms = timeit(job, warm=2, reps=5)
            report(f"GPR single series T={t} D=2 f64: log-lik + posterior SSM + marginals [{label}]", t,
                   (2 * 4 + 2 + 2 + 1 + 3 * 4 + 2 * 2) * 8, ms)
        lib.mf_set_tuning(2, 0)
        del ssm, h, y, kf
        torch.cuda.empty_cache()


def config4(b=256, t=10_000):
    """Config 4 at reduced T (the full T=1e5 needs 118 GB of inputs: in-place, 8 GPUs or B-chunks)."""
    diag, sub, rhs = bench_inputs.sum_kernel_posterior_precision(b, t, DEV)
    d0, s0 = diag.clone(), sub.clone()
    x = torch.empty_like(rhs)
    info = torch.empty(b, dtype=torch.int32, device=DEV)
    lib = _lib.lib()

    def step():
        diag.copy_(d0)
        sub.copy_(s0)

    def run():
        _lib.check(lib.mf_btd_cholesky(_lib.MF_F64, _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(rhs),
                                       _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(x), None, _lib.ptr(info),
                                       _lib.i64(b), _lib.i64(t), _lib.i64(17), _lib.current_stream()), "chol")

    # in-place: restore the inputs before every timed call (restore not timed)
    ts = []
    for i in range(5):
        step()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z))
    assert int(info.abs().max()) == 0
    ms = sorted(ts[1:])[len(ts[1:]) // 2]
    report(f"cholesky+solve config4 B={b} T={t} D=17 f64 in-place [warp per chain]", b * t,
           (4 * 289 + 2 * 17) * 8, ms)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kalman3", "kalman_batch", "cvi5", "ssm"]
    for w in which:
        globals()[w]()
        torch.cuda.empty_cache()
