"""Development check of the team (CTA-per-chain) large-block Cholesky kernel against the oracle and
the warp-per-chain kernel, plus timing on config-4 shapes:  python tools/team_check.py [time]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_inputs  # noqa: E402
from markovflow_b200 import _lib  # noqa: E402
from markovflow_b200._lib import check, current_stream, i64, ptr  # noqa: E402
from oracle import np_oracle as O  # noqa: E402
from tests.test_gpu_big_blocks import random_well_conditioned_spd_btd  # noqa: E402

DEV = torch.device("cuda:0")
lib = _lib.lib()


def run(diag, sub, rhs, dtype=torch.float64, inplace=False, variant=2, offset=0):
    """offset: shift all arrays by `offset` elements inside a bigger allocation (misalignment)."""
    b, t, d = diag.shape[:3]

    def dev(x):
        if x is None:
            return None
        flat = torch.empty(x.size + offset, dtype=dtype, device=DEV)
        v = flat[offset:].view(x.shape)
        v.copy_(torch.as_tensor(x).to(dtype))
        return v

    gd, gs, gr = dev(diag), dev(sub), dev(rhs)
    if inplace:
        od, os_, ox = gd, gs, gr
    else:
        od, os_, ox = dev(np.full_like(diag, np.nan)), dev(np.full_like(sub, np.nan)), (
            dev(np.full_like(rhs, np.nan)) if rhs is not None else None)
    logdet = torch.empty(b, dtype=dtype, device=DEV)
    info = torch.empty(b, dtype=torch.int32, device=DEV)
    lib.mf_set_tuning(7, variant)
    code = _lib.MF_F64 if dtype == torch.float64 else _lib.MF_F32
    check(lib.mf_btd_cholesky(code, ptr(gd), ptr(gs), ptr(gr), ptr(od), ptr(os_), ptr(ox), ptr(logdet),
                              ptr(info), i64(b), i64(t), i64(d), current_stream()), "chol")
    torch.cuda.synchronize()
    lib.mf_set_tuning(7, 0)
    f = lambda x: None if x is None else x.cpu().numpy().astype(np.float64)
    return f(od), f(os_), f(ox), f(logdet), info.cpu().numpy()


def err(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def parity():
    worst = 0.0
    for d in (17,):
        for t in (2, 3, 4, 5, 7, 8, 9, 33, 100):
            for (dtype, tol) in ((torch.float64, 1e-10), (torch.float32, 1e-4)):
                for offset in (0, 1):
                    for with_rhs in (True, False):
                        for inplace in (False, True):
                            b = 5
                            diag, sub, _, _ = random_well_conditioned_spd_btd((b,), t, d, rng=d * 100 + t)
                            rhs = np.random.default_rng(t).standard_normal((b, t, d)) if with_rhs else None
                            o_ld, o_ls = O.btd_cholesky(diag, sub)
                            gd, gs, gx, gl, info = run(diag, sub, rhs, dtype, inplace, 2, offset)
                            e = max(err(gd, o_ld), err(gs, o_ls), err(gl, O.btd_abs_log_det(o_ld)))
                            if with_rhs:
                                e = max(e, err(gx, O.btd_solve(o_ld, o_ls, rhs)))
                            assert np.all(np.triu(gd, 1) == 0.0)
                            assert info.max() == 0
                            worst = max(worst, e / tol)
                            assert e < tol, (d, t, dtype, offset, with_rhs, inplace, e)
    print("parity ok, worst err/tol", worst)
    # failure report
    diag, sub, _, _ = random_well_conditioned_spd_btd((3,), 12, 17, rng=1)
    bad = diag.copy()
    bad[1, 4] = -bad[1, 4]
    *_, info = run(bad, sub, None)
    print("info on failure", info)
    assert info[1] == 5 and info[0] == 0 and info[2] == 0


def config4_parity():
    diag, sub, rhs = bench_inputs.sum_kernel_posterior_precision(8, 500, DEV)
    d_, s_, r_ = diag.cpu().numpy(), sub.cpu().numpy(), rhs.cpu().numpy()
    gd, gs, gx, gl, info = run(d_, s_, r_, variant=2)
    wd, ws, wx, wl, _ = run(d_, s_, r_, variant=1)
    o_ld, o_ls = O.btd_cholesky(d_, s_)
    print("config4 inputs: team vs oracle Ld %.2e Ls %.2e x %.2e logdet %.2e | warp vs oracle Ld %.2e x %.2e" % (
        err(gd, o_ld), err(gs, o_ls), err(gx, O.btd_solve(o_ld, o_ls, r_)), err(gl, O.btd_abs_log_det(o_ld)),
        err(wd, o_ld), err(wx, O.btd_solve(o_ld, o_ls, r_))))
    rec_d = gd @ np.swapaxes(gd, -1, -2)
    rec_d[:, 1:] += gs @ np.swapaxes(gs, -1, -2)
    print("  reconstruction err %.2e" % err(np.tril(rec_d), np.tril(d_)))


def timing(b=256, t=4000):
    diag, sub, rhs = bench_inputs.sum_kernel_posterior_precision(b, t, DEV)
    d = diag.shape[-1]
    od, os_, ox = torch.empty_like(diag), torch.empty_like(sub), torch.empty_like(rhs)
    info = torch.empty(b, dtype=torch.int32, device=DEV)
    for variant in (1, 2):
        lib.mf_set_tuning(7, variant)
        fn = lambda: check(lib.mf_btd_cholesky(_lib.MF_F64, ptr(diag), ptr(sub), ptr(rhs), ptr(od), ptr(os_),
                                               ptr(ox), None, ptr(info), i64(b), i64(t), i64(d),
                                               current_stream()), "chol")
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 5
        sps = b * t / (ms * 1e-3)
        print(f"variant {variant}: B={b} T={t} D={d}: {ms:.3f} ms, {sps:.3e} state-steps/s, "
              f"{sps * 9520 / 1e9:.0f} GB/s, cycles/step/chain @1.965GHz = {1.965e9 * ms * 1e-3 / t:.0f}")
    lib.mf_set_tuning(7, 0)


if __name__ == "__main__":
    parity()
    config4_parity()
    if "time" in sys.argv:
        timing()
        timing(b=296, t=2000)
        timing(b=128, t=4000)
