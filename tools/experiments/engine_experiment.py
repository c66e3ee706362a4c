"""A/B timing of the sweep2 (LDGSTS) engine variants against the default paths (tuning knob 6)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_inputs
import markovflow_b200 as mf
from markovflow_b200 import _lib
from tools.bench_paths import timeit

dev = torch.device("cuda:0")
lib = _lib.lib()
which = sys.argv[1:] or ["chol", "kalman"]

if "chol" in which:
    b, t, d = 4096, 10000, 3
    diag, sub, rhs = bench_inputs.matern52_posterior_precision(b, t, dev)
    od, os_, ox = torch.empty_like(diag), torch.empty_like(sub), torch.empty_like(rhs)
    info = torch.empty(b, dtype=torch.int32, device=dev)

    def run():
        _lib.check(lib.mf_btd_cholesky(_lib.MF_F64, _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(rhs), _lib.ptr(od),
                                       _lib.ptr(os_), _lib.ptr(ox), None, _lib.ptr(info), _lib.i64(b),
                                       _lib.i64(t), _lib.i64(d), _lib.current_stream()), "chol")

    lib.mf_set_tuning(6, 0)
    run(); torch.cuda.synchronize()
    ref = (od.clone(), os_.clone(), ox.clone())
    print(json.dumps({"chol": "default TMA ring", "ms": round(timeit(run), 4)}), flush=True)
    for v in range(7, 14):
        lib.mf_set_tuning(6, v + 1)
        try:
            od.zero_(); os_.zero_(); ox.zero_()
            run(); torch.cuda.synchronize()
            same = all(torch.equal(a, r) for a, r in zip((od, os_, ox), ref))
            err = max(float((a - r).abs().max()) for a, r in zip((od, os_, ox), ref))
            print(json.dumps({"chol": f"sweep2 variant {v}", "ms": round(timeit(run), 4), "bit_equal": same,
                              "max_abs_diff": err, "info": int(info.abs().max())}), flush=True)
        except Exception as e:
            print("variant", v, "failed:", str(e)[:100], flush=True)
    lib.mf_set_tuning(6, 0)
    del diag, sub, rhs, od, os_, ox, ref
    torch.cuda.empty_cache()

if "kalman" in which:
    t = 10_000_000
    ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t, dev)
    lib.mf_set_tuning(6, 0)
    ll0 = float(mf.kalman_log_likelihood(ssm, h, y, lr))
    print(json.dumps({"kalman": "default", "ms": round(timeit(lambda: mf.kalman_log_likelihood(ssm, h, y, lr)), 4), "ll": ll0}), flush=True)
    cpw = {0: 128, 1: 160, 2: 96, 3: 192, 4: 128, 5: 64, 6: 128}
    for v in range(7):
        lib.mf_set_tuning(6, v + 1)
        L = -(-t // (148 * cpw[v]))
        L = (L + 31) // 32 * 32
        lib.mf_set_tuning(3, L)
        lib.mf_set_tuning(2, 2)
        try:
            ll = float(mf.kalman_log_likelihood(ssm, h, y, lr))
            ms = timeit(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
            print(json.dumps({"kalman": f"sweep2 variant {v}", "L": L, "ms": round(ms, 4), "ll": ll,
                              "rel": abs(ll - ll0) / abs(ll0)}), flush=True)
        except Exception as e:
            print("variant", v, "failed:", str(e)[:100], flush=True)
    lib.mf_set_tuning(6, 0); lib.mf_set_tuning(3, 0); lib.mf_set_tuning(2, 0)
