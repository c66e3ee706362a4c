"""A/B timings of mf_kalman_matern_log_likelihood (in-kernel SSM construction) on the
config-3 series: launch variant (knob 8) x virtual chains per SM (knob 9), against the materialised-SSM
kernel.  Usage: python tools/matern_bench.py [T]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf
from markovflow_b200 import _lib


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    t = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda:0")
    ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t, dev)
    ref = float(mf.kalman_log_likelihood(ssm, h, y, lr))
    ms_mat = timed(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
    print(json.dumps({"path": "materialised SSM", "ms": ms_mat, "loglik": ref}))
    dts = bench_inputs.matern32_time_deltas(1, t, dev)
    y2 = y.reshape(1, t).contiguous()
    one = torch.ones(1, dtype=torch.float64, device=dev)
    lib = _lib.lib()
    for variant in (0, 5, 6, 1, 4):
        for wps in {0: (8, 12, 16), 5: (12, 16, 20), 6: (12, 16, 20, 24)}.get(variant, (4,)):
            lib.mf_set_tuning(8, variant)
            lib.mf_set_tuning(9, wps)
            fn = lambda: mf.matern_kalman_log_likelihood(2, one, one, y2, lr, time_deltas=dts)
            try:
                ll = float(fn()[0])
                ms = timed(fn)
                print(json.dumps({"variant": variant, "warps_per_sm": wps, "ms": round(ms, 4),
                                  "rel_diff": abs(ll - ref) / abs(ref), "speedup": round(ms_mat / ms, 2)}))
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"variant": variant, "warps_per_sm": wps, "error": str(e)}))
    lib.mf_set_tuning(8, 0)
    lib.mf_set_tuning(9, 0)
    # f32 and the other orders at the default geometry
    for d in (1, 3):
        fn = lambda: mf.matern_kalman_log_likelihood(d, one, one, y2, lr, time_deltas=dts)
        print(json.dumps({"state_dim": d, "ms": round(timed(fn), 4)}))
    y32, dt32, one32 = y2.float(), dts.float(), one.float()
    fn = lambda: mf.matern_kalman_log_likelihood(2, one32, one32, y32, lr.float(), time_deltas=dt32)
    ll = float(fn()[0])
    print(json.dumps({"dtype": "f32", "ms": round(timed(fn), 4), "rel_diff": abs(ll - ref) / abs(ref)}))


if __name__ == "__main__":
    main()
