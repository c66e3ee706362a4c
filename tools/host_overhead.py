"""Host-side cost of one naturals_to_ssm_params / ssm_to_expectations call on config 5 (time to ENQUEUE the work with
the failure check off), against the device time of the same call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf

dev = torch.device("cuda:0")
th = bench_inputs.cvi_naturals_config5(1024, 10_000, dev, dtype=torch.float64)
got = mf.naturals_to_ssm_params(*th)
q = mf.StateSpaceModel(*(g.contiguous() for g in (got[4], got[2], got[0], got[1], got[3])))
for name, fn in (("naturals_to_ssm_params", lambda: mf.naturals_to_ssm_params(*th)),
                 ("ssm_to_expectations", lambda: mf.ssm_to_expectations(q))):
    for check in (True, False):
        mf.set_check_numerics(check)
        for _ in range(5):
            fn()
        host = []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            host.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
        host.sort()
        print(f"{name:26s} check_numerics={check}: host time per call (median) {host[len(host) // 2] * 1e6:.0f} us", flush=True)
mf.set_check_numerics(True)
