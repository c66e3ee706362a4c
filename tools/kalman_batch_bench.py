"""Materialised-SSM Kalman log-likelihood on mid-size batches: cut in time (default plan) against one
chain per series (tuning knob 2 = 1).  Usage: python tools/kalman_batch_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf
from markovflow_b200 import _lib
from tools.matern_bench import timed

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    for b, t in ((256, 10_000), (1024, 10_000), (4096, 10_000), (8192, 4_000)):
        mu0, l0, a, off, lq, h = bench_inputs.matern32_ssm(b, t, dev, jitter_hyper=True)
        ssm = mf.StateSpaceModel(mu0, l0, a, off, lq)
        y = torch.randn(b, t, 1, dtype=torch.float64, device=dev)
        lr = torch.tensor([[0.1]], dtype=torch.float64, device=dev)
        fn = lambda: mf.kalman_log_likelihood(ssm, h[0], y, lr)
        res = {}
        for knob in (0, 1):
            lib.mf_set_tuning(2, knob)
            ll = fn()
            res[knob] = (timed(fn), ll)
        lib.mf_set_tuning(2, 0)
        err = float((res[0][1] - res[1][1]).abs().max() / res[1][1].abs().max())
        print(json.dumps({"B": b, "T": t, "ms_default_plan": round(res[0][0], 4), "ms_one_chain_per_series": round(res[1][0], 4),
                          "GBps_default": b * t * 104 / (res[0][0] * 1e-3) / 1e9, "max_rel_diff": err}))
        del ssm, mu0, l0, a, off, lq, h, y
        torch.cuda.empty_cache()
