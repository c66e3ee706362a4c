"""A few launches of the Matern-prior Kalman log-likelihood (in-kernel SSM) on the
config-3 series, for ncu.  Usage: python tools/profile_matern.py [T] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf

if __name__ == "__main__":
    t = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    dts = bench_inputs.matern32_time_deltas(1, t, dev)
    y = torch.randn(1, t, generator=g, dtype=torch.float64, device=dev)
    one = torch.ones(1, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(reps):
        ll = mf.matern_kalman_log_likelihood(2, one, one, y, 0.1, time_deltas=dts)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(float(ll[0]))
