"""Config 3 (one Matern32 series, T = 1e7, f64 Kalman log-likelihood) for ncu: warm-up, then profiled calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf

if __name__ == "__main__":
    t = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda:0")
    ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t, dev)
    kf = mf.KalmanFilter(ssm, mf.EmissionModel(h), y, lr)
    for _ in range(3):
        kf.log_likelihood()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(2):
        kf.log_likelihood()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
