"""Config 4 (D=17, B=256, f64, in place) at T steps for ncu: one warm-up call, then one profiled call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
from tools.config4_full import chol

if __name__ == "__main__":
    t = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda:0")
    for it in range(2):
        d, s, r = bench_inputs.sum_kernel_posterior_precision(b, t, dev)
        x = torch.empty_like(r)
        info = torch.empty(b, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.cudart().cudaProfilerStart()
        chol(d, s, r, d, s, x, info, b, t)
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.cudart().cudaProfilerStop()
        del d, s, r, x
