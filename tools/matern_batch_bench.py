"""Many independent series through mf_kalman_matern_log_likelihood (one chain per series, plain filter
core): B x T state-steps per launch.  Usage: python tools/matern_batch_bench.py [B] [T]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import markovflow_b200 as mf
from tools.matern_bench import timed

if __name__ == "__main__":
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(71892305)
    ls = 0.5 + 1.5 * torch.rand(b, generator=g, dtype=torch.float64, device=dev)
    var = 0.5 + 1.5 * torch.rand(b, generator=g, dtype=torch.float64, device=dev)
    dts = 0.05 + 0.1 * torch.rand(b, t - 1, generator=g, dtype=torch.float64, device=dev)
    y = torch.randn(b, t, generator=g, dtype=torch.float64, device=dev)
    for dtype in (torch.float64, torch.float32):
        for d in (1, 2, 3):
            args = (d, ls.to(dtype), var.to(dtype), y.to(dtype), 0.1)
            kw = dict(time_deltas=dts.to(dtype))
            fn = lambda: mf.matern_kalman_log_likelihood(*args, **kw)
            ll = fn()
            ms = timed(fn)
            es = 8 if dtype == torch.float64 else 4
            print(json.dumps({"B": b, "T": t, "state_dim": d, "dtype": str(dtype).split(".")[-1], "ms": round(ms, 4),
                              "state_steps_per_s": b * t / (ms * 1e-3), "GBps": b * t * 2 * es / (ms * 1e-3) / 1e9,
                              "finite": bool(torch.isfinite(ll).all())}))
