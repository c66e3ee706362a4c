"""Throughput of the large-block (D = 17) Cholesky + solve kernel as a function of the number of chains:
latency-bound (256 chains: 2 warps per SM) against throughput-bound (thousands of chains) -- the number that
decides what a parallel-in-time evaluation of config 4 can gain.  Usage: python tools/big_throughput.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
from markovflow_b200 import _lib


def main():
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    d = 17
    base_d, base_s, base_r = bench_inputs.sum_kernel_posterior_precision(64, 1000, dev)
    for b, t in ((256, 1000), (1024, 1000), (2048, 1000), (4096, 1000), (8192, 500)):
        reps = (b + 63) // 64
        diag = base_d[:, :t].repeat(reps, 1, 1, 1)[:b].contiguous()
        sub = base_s[:, :t - 1].repeat(reps, 1, 1, 1)[:b].contiguous()
        rhs = base_r[:, :t].repeat(reps, 1, 1)[:b].contiguous()
        od, os_, ox = torch.empty_like(diag), torch.empty_like(sub), torch.empty_like(rhs)
        info = torch.empty(b, dtype=torch.int32, device=dev)

        def run():
            _lib.check(lib.mf_btd_cholesky(_lib.MF_F64, _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(rhs), _lib.ptr(od),
                                           _lib.ptr(os_), _lib.ptr(ox), None, _lib.ptr(info), _lib.i64(b), _lib.i64(t),
                                           _lib.i64(d), _lib.current_stream()), "chol")
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"chains": b, "T": t, "ms": round(ms, 3), "state_steps_per_s": b * t / (ms * 1e-3),
                          "GBps": b * t * 9520 / (ms * 1e-3) / 1e9, "cycles_per_step_per_sm": ms * 1e-3 * 1.965e9 * 148 / (b * t)}))
        del diag, sub, rhs, od, os_, ox


if __name__ == "__main__":
    main()
