"""Config 4 (D=17, B=256, f64, in place): time of one mf_btd_cholesky call against the number of time segments
(tuning knob 7; 0 = the library's plan)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
from markovflow_b200 import _lib
from tools.config4_full import chol

if __name__ == "__main__":
    t = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    b = 256
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    for knob in [0, 6, 8, 9, 10, 11, 12, 14, 18, 20]:
        lib.mf_set_tuning(7, knob)
        best = 1e9
        for it in range(3):
            d, s, r = bench_inputs.sum_kernel_posterior_precision(b, t, dev)
            x = torch.empty_like(r)
            info = torch.empty(b, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            chol(d, s, r, d, s, x, info, b, t)
            e1.record()
            torch.cuda.synchronize()
            if it > 0:
                best = min(best, e0.elapsed_time(e1))
            del d, s, r, x
        print(f"T={t} segments knob={knob}: {best:.2f} ms  ({b * t / best / 1e3:.3e} state-steps/s)", flush=True)
    lib.mf_set_tuning(7, 0)
