"""Config-5 transforms (B=1024 x M=1e4, D=2): time per call for the 1-D engine (knob 13 = 1) and every tile
geometry of the tensor-map engine (knob 14; 0 = in-place stages, 8 = separate rings)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf
from markovflow_b200 import _lib


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    for dtype in (torch.float64, torch.float32):
        th = tuple(x.to(dtype) for x in bench_inputs.cvi_naturals_config5(B, T, dev, dtype=torch.float64))
        got = mf.naturals_to_ssm_params(*th)
        q = mf.StateSpaceModel(*(g.contiguous() for g in (got[4], got[2], got[0], got[1], got[3])))
        for k13, k14, k3 in [(1, 0, 0), (0, 0, 0), (0, 8, 0)] + [(0, 0, int(x)) for x in os.environ.get('C5_SEGS', '').split(',') if x]:
            lib.mf_set_tuning(13, k13)
            lib.mf_set_tuning(14, k14)
            lib.mf_set_tuning(3, k3)
            t_nat = timed(lambda: mf.naturals_to_ssm_params(*th))
            t_exp = timed(lambda: mf.ssm_to_expectations(q))
            t_mar = timed(lambda: q.marginals)  # q's parameters are dense: no compaction copy inside the timing
            es = 8 if dtype == torch.float64 else 4
            gb = B * T * 20 * es / 1e9
            print(f"{str(dtype):14s} knob13={k13} knob14={k14} seg={k3}  nat->ssm {t_nat:.3f} ms ({gb / t_nat:.2f} TB/s)  "
                  f"ssm->exp {t_exp:.3f} ms ({gb / t_exp:.2f} TB/s)  marginals {t_mar:.3f} ms", flush=True)
        lib.mf_set_tuning(13, 0)
        lib.mf_set_tuning(14, 0)
        lib.mf_set_tuning(3, 0)
