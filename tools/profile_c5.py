"""Config-5 transforms (B=1024 x M=1e4, D=2) for ncu: naturals_to_ssm_params + ssm_to_expectations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf

if __name__ == "__main__":
    dtype = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    if len(sys.argv) > 3:
        from markovflow_b200 import _lib
        _lib.lib().mf_set_tuning(10, int(sys.argv[3]))
    if len(sys.argv) > 4:
        from markovflow_b200 import _lib
        _lib.lib().mf_set_tuning(11, int(sys.argv[4]))
    dev = torch.device("cuda:0")
    th = tuple(x.to(dtype) for x in bench_inputs.cvi_naturals_config5(1024, 10_000, dev, dtype=torch.float64))
    got = mf.naturals_to_ssm_params(*th)
    q = mf.StateSpaceModel(*(g.contiguous() for g in (got[4], got[2], got[0], got[1], got[3])))
    mf.ssm_to_expectations(q)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(reps):
        mf.naturals_to_ssm_params(*th)
        mf.ssm_to_expectations(q)
        if os.environ.get("C5_MARGINALS"):
            q.marginals
            q.covariance_blocks()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
