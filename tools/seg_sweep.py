import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch, bench_inputs, markovflow_b200 as mf
from markovflow_b200 import _lib
from bench_paths import timeit, DEV
lib = _lib.lib()
th = bench_inputs.cvi_naturals_config5(1024, 10000, DEV, dtype=torch.float64)
p = mf.naturals_to_ssm_params(*th)
q = mf.StateSpaceModel(p[4], p[2], p[0], p[1], p[3])
for seg, geo in ((0, 0), (80, 0), (120, 0), (200, 0), (358, 0), (1000, 0)):
    lib.mf_set_tuning(3, seg)
    lib.mf_set_tuning(5, geo)
    a = timeit(lambda: mf.naturals_to_ssm_params(*th))
    b = timeit(lambda: mf.ssm_to_expectations(q))
    c = timeit(lambda: q.marginals)
    print(f"seg {seg} geometry knob {geo}: nat_to_ssm {a:.3f} ms, to_expectations {b:.3f} ms, marginals {c:.3f} ms", flush=True)
lib.mf_set_tuning(3, 0)
lib.mf_set_tuning(5, 0)
