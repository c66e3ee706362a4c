"""inverse subset / U D U^T at D=2, f64: tensor-map engine vs 1-D copies.  Needs the two cores to opt in (a
`tm_describe` in btd_sweep_cores.cuh: every stream {base, T or T-1, shift 0}); measured once in round 2 and not kept:
B=4096 x T=1e4: 1.84 / 2.39 ms on 1-D copies against 2.30 / 3.15 ms on tensor maps; B=16384 x T=2500: 0.74 / 1.49 against
0.88 / 1.30 ms; B=64 x T=160000 (parallel in time): 0.36 / 0.58 against 0.46 / 0.73 ms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_inputs
import markovflow_b200 as mf
from markovflow_b200 import _lib
from tools.bench_paths import timeit

lib = _lib.lib()
dev = torch.device("cuda:0")
for d, b, t in ((2, 4096, 10_000), (2, 16384, 2500), (2, 64, 160_000)):
    if d == 3:
        diag, sub, rhs = bench_inputs.matern52_posterior_precision(b, t, dev)
    else:
        g = torch.Generator(device=dev); g.manual_seed(1)
        ld = torch.tril(0.3 * torch.randn(b, t, d, d, generator=g, device=dev, dtype=torch.float64)) + 1.5 * torch.eye(d, device=dev, dtype=torch.float64)
        ls = 0.3 * torch.randn(b, t - 1, d, d, generator=g, device=dev, dtype=torch.float64)
        diag = ld @ ld.transpose(-1, -2)
        diag[:, 1:] += ls @ ls.transpose(-1, -2)
        sub = ls @ ld[:, :-1].transpose(-1, -2)
        del ld, ls
    m = mf.SymmetricBlockTriDiagonal(diag, sub)
    chol = m.cholesky
    for knob13, knob14, label in ((1, 0, "1-D copies"), (2, 0, "tensor maps")):
        lib.mf_set_tuning(13, knob13)
        ms1 = timeit(lambda: chol.block_diagonal_of_inverse())
        ms2 = timeit(lambda: m.upper_diagonal_lower())
        gb1, gb2 = b * t * 3 * d * d * 8 / 1e9, b * t * 4 * d * d * 8 / 1e9
        print(f"D={d} B={b} T={t} [{label}] inverse subset {ms1:.3f} ms ({gb1 / ms1:.2f} TB/s)   U D U^T {ms2:.3f} ms ({gb2 / ms2:.2f} TB/s)", flush=True)
    lib.mf_set_tuning(13, 0)
    del m, chol, diag, sub
    torch.cuda.empty_cache()
