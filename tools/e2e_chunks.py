"""End-to-end (host buffers) Cholesky + solve on config 2 for several chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_inputs
from markovflow_b200.host import cholesky_solve_host

dev = torch.device("cuda:0")
b, t = 4096, 10000
diag, sub, rhs = bench_inputs.matern52_posterior_precision(b, t, dev)
hd = torch.empty(diag.shape, dtype=diag.dtype, pin_memory=True); hd.copy_(diag)
hs = torch.empty(sub.shape, dtype=sub.dtype, pin_memory=True); hs.copy_(sub)
hr = torch.empty(rhs.shape, dtype=rhs.dtype, pin_memory=True); hr.copy_(rhs)
del diag, sub, rhs
out = (torch.empty(hd.shape, dtype=hd.dtype, pin_memory=True), torch.empty(hs.shape, dtype=hs.dtype, pin_memory=True),
       torch.empty(hr.shape, dtype=hr.dtype, pin_memory=True), torch.empty(b, dtype=torch.int32, pin_memory=True))
for chunk in (64, 128, 256, 512, 1024):
    for _ in range(2):
        cholesky_solve_host(hd, hs, hr, out=out, chunk=chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 4
    for _ in range(n):
        cholesky_solve_host(hd, hs, hr, out=out, chunk=chunk)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"chunk {chunk}: {ms:.1f} ms per batch, {b * t / ms * 1e3:.3e} state-steps/s, "
          f"{2 * 6.881 / ms * 1e3:.1f} GB/s over PCIe (both directions)", flush=True)
