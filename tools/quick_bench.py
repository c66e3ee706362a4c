"""Quick device-resident timing of the fused Cholesky+solve sweep (development aid, not bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from markovflow_b200 import _lib

def make(b, t, d, dtype=torch.float64, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ld = 0.3 * torch.tril(torch.randn(b, t, d, d, generator=g, device="cuda", dtype=dtype), -1)
    ld = ld + torch.diag_embed(1.0 + torch.rand(b, t, d, generator=g, device="cuda", dtype=dtype))
    ls = 0.3 * torch.randn(b, t - 1, d, d, generator=g, device="cuda", dtype=dtype)
    diag = ld @ ld.transpose(-1, -2)
    diag[:, 1:] += ls @ ls.transpose(-1, -2)
    sub = ls @ ld[:, :-1].transpose(-1, -2)
    rhs = torch.randn(b, t, d, generator=g, device="cuda", dtype=dtype)
    return diag.contiguous(), sub.contiguous(), rhs

def run(b, t, d, reps=5, dtype=torch.float64):
    diag, sub, rhs = make(b, t, d, dtype)
    od, os_, ox = torch.empty_like(diag), torch.empty_like(sub), torch.empty_like(rhs)
    info = torch.empty(b, dtype=torch.int32, device="cuda")
    L = _lib.lib()
    def call():
        st = L.mf_btd_cholesky(_lib.dtype_code(dtype), _lib.ptr(diag), _lib.ptr(sub), _lib.ptr(rhs), _lib.ptr(od),
                               _lib.ptr(os_), _lib.ptr(ox), None, _lib.ptr(info), _lib.i64(b), _lib.i64(t), _lib.i64(d),
                               _lib.current_stream())
        assert st == 0, st
    for _ in range(2): call()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    es = diag.element_size()
    bytes_ = b * t * (4 * d * d + 2 * d) * es
    print(f"B={b} T={t} D={d} {dtype}: {ms:.3f} ms  {b*t/ms*1e3:.3e} steps/s  {bytes_/ms/1e6:.1f} GB/s "
          f"({bytes_/ms/1e6/6550.1*100:.1f}% of 6550 GB/s)  info_max={int(info.max())}", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for variant, k in ((1, 0), (0, 0), (0, 4)):
        _lib.lib().mf_set_tuning(0, variant); _lib.lib().mf_set_tuning(1, k)
        print("variant", variant, "K", k)
        run(4096, 10000, 3)
        run(4736, 10000, 3)
        run(16384, 2500, 3)
    _lib.lib().mf_set_tuning(0, 0); _lib.lib().mf_set_tuning(1, 0)
    run(4096, 10000, 3)
    run(4096, 2000, 3)
    run(8192, 5000, 3)
    run(32768, 1000, 3)
    run(4096, 10000, 2)
    run(4096, 10000, 3, dtype=torch.float32)
    run(1024, 10000, 5)
    run(512, 5000, 8)
