cat > /tmp/c4.py <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
import bench_inputs
from tools.config4_full import chol
from markovflow_b200 import _lib
dev = torch.device("cuda:0")
b, t, knob = 256, int(sys.argv[1]), int(sys.argv[2])
d, s, r = bench_inputs.sum_kernel_posterior_precision(b, t, dev)
d0 = None
x = torch.empty_like(r); info = torch.empty(b, dtype=torch.int32, device=dev)
_lib.lib().mf_set_tuning(7, knob)
for i in range(3):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chol(d, s, r, d, s, x, info, b, t)
    w1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print(t, knob, "call", i, "event ms %.2f" % e0.elapsed_time(e1), "host launch ms %.2f" % ((w1 - w0) * 1e3), "info", int(info.abs().max()))
PY
python /tmp/c4.py 20000 34
python /tmp/c4.py 100000 0
