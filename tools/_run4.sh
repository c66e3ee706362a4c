python tools/big_throughput.py | tail -1
python /tmp/c4.py 2>/dev/null
cat > /tmp/c4.py <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
import bench_inputs
from tools.config4_full import chol
from markovflow_b200 import _lib
dev = torch.device("cuda:0")
b, t, knob = 256, int(sys.argv[1]), int(sys.argv[2])
d, s, r = bench_inputs.sum_kernel_posterior_precision(b, t, dev)
x = torch.empty_like(r); info = torch.empty(b, dtype=torch.int32, device=dev)
_lib.lib().mf_set_tuning(7, knob)
for i in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chol(d, s, r, d, s, x, info, b, t)
    e1.record()
    torch.cuda.synchronize()
    print(t, knob, "call", i, "event ms %.2f" % e0.elapsed_time(e1), "steps/s %.3e" % (b * t / e0.elapsed_time(e1) * 1e3))
PY
python /tmp/c4.py 20000 0
