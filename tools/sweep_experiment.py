"""Times the config-3 summary sweep under alternative tile configurations (tuning knob 5)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_inputs
import markovflow_b200 as mf
from markovflow_b200 import _lib
from tools.bench_paths import timeit

dev = torch.device("cuda:0")
t = 10_000_000
ssm, h, y, lr = bench_inputs.kalman_inputs_config3(t, dev)
lib = _lib.lib()
names = {0: "default C96 K8 S2", 13: "C96 K10 S2", 14: "C128 K6 S2", 15: "C64 K16 S2", 16: "C64 K12 S2"}
cpw = {0: 96, 13: 96, 14: 128, 15: 64, 16: 64}
for knob, name in names.items():
    for waves in (1,):
        lib.mf_set_tuning(5, knob)
        L = -(-t // (148 * cpw[knob] * waves))
        L = (L + 7) // 8 * 8
        lib.mf_set_tuning(3, L)
        lib.mf_set_tuning(2, 2)
        try:
            ms = timeit(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
            ll = float(mf.kalman_log_likelihood(ssm, h, y, lr))
            print(json.dumps({"cfg": name, "waves": waves, "L": L, "ms": round(ms, 4), "ll": ll}), flush=True)
        except Exception as e:
            print(name, "failed:", e, flush=True)
lib.mf_set_tuning(5, 0); lib.mf_set_tuning(3, 0); lib.mf_set_tuning(2, 0)
