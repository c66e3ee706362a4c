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
names = {0: "C64 K16 S2", 5: "C96 K8 S2", 6: "C32 K8 S3", 7: "C128 K4 S3", 8: "C192 K4 S2", 9: "C64 K4 S3",
         10: "C32 K4 S3", 11: "C64 K8 S2", 12: "C160 K4 S2"}
cpw = {0: 64, 5: 96, 6: 64, 7: 128, 8: 192, 9: 128, 10: 128, 11: 128, 12: 160}
for knob, name in names.items():
    for waves in (1, 2, 3):
        lib.mf_set_tuning(5, knob)
        L = -(-t // (148 * cpw[knob] * waves))
        L = (L + 31) // 32 * 32
        lib.mf_set_tuning(3, L)
        lib.mf_set_tuning(2, 2)
        try:
            ms = timeit(lambda: mf.kalman_log_likelihood(ssm, h, y, lr))
            ll = float(mf.kalman_log_likelihood(ssm, h, y, lr))
            print(json.dumps({"cfg": name, "waves": waves, "L": L, "ms": round(ms, 4), "ll": ll}), flush=True)
        except Exception as e:
            print(name, "failed:", e, flush=True)
lib.mf_set_tuning(5, 0); lib.mf_set_tuning(3, 0); lib.mf_set_tuning(2, 0)
