"""Per-stage clock64 timeline of one step of the team kernel (library built with -DMF_TEAM_DEBUG)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tools.team_check as tc  # noqa: E402

tc.timing(b=int(sys.argv[1]) if len(sys.argv) > 1 else 256, t=400)
buf = (ctypes.c_longlong * 160)()
tc.lib.mf_set_tuning(7, 2)
rc = tc.lib.mf_debug_team_timeline(buf)
a = np.array(buf[:]).reshape(4, 40)
t0 = a[1, 0]
print("F (W1) stage ends rel. to F start:", (a[1, 1:18] - t0).tolist())
print("A0 start, stage ends:", (a[2, 0] - t0), (a[2, 1:18] - t0).tolist())
print("A1 start, stage ends:", (a[3, 0] - t0), (a[3, 1:18] - t0).tolist())
print("N (W0) stage ends:", (a[0, 20:37] - t0).tolist())
print("next F start:", a[0, 0] - t0)
