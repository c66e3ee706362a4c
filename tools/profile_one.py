"""One launch of the fused Cholesky+solve sweep at config-2 shape, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.quick_bench import run
if __name__ == "__main__":
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    run(b, t, d, reps=1)
