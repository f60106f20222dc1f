"""Profiling target: a few Heun steps of the bench workload (C3 sc 256^3, T = 100 K) for ncu.
    ncu --set full -k regex:stage -s 4 -c 2 ... python scripts/profile_step.py [n] [steps] [kernel] [T] [json options]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jams_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 1
T = float(sys.argv[4]) if len(sys.argv) > 4 else 100.0
import json
extra = json.loads(sys.argv[5]) if len(sys.argv) > 5 else {}
dims = tuple(int(v) for v in os.environ["JB_QB_DIMS"].split("x")) if os.environ.get("JB_QB_DIMS") else (n, n, n)
w = W.c3_sc(dims=dims, temperature=T)
s = W.make_solver(w, options=dict(extra, kernel=kernel), random_spins_seed=1, seed=3)
s.run(steps)
s.ctx.synchronize()
print("done", n, steps, kernel, s.ctx.kernel_launches())
