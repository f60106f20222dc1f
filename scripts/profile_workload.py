"""Profiling target for the other BASELINE configurations:
    ncu --set full -k regex:stage -s 4 -c 2 ... python scripts/profile_workload.py c2|c4|c3 [n] [steps] [T] [json options]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jams_b200 import workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
T = float(sys.argv[4]) if len(sys.argv) > 4 else 300.0
extra = json.loads(sys.argv[5]) if len(sys.argv) > 5 else {}
w = {"c2": W.c2_bcc_fe, "c4": W.c4_bcc_long_range, "c3": W.c3_sc}[which[:2]](n, temperature=T)
if which.endswith("rk4"):   # the RK4 solver on the same workload (stage ms: stages 1 + 2, stages 3 + 4)
    from jams_b200.solver import create_hamiltonian, create_solver
    s = create_solver(dict(module="llg-rk4-b200-gpu", t_step=W.T_STEP, t_max=1e-9, seed=3, options=dict(extra, verbose=1, time_kernels=1)), w["lattice"])
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, w["lattice"]))
    s.set_temperature(T)
    s.set_spins(w["lattice"].initial_spins(seed=1))
else:
    s = W.make_solver(w, options=dict(extra, verbose=1, time_kernels=1), random_spins_seed=1, seed=3)
s.run(2)   # warm-up: lazy module loading, first-launch attribute calls
s.ctx.synchronize()
s.ctx.last_step_kernel_ms()
s.run(steps)
s.ctx.synchronize()
ms = s.ctx.last_step_kernel_ms() / steps
print("done", which, n, steps, s.ctx.kernel_launches(), "stage ms", ms, "->", w["lattice"].num_spins / (ms.sum() * 1e-3) / 1e9, "G upd/s")
