set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu24.log; tail -25 gpurun_out/pytest_gpu24.log
python - <<'PY'
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from jams_b200 import workloads as W
from jams_b200.solver import create_solver, create_hamiltonian
w = W.c3_sc(dims=(256, 256, 256), temperature=100.0)
lat = w["lattice"]
for mod in ("llg-rk4-b200-gpu", "llg-heun-b200-gpu"):
    s = create_solver(dict(module=mod, t_step=W.T_STEP, t_max=1e-9, seed=1, options=dict(time_kernels=0)), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.set_temperature(100.0)
    s.set_spins(lat.initial_spins(seed=1))
    s.run(5); s.ctx.synchronize()
    t0 = time.perf_counter(); s.run(50); s.ctx.synchronize(); dt = (time.perf_counter() - t0) / 50
    print(f"{mod}: {dt*1e3:.3f} ms/step -> {lat.num_spins/dt/1e9:.2f} G spin-updates/s (RK4: 408 B/update model -> {408*lat.num_spins/dt/1e12:.2f} TB/s)", flush=True)
    s.ctx.close()
PY
