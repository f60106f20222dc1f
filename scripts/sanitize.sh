# compute-sanitizer over a small run of every stage-kernel generation (memcheck: out-of-bounds / misaligned accesses in
# global and shared memory; racecheck: shared-memory hazards of the fused kernel's s* ring and the entry tables)
set -x
for K in 2 3 1; do
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanity_tile.py "{\"kernel\": $K}" 50.0 > gpurun_out/memcheck_k$K.log 2>&1; echo "memcheck kernel $K rc=$?"; tail -3 gpurun_out/memcheck_k$K.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from jams_b200 import workloads as W
from jams_b200.solver import create_solver, create_hamiltonian
w = W.c2_bcc_fe(6, temperature=50.0); lat = w['lattice']
for mod in ('llg-rk4-b200-gpu', 'llg-heun-b200-gpu'):
    for k in (2, 3):
        s = create_solver(dict(module=mod, t_step=1e-16, t_max=1e-9, options=dict(kernel=k)), lat)
        for h in w['hamiltonians']: s.register_hamiltonian(create_hamiltonian(h, lat))
        s.set_temperature(50.0); s.set_spins(lat.initial_spins(seed=1)); s.run(4); print(mod, k, np.abs(s.spins()).max())
" > gpurun_out/memcheck_bcc.log 2>&1; echo "memcheck bcc rc=$?"; tail -6 gpurun_out/memcheck_bcc.log
for K in 3 2; do
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanity_tile.py "{\"kernel\": $K}" 0.0 > gpurun_out/racecheck_k$K.log 2>&1; echo "racecheck kernel $K rc=$?"; grep -c "hazard" gpurun_out/racecheck_k$K.log; tail -4 gpurun_out/racecheck_k$K.log
done
