set -x
export JB_QB_EXTRA='[{"recover_u":0,"noise_warp":0},{"recover_u":0,"noise_warp":1},{"recover_u":0,"noise_warp":2},{"recover_u":1,"noise_warp":1},{"recover_u":1,"noise_warp":2}]'
timeout 600 python scripts/quick_bench.py 256 100 > gpurun_out/quick_bench51.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench51.log
timeout 600 python - <<'PY' 2>&1 | tail -12
import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from jams_b200 import workloads as W
# noise-warp variants must reproduce the in-thread draws bit for bit (same Philox keys, same arithmetic)
for dims in ((20, 18, 70), (7, 9, 37), (33, 5, 130), (16, 16, 256)):
    w = W.c3_sc(dims=dims, temperature=100.0)
    res = {}
    for ru in (0, 1):
        for nw in (0, 1, 2):
            s = W.make_solver(w, options=dict(kernel=2, recover_u=ru, noise_warp=nw), random_spins_seed=1, seed=3)
            s.run(12)
            res[(ru, nw)] = s.spins()
            s.ctx.close()
    for ru in (0, 1):
        for nw in (1, 2):
            print(dims, "recover_u", ru, "noise_warp", nw, "max diff vs in-thread", float(np.abs(res[(ru, nw)] - res[(ru, 0)]).max()), flush=True)
PY
