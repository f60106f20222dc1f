set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu30.log; tail -6 gpurun_out/pytest_gpu30.log
timeout 900 python scripts/other_configs.py > gpurun_out/other_configs30.log 2>&1; cat gpurun_out/other_configs30.log
python - <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from jams_b200 import workloads as W
import other_configs as oc
for opts in (dict(tile_y=8, tile_z=32), dict(tile_y=4, tile_z=64), dict(tile_y=2, tile_z=128), dict(kernel=3), dict(kernel=0), dict(kernel=1)):
    oc.run("C2 bcc 128^3 T=300 " + str(opts), W.c2_bcc_fe(128, temperature=300.0), 30, dict(opts, verbose=1))
PY
