set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu50.log; tail -15 gpurun_out/pytest_gpu50.log
export JB_QB_EXTRA='[{"recover_u":0},{"recover_u":1},{"recover_u":1,"ring_u":3},{"recover_u":1,"ring":5}]'
timeout 900 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench50.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench50.log
