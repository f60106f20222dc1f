set -x
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu57.log; tail -5 gpurun_out/pytest_gpu57.log
export JB_QB_EXTRA='[{"recover_u":0}]'
timeout 120 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench57.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench57.log
