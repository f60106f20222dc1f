set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu21.log; tail -15 gpurun_out/pytest_gpu21.log
export JB_QB_EXTRA='[{"kernel":2},{"kernel":2,"tile_y":8,"tile_z":64},{"kernel":2,"spt":2,"tile_y":8,"tile_z":128},{"kernel":3}]'
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench21.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench21.log
