set -x
export JB_QB_EXTRA='[{"kernel":2},{"kernel":2,"tile_y":8,"tile_z":64}]'
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench22.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench22.log; grep "stage 1" gpurun_out/quick_bench22.log | sort -u | head -3
