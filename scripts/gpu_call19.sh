set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity19.log 2>&1; head -12 gpurun_out/sanity19.log
export JB_QB_EXTRA='[{"kernel":2,"ring":4},{"kernel":2,"ring":4,"row_offset":8},{"kernel":2,"ring":4,"row_offset":4},{"kernel":2,"ring":4,"row_offset":8,"debug_skip":12},{"kernel":2,"ring":4,"row_offset":4,"debug_skip":12},{"kernel":2,"ring":4,"row_offset":4,"tile_y":4,"tile_z":128},{"kernel":2,"ring":5,"row_offset":4},{"kernel":3,"row_offset":4},{"kernel":3,"row_offset":8}]'
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench19.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench19.log
