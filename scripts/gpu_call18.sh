set -x
export JB_QB_DIMS=256x512x128
export JB_QB_EXTRA='[{"kernel":2,"ring":4,"tile_y":8,"tile_z":64},{"kernel":2,"ring":4,"tile_y":4,"tile_z":128},{"kernel":2,"ring":4,"tile_y":2,"tile_z":128},{"kernel":2,"ring":4,"tile_y":8,"tile_z":64,"debug_skip":12},{"kernel":2,"ring":4,"tile_y":4,"tile_z":128,"debug_skip":12},{"kernel":2,"ring":4,"tile_y":2,"tile_z":128,"debug_skip":12},{"kernel":2,"ring":5,"tile_y":4,"tile_z":128,"debug_skip":12},{"kernel":3}]'
timeout 1500 python scripts/quick_bench.py 256 0 > gpurun_out/quick_bench18.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench18.log
export JB_QB_DIMS=512x512x64
export JB_QB_EXTRA='[{"kernel":2,"ring":4,"tile_y":8,"tile_z":64},{"kernel":2,"ring":4,"tile_y":8,"tile_z":64,"debug_skip":12},{"kernel":2,"ring":4,"tile_y":4,"tile_z":64,"debug_skip":12}]'
timeout 1500 python scripts/quick_bench.py 256 0 > gpurun_out/quick_bench18b.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench18b.log
