set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity14.log 2>&1; cat gpurun_out/sanity14.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu14.log; tail -25 gpurun_out/pytest_gpu14.log
export JB_QB_EXTRA='[{"kernel":3},{"kernel":3,"ring":4},{"kernel":3,"ring":5},{"kernel":3,"tile_y":4,"tile_z":64},{"kernel":3,"tile_y":8,"tile_z":32},{"kernel":3,"chunks":4},{"kernel":3,"chunks":16},{"kernel":2,"ring":4}]'
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench14.log 2>&1; cat gpurun_out/quick_bench14.log
timeout 300 scripts/stream_probe > gpurun_out/stream_probe14.log 2>&1; cat gpurun_out/stream_probe14.log
