set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity7.log 2>&1; cat gpurun_out/sanity7.log
timeout 1500 python scripts/quick_bench.py 256 100 > gpurun_out/quick_bench7.log 2>&1; cat gpurun_out/quick_bench7.log
