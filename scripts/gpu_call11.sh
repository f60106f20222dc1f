set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity11.log 2>&1; cat gpurun_out/sanity11.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu11.log; tail -15 gpurun_out/pytest_gpu11.log
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench11.log 2>&1; cat gpurun_out/quick_bench11.log
