set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity10.log 2>&1; cat gpurun_out/sanity10.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu10.log; tail -15 gpurun_out/pytest_gpu10.log
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench10.log 2>&1; cat gpurun_out/quick_bench10.log
