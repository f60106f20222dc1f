set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity12.log 2>&1; cat gpurun_out/sanity12.log
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu12.log; tail -15 gpurun_out/pytest_gpu12.log
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench12.log 2>&1; cat gpurun_out/quick_bench12.log
