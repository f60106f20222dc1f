set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu25.log; tail -25 gpurun_out/pytest_gpu25.log
