set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu55.log; tail -8 gpurun_out/pytest_gpu55.log
