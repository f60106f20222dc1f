set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu32.log; tail -6 gpurun_out/pytest_gpu32.log
timeout 900 python scripts/other_configs.py > gpurun_out/other_configs32.log 2>&1; cat gpurun_out/other_configs32.log
