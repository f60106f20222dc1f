set -x
timeout 1800 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -40 > gpurun_out/r02h_pytest.log; tail -40 gpurun_out/r02h_pytest.log
