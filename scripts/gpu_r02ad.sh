set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -4
timeout 600 python scripts/other_configs.py 2>&1 | grep -v "^$" > gpurun_out/r02ad_other_configs.log; cat gpurun_out/r02ad_other_configs.log
timeout 300 python scripts/profile_workload.py c2rk4 64 100 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
