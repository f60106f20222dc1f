set -x
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 -x -k "rk4 or slab or biquadratic or pinned" 2>&1 | tail -5
timeout 300 python scripts/profile_workload.py c3rk4 256 20 100 2>&1 | grep -v "^$" | tail -1 | cut -c1-330
timeout 300 python scripts/profile_workload.py c3rk4 256 20 0 2>&1 | grep -v "^$" | tail -1 | cut -c1-330
timeout 300 python scripts/profile_workload.py c2rk4 128 20 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 8 -c 4 -f -o gpurun_out/r02w_rk4_c3_T100 python scripts/profile_workload.py c3rk4 256 4 100 > gpurun_out/r02w_ncu_rk4.log 2>&1; tail -2 gpurun_out/r02w_ncu_rk4.log
