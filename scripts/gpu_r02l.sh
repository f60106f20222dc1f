set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 6 -c 2 -f -o gpurun_out/r02l_c2_T300 python scripts/profile_workload.py c2 128 6 300 > gpurun_out/r02l_ncu_c2.log 2>&1; tail -3 gpurun_out/r02l_ncu_c2.log
