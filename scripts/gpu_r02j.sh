set -x
timeout 300 python scripts/profile_workload.py c2 128 20 300 2>&1 | grep -v "^$" | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 6 -c 2 -f -o gpurun_out/r02j_c2_T300 python scripts/profile_workload.py c2 128 6 300 > gpurun_out/r02j_ncu_c2.log 2>&1; tail -3 gpurun_out/r02j_ncu_c2.log
timeout 300 python scripts/profile_workload.py c4 128 10 0 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage -s 6 -c 2 -f -o gpurun_out/r02j_c4_T0 python scripts/profile_workload.py c4 128 6 0 > gpurun_out/r02j_ncu_c4.log 2>&1; tail -3 gpurun_out/r02j_ncu_c4.log
