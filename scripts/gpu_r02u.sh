set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -k "T0_traj or thermal_traj or midsize or partial or slab or bench_code or functional or vacancies" 2>&1 | tail -3
timeout 300 python scripts/profile_workload.py c2 128 50 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 300 python scripts/profile_workload.py c2 128 50 0 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 300 python scripts/profile_workload.py c2 64 100 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 300 python scripts/profile_workload.py c3 256 50 100 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 6 -c 2 -f -o gpurun_out/r02u_c2_T300 python scripts/profile_workload.py c2 128 6 300 > gpurun_out/r02u_ncu_c2.log 2>&1; tail -2 gpurun_out/r02u_ncu_c2.log
