set -x
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -x -k "slab or plan or kernel_choice" 2>&1 | tail -3
timeout 300 python scripts/profile_workload.py c2 128 50 300 2>&1 | grep -v "^$" | tail -3 | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 6 -c 2 -f -o gpurun_out/r02s_c2_T300 python scripts/profile_workload.py c2 128 6 300 > gpurun_out/r02s_ncu_c2.log 2>&1; tail -2 gpurun_out/r02s_ncu_c2.log
for O in '{"tile_y":8,"tile_z":32}' '{"tile_y":2,"tile_z":128}' '{"ctas_per_sm":1}' '{"tile_y":4,"tile_z":64,"chunks":8}' '{"recover_u":0}'; do
timeout 300 python scripts/profile_workload.py c2 128 50 300 "$O" 2>&1 | grep -v "^$" | tail -2 | cut -c1-330
done
timeout 300 python scripts/profile_workload.py c2 128 50 0 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
