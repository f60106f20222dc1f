set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -k "rows or deep or kernel_choice or midsize or partial or slab or T0_traj or thermal_traj" 2>&1 | tail -15
for O in '{"rows_mode":0}' '{"rows_mode":1}' '{"rows_mode":0,"tile_y":8}' '{"rows_mode":1,"tile_y":8}' '{"rows_warps":4,"tile_y":8}' '{"rows_warps":4,"rows_mode":0,"tile_y":8}' '{"rows_mode":1,"tile_y":12}'; do
timeout 300 python scripts/profile_workload.py c4 128 10 0 "$O" 2>&1 | grep -v "^$" | tail -2 | cut -c1-330
done
timeout 300 python scripts/profile_workload.py c4 128 10 300 2>&1 | grep -v "^$" | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_rows -s 6 -c 2 -f -o gpurun_out/r02n_c4_T0 python scripts/profile_workload.py c4 128 6 0 > gpurun_out/r02n_ncu_c4.log 2>&1; tail -3 gpurun_out/r02n_ncu_c4.log
