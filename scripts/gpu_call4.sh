set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu4.log; tail -15 gpurun_out/pytest_gpu4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_tile -s 4 -c 2 -o gpurun_out/prof_tile_r01b -f python scripts/profile_step.py 256 4 1 100 > gpurun_out/ncu_tile4.log 2>&1; tail -3 gpurun_out/ncu_tile4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_direct -s 4 -c 2 -o gpurun_out/prof_direct_r01b -f python scripts/profile_step.py 256 4 0 100 > gpurun_out/ncu_direct4.log 2>&1; tail -3 gpurun_out/ncu_direct4.log
