set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity15.log 2>&1; head -8 gpurun_out/sanity15.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu15.log; tail -25 gpurun_out/pytest_gpu15.log
export JB_QB_EXTRA='[{"kernel":3},{"kernel":3,"ring":4},{"kernel":3,"chunks":8},{"kernel":3,"chunks":12},{"kernel":3,"tile_y":4,"tile_z":64},{"kernel":3,"tile_y":6,"tile_z":64},{"kernel":2,"ring":4}]'
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench15.log 2>&1; cat gpurun_out/quick_bench15.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_fused -s 3 -c 1 -f -o gpurun_out/fused_T0 python scripts/profile_step.py 256 5 3 0 > gpurun_out/ncu15_T0.log 2>&1; tail -3 gpurun_out/ncu15_T0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_fused -s 3 -c 1 -f -o gpurun_out/fused_T100 python scripts/profile_step.py 256 5 3 100 > gpurun_out/ncu15_T100.log 2>&1; tail -3 gpurun_out/ncu15_T100.log
