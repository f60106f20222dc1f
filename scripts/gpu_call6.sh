set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity6.log 2>&1; cat gpurun_out/sanity6.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu6.log; tail -5 gpurun_out/pytest_gpu6.log
timeout 1500 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench6.log 2>&1; cat gpurun_out/quick_bench6.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_tile -s 4 -c 2 -o gpurun_out/prof_tile_r01d_spt1 -f python scripts/profile_step.py 256 4 1 100 '{"tile_y":8,"tile_z":64,"spt":1,"ring":4,"ring_u":2}' > gpurun_out/ncu6a.log 2>&1; tail -2 gpurun_out/ncu6a.log
