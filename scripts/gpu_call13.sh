set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu13.log; tail -15 gpurun_out/pytest_gpu13.log
export JB_QB_EXTRA='[{"ring":4},{"ring":4,"debug_skip":1},{"ring":4,"debug_skip":2},{"ring":6,"ring_u":2,"tile_y":8,"tile_z":64,"ctas_per_sm":1},{"ring":6,"debug_skip":1},{"ring":6,"debug_skip":2},{"ring":4,"smem_pad":16},{"ring":4,"smem_pad":32},{"ring":4,"ctas_per_sm":1},{"ring":4,"tile_y":7},{"ring":4,"chunks":2},{"ring":4,"chunks":4},{"ring":4,"tile_y":4,"tile_z":128},{"ring":4,"tile_y":2,"tile_z":256},{"ring":4,"tile_y":4,"tile_z":256,"ctas_per_sm":1},{"ring":4,"tile_y":16,"tile_z":32},{"ring":4,"tile_y":6,"tile_z":64},{"ring":4,"tile_y":5,"tile_z":64}]'
timeout 1500 python scripts/quick_bench.py 256 0 > gpurun_out/quick_bench13.log 2>&1; cat gpurun_out/quick_bench13.log
for R in 4 6; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 4 -c 2 -f -o gpurun_out/pair_R${R}_T0 python scripts/profile_step.py 256 4 2 0 "{\"ring\":$R}" > gpurun_out/ncu13_R${R}.log 2>&1; tail -3 gpurun_out/ncu13_R${R}.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 4 -c 2 -f -o gpurun_out/pair_R4_T100 python scripts/profile_step.py 256 4 2 100 '{"ring":4}' > gpurun_out/ncu13_R4T.log 2>&1; tail -3 gpurun_out/ncu13_R4T.log
ls -la gpurun_out/
