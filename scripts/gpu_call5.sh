set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity5.log 2>&1; cat gpurun_out/sanity5.log
if grep -q "CRASHED" gpurun_out/sanity5.log; then
  timeout 300 compute-sanitizer --tool memcheck python scripts/sanity_tile.py '{"kernel": 1, "u_tma": 1}' 0.0 2>&1 | grep -v "Host Frame" | head -40 > gpurun_out/sanitizer5.log; cat gpurun_out/sanitizer5.log
fi
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu5.log; tail -15 gpurun_out/pytest_gpu5.log
timeout 1500 python scripts/quick_bench.py 256 100 > gpurun_out/quick_bench5.log 2>&1; cat gpurun_out/quick_bench5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_tile -s 4 -c 2 -o gpurun_out/prof_tile_r01c_spt1 -f python scripts/profile_step.py 256 4 1 100 '{"tile_y":8,"tile_z":64,"spt":1,"ring":4,"ring_u":2}' > gpurun_out/ncu5a.log 2>&1; tail -2 gpurun_out/ncu5a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_tile -s 4 -c 2 -o gpurun_out/prof_tile_r01c_spt2 -f python scripts/profile_step.py 256 4 1 100 '{"tile_y":8,"tile_z":64,"spt":2,"ring":4,"ring_u":2}' > gpurun_out/ncu5b.log 2>&1; tail -2 gpurun_out/ncu5b.log
