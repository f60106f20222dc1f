set -x
timeout 600 python scripts/sanity_tile.py > gpurun_out/sanity3.log 2>&1; cat gpurun_out/sanity3.log
if grep -q "CRASHED" gpurun_out/sanity3.log; then
  timeout 300 compute-sanitizer --tool memcheck python scripts/sanity_tile.py '{"kernel": 1, "u_tma": 1}' 0.0 2>&1 | grep -v "Host Frame" | head -40 > gpurun_out/sanitizer3.log; cat gpurun_out/sanitizer3.log
fi
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu3.log; tail -40 gpurun_out/pytest_gpu3.log
timeout 1200 python scripts/quick_bench.py 256 > gpurun_out/quick_bench3.log 2>&1; cat gpurun_out/quick_bench3.log
