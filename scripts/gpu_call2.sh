set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu2.log; tail -30 gpurun_out/pytest_gpu2.log
if grep -q "failed" gpurun_out/pytest_gpu2.log; then
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "extension_is_loaded or T0_trajectories" 2>&1 | grep -v "^$" | head -80 > gpurun_out/sanitizer2.log; head -60 gpurun_out/sanitizer2.log
fi
timeout 1200 python scripts/quick_bench.py 256 > gpurun_out/quick_bench2.log 2>&1; cat gpurun_out/quick_bench2.log
