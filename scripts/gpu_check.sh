set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu52.log; tail -6 gpurun_out/pytest_gpu52.log
timeout 900 python scripts/other_configs.py > gpurun_out/other_configs52.log 2>&1; cat gpurun_out/other_configs52.log
export JB_QB_EXTRA='[{"recover_u":0},{"recover_u":1}]'
timeout 600 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench52.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench52.log
timeout 600 python bench.py --strong --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_strong_1gpu.json 2> gpurun_out/bench_strong_1gpu.err; cut -c1-900 gpurun_out/bench_strong_1gpu.json; tail -2 gpurun_out/bench_strong_1gpu.err
