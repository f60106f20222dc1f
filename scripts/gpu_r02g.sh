set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; cat gpurun_out/r02g_bench.json; tail -5 gpurun_out/r02g_bench.err
